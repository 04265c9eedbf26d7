"""Attention-step kernel timing vs number of rows per launch (HBM efficiency at reduced grid sizes)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import __graft_entry__  # noqa: E402

__graft_entry__.build()
from icd_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
P, C, A = 196, 2048, 512
B = 1184
enc16 = torch.randn(B, P, C, device=dev).clamp_min_(0).bfloat16()
att16 = (torch.randn(B, P, A, device=dev) * 0.5).bfloat16()
att_dec = torch.randn(B, A, device=dev) * 0.5
wf = torch.randn(A, device=dev) * 0.2
bf = torch.zeros(1, device=dev)
fb = torch.randn(B, C, device=dev)
FEW = "--few-rows" in sys.argv      # rows-per-launch sweep under the row-sharing variants (ICD_ATT_FWD_SPLIT / ICD_ATT_BWD_SPLIT)


def timed(fn, los, reps=10, warm=2):
    for _ in range(warm):
        for lo in los:
            fn(lo)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        for lo in los:
            fn(lo)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / (reps * len(los))


if FEW:
    alpha, awe, gate, gated, gated16 = ops.attention_step_fwd_bf16(enc16, att16, att_dec, wf, bf, fb)
    d_gated = torch.randn(B, C, device=dev)
    print("rows | fwd us: 1 CTA/row, 2, 4 | bwd us: whole rows, balance rule, 2 CTAs/row")
    for rows in (16, 32, 64, 100, 148, 200, 222, 256, 296, 298, 320, 338, 370, 400, 444, 468, 480, 512):
        los = [lo for lo in range(0, B, rows) if lo + rows <= B]
        out = []
        for mode in ("1", "2", "4"):
            os.environ["ICD_ATT_FWD_SPLIT"] = mode
            out.append(timed(lambda lo: ops.attention_step_fwd_bf16(enc16[lo:lo + rows], att16[lo:lo + rows], att_dec[lo:lo + rows],
                                                                    wf, bf, fb[lo:lo + rows]), los))
        os.environ.pop("ICD_ATT_FWD_SPLIT", None)
        for mode in ("0", None, "2"):
            if mode is None:
                os.environ.pop("ICD_ATT_BWD_SPLIT", None)
            else:
                os.environ["ICD_ATT_BWD_SPLIT"] = mode
            out.append(timed(lambda lo: ops.attention_step_bwd_bf16(enc16[lo:lo + rows], att16[lo:lo + rows], att_dec[lo:lo + rows], wf,
                                                                    alpha[lo:lo + rows], gate[lo:lo + rows], awe[lo:lo + rows],
                                                                    d_gated[lo:lo + rows], None), los))
        os.environ.pop("ICD_ATT_BWD_SPLIT", None)
        print("%4d | %6.1f %6.1f %6.1f | %6.1f %6.1f %6.1f" % (rows, *out), flush=True)
    sys.exit(0)

for rows in (592, 574, 512, 444, 296, 256, 148):
    def run(lo):
        return ops.attention_step_fwd_bf16(enc16[lo:lo + rows], att16[lo:lo + rows], att_dec[lo:lo + rows], wf, bf, fb[lo:lo + rows])
    los = [lo for lo in range(0, B, rows) if lo + rows <= B]
    for _ in range(3):
        for lo in los:
            run(lo)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 0
    e0.record()
    for _ in range(10):
        for lo in los:
            run(lo); n += 1
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / n
    print("fwd rows=%3d  %.1f us per launch  %.0f GB/s" % (rows, us, rows * 1022736 / us / 1e3))

# backward kernel at the same row counts
alpha, awe, gate, gated, gated16 = ops.attention_step_fwd_bf16(enc16, att16, att_dec, wf, bf, fb)
d_gated = torch.randn(B, C, device=dev)
for rows in (592, 574, 512, 480, 444, 400, 320, 296, 290, 256, 200, 148):
    def runb(lo):
        sl = slice(lo, lo + rows)
        return ops.attention_step_bwd_bf16(enc16[sl], att16[sl], att_dec[sl], wf, alpha[sl], gate[sl], awe[sl], d_gated[sl], None)
    los = [lo for lo in range(0, B, rows) if lo + rows <= B]
    for _ in range(2):
        for lo in los:
            runb(lo)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 0
    e0.record()
    for _ in range(10):
        for lo in los:
            runb(lo); n += 1
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / n
    print("bwd rows=%3d  %.1f us per launch  %.0f GB/s (incl. the wrapper's output allocations)" % (rows, us, rows * 1042736 / us / 1e3))
