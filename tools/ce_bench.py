"""Diagnostic: time of the cross-entropy kernels on the (B*T, V) = (12288, 9490) logits of the timed configuration."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import __graft_entry__
__graft_entry__.build()
from icd_b200 import ops
dev = torch.device("cuda:0")
R, V = 12288, 9490
x = torch.randn(R, V, device=dev) * 3
t = torch.randint(0, V, (R,), device=dev)
def timeit(f, iters=20):
    for _ in range(3): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / iters
rl, lse = ops.cross_entropy_fwd(x, t)
print("fwd (loss + lse)            %7.1f us   (%.0f MB read)" % (timeit(lambda: ops.cross_entropy_fwd(x, t)), R * V * 4 / 1e6))
print("fwd + bf16 gradient, 1 pass %7.1f us   (%.0f MB read + %.0f MB written)" % (timeit(lambda: ops.cross_entropy_fwd_grad16(x, t, 1.0 / R)), R * V * 4 / 1e6, R * 9496 * 2 / 1e6))
print("bwd fp32 + bf16             %7.1f us" % timeit(lambda: ops.cross_entropy_bwd(x, t, lse, 1.0 / R, upstream=torch.ones(1, device=dev), want_bf16=True)))
print("bwd bf16 only               %7.1f us" % timeit(lambda: ops.cross_entropy_bwd(x, t, lse, 1.0 / R, upstream=torch.ones(1, device=dev), want_bf16=True, want_fp32=False)))
