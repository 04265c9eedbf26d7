"""Diagnostic: phase boundaries (globaltimer, ns) of every CTA of one attention-step backward launch at B = 512.
Needs a library built with NVCC_EXTRA=-DICD_ATT_TRACE."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import __graft_entry__
__graft_entry__.build()
from icd_b200 import ops, _lib
dev = torch.device("cuda:0")
R, P, C, A = 512, 196, 2048, 512
g = torch.Generator().manual_seed(1)
enc16 = torch.randn(R, P, C, generator=g).clamp_min_(0).bfloat16().to(dev)
att_enc16 = (torch.randn(R, P, A, generator=g) * 0.5).bfloat16().to(dev)
att_dec = (torch.randn(R, A, generator=g) * 0.5).to(dev)
wf = (torch.randn(A, generator=g) * 0.2).to(dev); bf = torch.randn(1, generator=g).to(dev)
fb = torch.randn(R, C, generator=g).to(dev)
alpha, awe, gate, gated, gated16 = ops.attention_step_fwd_bf16(enc16, att_enc16, att_dec, wf, bf, fb, None)
d_gated = torch.randn(R, C, generator=g).to(dev)
for _ in range(3):
    ops.attention_step_bwd_bf16(enc16, att_enc16, att_dec, wf, alpha, gate, awe, d_gated, None)
torch.cuda.synchronize()
buf = (ctypes.c_longlong * (8 * 1024))()
_lib.lib().icd_att_trace_read(buf)
t = np.array(buf[:], dtype=np.int64).reshape(8, 1024)[:5, :R].astype(np.float64)
t0 = t[0].min()
names = ["start", "A done (gate adjoint)", "B done (d_alpha over enc)", "C done (softmax bwd)", "D done (masks over att_enc)"]
print("ns since the first CTA started: min / median / max over the %d CTAs" % R)
for i, n in enumerate(names):
    print("%-30s %8.0f %8.0f %8.0f" % (n, t[i].min() - t0, np.median(t[i]) - t0, t[i].max() - t0))
d = np.diff(t, axis=0)
print("phase durations (median ns): A %.0f  B %.0f  C %.0f  D %.0f" % tuple(np.median(d, axis=1)))
