"""Diagnostic: per-k-block clock stamps of the TMA producer and the MMA issuer of CTA 0 of one contraction.
Needs a library built with -DICD_GEMM_TRACE:  NVCC_EXTRA=-DICD_GEMM_TRACE python -m ... build --force  (see build.py)."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import __graft_entry__
__graft_entry__.build()
from icd_b200 import ops, _lib
dev = torch.device("cuda:0")
M, N, K = 512, 2048, 2048
os.environ["ICD_GEMM_PAIR"] = "0"
os.environ["ICD_GEMM_FORCE_PLAN"] = sys.argv[1] if len(sys.argv) > 1 else "64,1"
a = torch.randn(M, K, device=dev).bfloat16(); b = torch.randn(N, K, device=dev).bfloat16()
out = torch.empty(M, N, device=dev)
for _ in range(3):
    ops.gemm_bf16(a, b, M, N, K, out=out, want_fp32=True, ldc=N)
torch.cuda.synchronize()
buf = (ctypes.c_longlong * (6 * 512))()
rc = _lib.lib().icd_gemm_trace_read(buf)
t = [[buf[s * 512 + i] for i in range(K // 64)] for s in range(6)]
t0 = t[0][0]
print("kb   A-prod:slot-free  B-prod:slot-free  mma:loop-top  mma:wait-done  mma:elected   mma:issued   (clocks since the first stamp)")
for kb in range(K // 64):
    print("%2d %12d %12d %12d %12d %12d %12d" % (kb, t[0][kb] - t0, t[1][kb] - t0, t[4][kb] - t0, t[5][kb] - t0, t[2][kb] - t0, t[3][kb] - t0))
