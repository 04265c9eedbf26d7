"""Calibration only (not on the product path): what the vendor library reaches on the in-loop contraction shapes, warm L2,
CUDA-graph replay — the bar for csrc/gemm_tc.cu on the M = batch shapes.   python tools/cublas_calibration.py"""
import torch
dev = torch.device("cuda:0")
shapes = [("K2 z", 512, 4608, 512), ("K4 gates", 512, 2048, 2048), ("d_gated", 512, 2048, 2048), ("dh", 512, 512, 4608),
          ("K1 enc_att", 100352, 512, 2048), ("fc fwd", 12288, 9490, 512)]
for name, M, N, K in shapes:
    a = torch.randn(M, K, device=dev, dtype=torch.bfloat16)
    b = torch.randn(N, K, device=dev, dtype=torch.bfloat16)
    out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    iters = 50 if M <= 512 else 5
    for _ in range(3):
        torch.matmul(a, b.t(), out=out)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(iters):
            torch.matmul(a, b.t(), out=out)
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        g.replay()
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / (5 * iters)
    print("%-12s M=%6d N=%5d K=%5d  cuBLAS bf16 %7.1f us %7.1f TF/s" % (name, M, N, K, us, 2.0 * M * N * K / us / 1e6), flush=True)
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for name, M, N, K in shapes[:4]:
        a = torch.randn(M, K, device=dev, dtype=torch.bfloat16); b = torch.randn(N, K, device=dev, dtype=torch.bfloat16)
        torch.matmul(a, b.t())
    torch.cuda.synchronize()
for ev in prof.key_averages():
    if "gemm" in ev.key.lower() or "cutlass" in ev.key.lower() or "nvjet" in ev.key.lower():
        print(ev.key[:150], "%.1f us" % ev.device_time_total)
