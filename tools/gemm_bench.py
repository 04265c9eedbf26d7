"""Isolated timing of the tcgen05 contraction on the shapes of the train step (CUDA events, warm L2).
    python tools/gemm_bench.py            # on the GPU box
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import __graft_entry__  # noqa: E402

__graft_entry__.build()
from icd_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
GRAPH = "--graph" in sys.argv              # time a CUDA-graph replay of the launch chain (removes the Python launch cost)
PROFILE = "--profile" in sys.argv          # one warm-up + one launch per shape (for ncu --set full)
B, T, P, C, A, D, E, V = 512, 24, 196, 2048, 512, 512, 512, 9490
NZ = A + C + 4 * D


def run(name, M, N, K, a_mn=False, b_mn=False, fp32=True, bf16=False, mask=False, bias=False, add=False, ldc=None, iters=20):
    a = torch.randn((K, M) if a_mn else (M, K), device=dev).bfloat16()
    b = torch.randn((K, N) if b_mn else (N, K), device=dev).bfloat16()
    ldc = ldc or N
    out = torch.empty(M, ldc, device=dev) if fp32 else None
    out16 = torch.empty(M, (N + 7) // 8 * 8, device=dev, dtype=torch.bfloat16) if bf16 else None
    kw = dict(a_mn=a_mn, b_mn=b_mn, out=out, out16=out16, want_fp32=fp32, ldc=ldc,
              bias1=torch.randn(N, device=dev) if bias else None,
              add1=torch.randn(M, N, device=dev) if add else None,
              row_mask=torch.ones(M, device=dev, dtype=torch.uint8) if mask else None)
    if PROFILE:
        iters = 1
    for _ in range(1 if PROFILE else 3):
        ops.gemm_bf16(a, b, M, N, K, **kw)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if GRAPH:                                   # replay a captured chain: no host launch overhead between the kernels
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(iters):
                ops.gemm_bf16(a, b, M, N, K, **kw)
        g.replay(); torch.cuda.synchronize()
        e0.record()
        for _ in range(5):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / (5 * iters)
    else:
        e0.record()
        for _ in range(iters):
            ops.gemm_bf16(a, b, M, N, K, **kw)
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / iters
    print("%-34s M=%6d N=%5d K=%6d  %8.1f us  %7.1f TF/s" % (name, M, N, K, us, 2.0 * M * N * K / us / 1e6), flush=True)


if "--fc" in sys.argv:
    # the vocabulary layer (K6) and the hoisted input projection (K5): short K, fp32 output — tile width, pairing, multicast clusters
    for name, M, N, K, kw in [("K6 fc fwd (ldc=V, mask)", B * T, V, D, dict(bias=True, mask=True)),
                              ("K6 fc fwd (ldc=9496)", B * T, V, D, dict(bias=True, mask=True, ldc=9496)),
                              ("K5 emb*W_ih", B * T, 4 * D, E, dict(bias=True))]:
        for pair, plan, cl in [("0", "256,1", "1,1"), ("0", "128,1", "1,1"), ("2", "256,1", "1,1"), ("2", "128,1", "1,1"),
                               ("0", "256,1", "2,1"), ("0", "256,1", "1,2"), ("0", "256,1", "2,2"), ("0", "128,1", "2,2"), ("0", "256,1", "4,1"),
                               ("0", "256,1", "4,2")]:
            os.environ["ICD_GEMM_FORCE_PLAN"] = plan
            os.environ["ICD_GEMM_CLUSTER"] = cl
            ops.gemm_set_pair_mode(int(pair))
            run("%s pair %s plan %s cluster %s" % (name, pair, plan, cl), M, N, K, iters=20, **kw)
    sys.exit(0)
if "--fck" in sys.argv:
    # K-slope of the vocabulary-layer shape: the K -> 0 intercept is the epilogue + store time of the real kernel
    for ldc in (V, 9496, 9600):
        for K_ in (64, 128, 256, 512, 1024):
            run("fc shape ldc=%d" % ldc, B * T, V, K_, bias=True, mask=True, ldc=ldc, iters=10)
    for K_ in (64, 256, 512):
        run("K5 shape", B * T, 4 * D, K_, bias=True, iters=20)
    sys.exit(0)
if "--clusters" in sys.argv:
    # sweep of the cm x cn multicast clusters and tile / split-K plans on the in-loop (M = batch) contractions
    shapes = [("K2 z", B, NZ, D, dict(bias=True), ["128,1", "64,1", "256,1"]),
              ("K4 gates", B, 4 * D, C, dict(add=True), ["64,1", "128,1", "128,2", "256,2", "256,4"]),
              ("d_gated", B, C, 4 * D, dict(b_mn=True), ["64,1", "128,1", "128,2", "256,2", "256,4"]),
              ("dh", B, D, NZ, dict(b_mn=True), ["128,8", "64,4", "128,4", "64,8", "256,8", "256,16"])]
    for name, M, N, K, kw, plans in shapes:
        for plan in plans:
            for cl in ["1,1", "2,1", "1,2", "2,2", "4,1", "4,2", "2,4", "1,4"]:
                os.environ["ICD_GEMM_FORCE_PLAN"] = plan
                os.environ["ICD_GEMM_CLUSTER"] = cl
                os.environ["ICD_GEMM_PAIR"] = "0"
                run("%s plan %s cluster %s" % (name, plan, cl), M, N, K, iters=100, **kw)
    sys.exit(0)
if "--plans" in sys.argv:
    # tile / split-K / pairing sweep on the in-loop (M = batch) contractions: the data the cost model of make_plan is fitted to
    pair = os.environ.get("ICD_GEMM_PAIR", "0")
    shapes = [("K2 z", B, NZ, D, dict(bias=True)), ("K4 gates", B, 4 * D, C, dict(add=True)),
              ("d_gated", B, C, 4 * D, dict(b_mn=True)), ("dh", B, D, NZ, dict(b_mn=True)), ("h_lin", B, D, C, dict(bias=True, bf16=True))]
    for name, M, N, K, kw in shapes:
        for bn in [64, 128, 256]:
            for sp in [1, 2, 3, 4, 6, 8, 12, 16]:
                nkb = (K + 63) // 64
                if sp > 1 and (nkb // sp < 2 or ((M + 127) // 128) * ((N + bn - 1) // bn) * sp > 160):
                    continue
                os.environ["ICD_GEMM_FORCE_PLAN"] = "%d,%d" % (bn, sp)
                run("pair %s %s plan %d,%d" % (pair, name, bn, sp), M, N, K, iters=100, **kw)
    sys.exit(0)
if "--kslope" in sys.argv:
    # time against K at fixed M x N and plan: slope = cost of one k-block, intercept = launch + prologue + epilogue
    os.environ["ICD_GEMM_PAIR"] = "0"
    for plan in ["64,1", "128,1", "256,1"]:
        os.environ["ICD_GEMM_FORCE_PLAN"] = plan
        for K in [64, 256, 512, 1024, 2048, 4096, 8192]:
            run("plan %s" % plan, B, 4 * D, K, iters=100)
    for K in [64, 256, 512, 1024, 2048, 4096, 8192]:
        a = torch.randn(B, K, device=dev, dtype=torch.bfloat16); b = torch.randn(4 * D, K, device=dev, dtype=torch.bfloat16)
        out = torch.empty(B, 4 * D, device=dev, dtype=torch.bfloat16)
        for _ in range(3):
            torch.matmul(a, b.t(), out=out)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(100):
                torch.matmul(a, b.t(), out=out)
        g.replay(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            g.replay()
        e1.record(); torch.cuda.synchronize()
        print("cuBLAS (calibration) K=%5d  %7.1f us" % (K, e0.elapsed_time(e1) * 1e3 / 500), flush=True)
    sys.exit(0)
if "--fc" in sys.argv:             # the vocabulary layer's forward contraction only (for ncu)
    PROFILE = "--profile" in sys.argv
    run("K6 fc fwd (ldc=V, mask)", B * T, V, D, bias=True, mask=True)
    sys.exit(0)
if "--one" in sys.argv:            # one shape, for ncu: python tools/gemm_bench.py --one  (K4 gates, plan from the environment)
    PROFILE = True
    run("K4 gates (step)", B, 4 * D, C, add=True)
    sys.exit(0)
if "--dh" in sys.argv:
    run("dh (step)", B, D, NZ, b_mn=True, iters=200)
    run("d_gated (step)", B, C, 4 * D, b_mn=True, iters=200)
    run("K4 gates (step)", B, 4 * D, C, add=True, iters=200)
    run("K2 z (step)", B, NZ, D, bias=True, iters=200)
    sys.exit(0)
run("K1 enc_att (bf16 out)", B * P, A, C, fp32=False, bf16=True, bias=True)
run("K5 emb*W_ih (fp32 out)", T * B, 4 * D, E, bias=True)
run("K6 fc fwd (ldc=V, mask)", B * T, V, D, bias=True, mask=True)
run("K6 fc fwd (ldc=V, no mask)", B * T, V, D, bias=True)
run("K6 fc fwd (ldc=9496)", B * T, V, D, bias=True, mask=True, ldc=9496)
run("d_hdrop dY*Wfc", B * T, D, V - 2, b_mn=True)
run("d_fc_w dY^T*hdrop", V - 2, D, B * T, a_mn=True, b_mn=True)
run("d_w_cat", NZ, D, T * B, a_mn=True, b_mn=True)
run("d_w_ih(C)", 4 * D, C, T * B, a_mn=True, b_mn=True)
run("dW_e", A, C, B * P, a_mn=True, b_mn=True)
run("K2 z (step)", B, NZ, D, bias=True, iters=100)
run("K4 gates (step)", B, 4 * D, C, add=True, iters=100)
run("d_gated (step)", B, C, 4 * D, b_mn=True, iters=100)
run("dh (step)", B, D, NZ, b_mn=True, iters=100)
run("h_lin", B, D, C, bias=True, bf16=True, iters=100)
