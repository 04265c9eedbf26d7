// Microbenchmark (diagnostic, not product): how fast can ONE 128-row x 256-column fp32 output tile per SM be written with the
// access pattern of the contraction epilogue (a warp owns 32 rows x 32 columns at a time; rows ldc floats apart), as a function
// of the number of storing warps per SM, of the row stride and of the store mechanism:
//   mode 0: st.global.v4, lane -> 4 columns, 8 lanes per row, 4 rows per instruction (the epilogue's store loop)
//   mode 1: per-row bulk async copies shared -> global (cp.async.bulk.global.shared::cta, 128 B per row, one per lane)
//   mode 2: st.global.v4, lane -> its own row, 8 stores of 16 B (no transposition: what a direct TMEM -> global path would do)
// 148 CTAs, each walks over tiles exactly like gemm_tc_kernel (M-fastest or N-fastest rasterisation).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o store_pattern.bin store_pattern.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int MODE>
__global__ void __launch_bounds__(512, 1) store_kernel(float* C, long long ldc, int M, int N, int raster_n, int nwarps) {
    extern __shared__ __align__(128) float sm[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp >= nwarps) return;
    const int tiles_m = M / 128, tiles_n = (N + 255) / 256;
    float* sE = sm + warp * 32 * 36;                               // 32 rows x 144 B (padded: conflict-free row writes)
    for (int i = lane; i < 32 * 36; i += 32) sE[i] = (float)i;
    __syncwarp();
    const int q = warp & 3, groups = nwarps / 4, cg = warp >> 2;
    for (int tile = blockIdx.x; tile < tiles_m * tiles_n; tile += gridDim.x) {
        const int tm = raster_n ? tile / tiles_n : tile % tiles_m, tn = raster_n ? tile % tiles_n : tile / tiles_m;
        const int mrow0 = tm * 128 + q * 32;
        for (int c = cg; c < 8; c += groups) {
            const int nb = tn * 256 + c * 32;
            if (nb + 32 > N) continue;
            if (MODE == 0) {
                const int l3 = lane >> 3, j = lane & 7;
                float* dst = C + (long long)(mrow0 + l3) * ldc + nb + 4 * j;
                const float4 v = *reinterpret_cast<const float4*>(sE + l3 * 36 + 4 * j);
#pragma unroll
                for (int i = 0; i < 8; ++i) { *reinterpret_cast<float4*>(dst) = v; dst += 4 * ldc; }
            } else if (MODE == 1) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                float* dst = C + (long long)(mrow0 + lane) * ldc + nb;
                asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], 128;" :: "l"(dst), "r"(smem_u32(sE + lane * 36)) : "memory");
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            } else {
                float* dst = C + (long long)(mrow0 + lane) * ldc + nb;
                const float4 v = *reinterpret_cast<const float4*>(sE + lane * 36);
#pragma unroll
                for (int i = 0; i < 8; ++i) *reinterpret_cast<float4*>(dst + 4 * i) = v;
            }
            __syncwarp();
        }
    }
    if (MODE == 1) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

template <int MODE>
static void run(const char* name, float* C, long long ldc, int M, int N, int raster_n, int nwarps) {
    cudaFuncSetAttribute(store_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 16 * 32 * 36 * 4);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int i = 0; i < 2; ++i) store_kernel<MODE><<<148, 512, 16 * 32 * 36 * 4>>>(C, ldc, M, N, raster_n, nwarps);
    cudaEventRecord(e0);
    const int iters = 10;
    for (int i = 0; i < iters; ++i) store_kernel<MODE><<<148, 512, 16 * 32 * 36 * 4>>>(C, ldc, M, N, raster_n, nwarps);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    const cudaError_t err = cudaGetLastError();
    const double bytes = (double)M * (N / 32 * 32) * 4;
    printf("%-28s ldc %5lld raster_n %d warps %2d : %7.1f us  %6.2f TB/s  %s\n", name, ldc, raster_n, nwarps, ms * 1e3 / iters,
           bytes / (ms / iters * 1e-3) / 1e12, err == cudaSuccess ? "" : cudaGetErrorString(err));
}

int main() {
    const int M = 12288, N = 9490;
    float* C;
    cudaMalloc(&C, (size_t)M * 9496 * 4);
    for (int raster = 0; raster < 2; ++raster)
        for (int nw : {4, 8, 16}) {
            run<0>("st.v4 4 rows x 128 B", C, 9496, M, N, raster, nw);
            run<0>("st.v4 4 rows x 128 B", C, 9492, M, N, raster, nw);
            run<1>("bulk copy 128 B per row", C, 9496, M, N, raster, nw);
            run<1>("bulk copy 128 B per row", C, 9492, M, N, raster, nw);
            run<2>("st.v4 lane = row", C, 9496, M, N, raster, nw);
        }
    run<0>("st.v4 4 rows x 128 B (K5)", C, 2048, M, 2048, 0, 8);
    run<0>("st.v4 4 rows x 128 B (K5)", C, 2048, M, 2048, 1, 8);
    run<1>("bulk copy (K5)", C, 2048, M, 2048, 1, 8);
    return 0;
}
