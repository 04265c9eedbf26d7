// Microbenchmark (diagnostic): what in the MMA issuer's k-loop serialises with the tensor pipe?  One CTA, operands resident
// in shared memory, N = 64 (48 clk per MMA when issued back to back).  Per "k-block" of four MMAs the loop optionally adds
//   bit 0: tcgen05.commit -> mbarrier (never waited on inside the loop)
//   bit 1: tcgen05.fence::after_thread_sync
//   bit 2: an mbarrier.try_wait on a barrier whose phase is already complete
//   bit 3: the elect_one / __syncwarp structure of gemm_tc.cu (otherwise lane 0 issues everything in one block)
#include "../../image-captioning-with-different-decoders_b200/csrc/tc_common.cuh"
#include <cstdio>
void icd_set_error(const char*, ...) {}
long long g_icd_launches = 0;

template <int BN>
__global__ void __launch_bounds__(128, 1) mma_loop_kernel(int n_kb, int mode, long long* out) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* g = smem_raw + (base - smem_u32(smem_raw));
    const uint32_t sA = base, sB = base + 16384, bar = base + 16384 + BN * 128, bar_done = bar + 8, bar_end = bar + 16;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(g + 16384 + BN * 128 + 64);
    for (int i = threadIdx.x; i < (16384 + BN * 128) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(g)[i] = 0x3c003c00u;
    const int warp = (int)uniform_u32(threadIdx.x >> 5);
    if (threadIdx.x == 0) {
        mbar_init(bar, 1 << 20); mbar_init(bar_done, 1); mbar_init(bar_end, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        mbar_arrive(bar_done);                       // phase 0 of bar_done is complete: try_wait(parity 0) succeeds at once
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = uniform_u32(*tmem_slot);
    if (warp == 1) {
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        long long t0 = clock64();
        if (mode & 16) {                 // bit 4: ONE elected thread runs the whole k-loop (no reconvergence per k-block)
            if (elect_one()) {
                uint32_t stage = 0;
                for (int kb = 0; kb < n_kb; ++kb) {
                    if (mode & 4) mbar_wait(bar_done, 0);
                    if (mode & 2) tc_fence_after();
                    const uint64_t adesc = make_smem_desc(sA + (stage & 1) * 0, 0), bdesc = make_smem_desc(sB, 0);
#pragma unroll
                    for (int k = 0; k < 4; ++k) tc_mma_f16(tmem_base, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, 1u);
                    if (mode & 1) tc_commit(bar);
                    if (kb == n_kb - 1) tc_commit(bar_end);
                    if (++stage == 8) stage = 0;
                }
            }
            __syncwarp();
        } else
        for (int kb = 0; kb < n_kb; ++kb) {
            if (mode & 4) mbar_wait(bar_done, 0);
            if (mode & 2) tc_fence_after();
            if (elect_one()) {
                const uint64_t adesc = make_smem_desc(sA, 0), bdesc = make_smem_desc(sB, 0);
#pragma unroll
                for (int k = 0; k < 4; ++k) tc_mma_f16(tmem_base, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, 1u);
                if (mode & 1) tc_commit(bar);
                if (kb == n_kb - 1) tc_commit(bar_end);
            }
            if (mode & 8) __syncwarp();
        }
        long long t1 = clock64();
        mbar_wait(bar_end, 0);
        long long t2 = clock64();
        if ((threadIdx.x & 31) == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "r"(512u) : "memory");
    }
}

int main() {
    long long* d; cudaMalloc(&d, 16);
    const int n_kb = 1024;
    const size_t smem = 1024 + 16384 + 64 * 128 + 256;
    cudaFuncSetAttribute(mma_loop_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    for (int mode : {0, 7, 15, 16, 17, 19, 23}) {
        mma_loop_kernel<64><<<1, 128, smem>>>(n_kb, mode, d);
        cudaError_t e = cudaDeviceSynchronize();
        long long h[2] = {0, 0}; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
        printf("mode %2d [%s%s%s%s%s]: issue loop %7.1f clk per k-block, until the last MMA retired %7.1f clk per k-block (%s)\n", mode,
               mode & 1 ? "commit " : "", mode & 2 ? "fence " : "", mode & 4 ? "try_wait " : "", mode & 8 ? "syncwarp " : "", mode & 16 ? "single-thread-loop " : "",
               (double)h[0] / n_kb, (double)h[1] / n_kb, cudaGetErrorString(e));
    }
    return 0;
}
