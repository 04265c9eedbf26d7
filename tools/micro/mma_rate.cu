// Microbenchmark (diagnostic, not product): clocks per tcgen05.mma (M = 128, K = 16, bf16, operands resident in shared
// memory) as a function of N, of the accumulator dependency (1 or 2 accumulators alternating) and of how many MMAs are
// issued between two commits.   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_rate.bin mma_rate.cu
#include "../../image-captioning-with-different-decoders_b200/csrc/tc_common.cuh"
#include <cstdio>

void icd_set_error(const char*, ...) {}
long long g_icd_launches = 0;

template <int BN>
__global__ void __launch_bounds__(128, 1) mma_rate_kernel(int n_mma, int n_acc, int per_commit, int mn_major, long long* out, int commit_every) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* g = smem_raw + (base - smem_u32(smem_raw));
    const uint32_t sA = base, sB = base + 16384, bar = base + 16384 + BN * 128;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(g + 16384 + BN * 128 + 64);
    for (int i = threadIdx.x; i < (16384 + BN * 128) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(g)[i] = 0x3c003c00u;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) { mbar_init(bar, commit_every > 0 ? (uint32_t)(n_mma / commit_every) : 1u); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (warp == 1) {
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(128 >> 4) << 24) |
                               ((uint32_t)mn_major << 15) | ((uint32_t)mn_major << 16);
        const uint64_t kstep = mn_major ? (uint64_t)(2048 >> 4) : (uint64_t)(32 >> 4);
        long long t0 = 0, t1 = 0;
        uint32_t phase = 0;
        if (lane == 0) t0 = clock64();
        int issued = 0;
        while (issued < n_mma) {
            if (lane == 0) {
                const uint64_t adesc = make_smem_desc(sA, mn_major), bdesc = make_smem_desc(sB, mn_major);
                // unrolled by 4 with compile-time operand offsets, so the issue loop is as cheap as in gemm_tc.cu
                const uint32_t d0 = tmem_base, d1 = tmem_base + (uint32_t)((n_acc - 1) * BN);
                for (int j = 0; j < per_commit; j += 4) {
                    tc_mma_f16(d0, adesc, bdesc, idesc, 1u);
                    tc_mma_f16(d1, adesc + kstep, bdesc + kstep, idesc, 1u);
                    tc_mma_f16(d0, adesc + 2 * kstep, bdesc + 2 * kstep, idesc, 1u);
                    tc_mma_f16(d1, adesc + 3 * kstep, bdesc + 3 * kstep, idesc, 1u);
                    if (commit_every > 0 && ((j + 4) % commit_every) == 0 && j + 4 < per_commit) tc_commit(bar);
                }
                tc_commit(bar);
            }
            __syncwarp();
            issued += per_commit;
            if (issued >= n_mma || per_commit >= 64) { mbar_wait(bar, phase); phase ^= 1; }   // otherwise: commits pile up, waited at the end
            else if (false) {}
        }
        if (per_commit < 64) { /* drain: barrier has count 1 per commit; just wait for the last phase */ }
        if (lane == 0) { t1 = clock64(); out[0] = t1 - t0; }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "r"(512u) : "memory");
    }
}

template <int BN>
void run(int n_acc, int per_commit, int mn, int commit_every = 0) {
    long long* d; cudaMalloc(&d, 8);
    const int n = 4096;
    const size_t smem = 1024 + 16384 + BN * 128 + 256;
    cudaFuncSetAttribute(mma_rate_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    // every commit is waited for (per_commit MMAs per wait) when per_commit >= 64; use n_mma == per_commit for one long chain
    mma_rate_kernel<BN><<<1, 128, smem>>>(per_commit >= 64 ? n : per_commit, n_acc, per_commit, mn, d, commit_every);
    cudaError_t e = cudaDeviceSynchronize();
    long long h = 0; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
    const int total = per_commit >= 64 ? n : per_commit;
    if (commit_every) printf("[commit every %d MMAs, one wait at the end] ", commit_every);
    printf("N=%3d acc=%d mmas/commit=%4d %s: %8.1f clk per MMA (%lld clk, %s)\n", BN, n_acc, per_commit, mn ? "MN-major" : "K-major ",
           (double)h / total, h, cudaGetErrorString(e));
    cudaFree(d);
}

int main() {
    run<64>(1, 4096, 0, 4); run<64>(1, 4096, 0, 8); run<64>(1, 4096, 0, 16); run<128>(1, 4096, 0, 4); run<256>(1, 4096, 0, 4);
    for (int mn = 0; mn < 1; ++mn) {
        run<64>(1, 4096, mn); run<64>(2, 4096, mn); run<128>(1, 4096, mn); run<128>(2, 4096, mn); run<256>(1, 4096, mn); run<256>(2, 4096, mn);
        run<64>(1, 64, mn); run<64>(1, 4, mn); run<256>(1, 4, mn); run<32>(1, 4096, mn); run<16>(1, 4096, mn);
    }
    return 0;
}
