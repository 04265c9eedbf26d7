// tcgen05 / TMA / mbarrier PTX wrappers and the UMMA shared-memory descriptor shared by the tensor-core kernels of this
// library (gemm_tc.cu: the general contraction; lstm_persistent.cu: the persistent recurrent LSTM kernels).  sm_100a only.
#pragma once
#include "common.cuh"
#include <cuda.h>
#include <cuda_bf16.h>
#include <mutex>

namespace {

constexpr int TC_BK = 64;                            // k-block: 64 bf16 = one 128-byte swizzle span
constexpr int TC_MN_BLOCK_BYTES = TC_BK * 128;       // one 64-element MN block of an MN-major tile: BK rows x 128 B

// ---------------------------------------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok = 0;
    unsigned long long spins = 0;
    while (true) {
        asm volatile("{\n\t.reg .pred p;\n\t"
                     "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                     "selp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
        if (ok) break;
        if (++spins > (1ull << 24)) __trap();        // a protocol bug must fault, never hang the GPU
    }
}
// the same for a waiter that is not on the critical path (a TMA producer several ring slots ahead of the consumer): sleeps
// between polls so that its spin does not take issue slots from the epilogue warps of the same scheduler
__device__ __forceinline__ void mbar_wait_backoff(uint32_t bar, uint32_t parity) {
    uint32_t ok = 0;
    unsigned long long spins = 0;
    while (true) {
        asm volatile("{\n\t.reg .pred p;\n\t"
                     "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                     "selp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
        if (ok) break;
        __nanosleep(64);
        if (++spins > (1ull << 22)) __trap();
    }
}
// non-blocking probe of a phase (used to start the barrier read of the NEXT ring slot before the MMAs of the current one
// are issued: the ~150-clock latency of the probe then hides behind the issue of four UTCHMMAs)
__device__ __forceinline__ uint32_t mbar_test(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\t"
                 "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                 "selp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok;
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* tm, int c0, int c1, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 :: "r"(dst), "l"(tm), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_load_2d_mc(uint32_t dst, const CUtensorMap* tm, int c0, int c1, uint32_t bar, uint16_t mask) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;"
                 :: "r"(dst), "l"(tm), "r"(bar), "r"(c0), "r"(c1), "h"(mask) : "memory");
}
__device__ __forceinline__ void tc_commit_mc(uint32_t bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 :: "r"(bar), "h"(mask) : "memory");
}
// One lane of a fully converged warp (CUTLASS' elect_one_sync).  The single-thread roles (TMA producer, MMA issuer) run
// their loops with the WHOLE warp in uniform control flow and issue under this predicate: code under `if (lane == 0)` is
// divergent for the compiler, which then wraps every UTMALDG / UTCHMMA / UTCBAR in an ELECT + R2UR.BROADCAST "waterfall"
// loop (~100 clocks per instruction: measured 750 clocks per k-block of four MMAs, i.e. issue-bound at 15 % tensor use).
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ uint32_t uniform_u32(uint32_t x) { return __shfl_sync(0xffffffffu, x, 0); }   // provably warp-uniform copy
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 :: "r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void tc_ld_32x32b_x32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                 "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
                 "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                   "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                   "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                 : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// 128B-swizzled shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout):
//   [0,14) start address >> 4 | [16,30) leading byte offset >> 4 | [32,46) stride byte offset >> 4
//   [46,48) version = 1 | [61,64) layout = SWIZZLE_128B (2)
// K-major tile  (rows of 64 bf16 = 128 B, 8-row swizzle atoms of 1024 B): LBO unused (1), SBO = 1024 B between
//                8-row groups; the next UMMA_K = 16 slice starts 32 B further inside the 128 B span.
// MN-major tile (BK rows of 64 MN-elements = 128 B each, one row per k; 64-wide MN blocks TC_MN_BLOCK_BYTES apart):
//                canonical ((8,n),(8,k)) : ((1,LBO),(8,SBO)) in 16-byte units: LBO = MN-block stride, SBO = 1024 B
//                between groups of 8 k-rows; the next UMMA_K = 16 slice starts 2 k-groups = 2048 B further.
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t addr, int mn_major) {
    const uint64_t lbo = mn_major ? (uint64_t)(TC_MN_BLOCK_BYTES >> 4) : 1ull;
    return (uint64_t)((addr & 0x3FFFF) >> 4) | (lbo << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
inline int get_encode(EncodeTiledFn* out) {
    static EncodeTiledFn g_encode = nullptr;
    static std::once_flag g_encode_once;
    std::call_once(g_encode_once, [] {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            g_encode = reinterpret_cast<EncodeTiledFn>(fn);
    });
    if (!g_encode) { icd_set_error("gemm_tc: cuTensorMapEncodeTiled entry point not available"); return -3; }
    *out = g_encode;
    return 0;
}

// K-major operand:  memory [rows][K] (ld elements between rows)  -> dims {K, rows}, box {BK, box_rows}
// MN-major operand: memory [K][rows] (ld elements between k rows) -> dims {rows, K}, box {64, BK}
// box_k: k-extent of the box (a cluster CTA that fetches only a slice of a tile: box_rows rows of a K-major tile, box_k
// k-rows of an MN-major one)
inline int make_tmap(CUtensorMap* tm, const __nv_bfloat16* p, long long ld, int rows, int K, int box_rows, int mn_major,
                     int box_k = TC_BK) {
    EncodeTiledFn enc;
    ICD_TRY(get_encode(&enc));
    ICD_CHECK_ARG((reinterpret_cast<uintptr_t>(p) & 15) == 0 && (ld % 8) == 0,
                  "gemm_tc: bf16 operand must be 16-byte aligned with a leading dimension multiple of 8 (ld=%lld)", ld);
    cuuint64_t dims[2], strides[1] = {(cuuint64_t)ld * 2};
    cuuint32_t box[2], estr[2] = {1, 1};
    if (!mn_major) { dims[0] = (cuuint64_t)K; dims[1] = (cuuint64_t)rows; box[0] = TC_BK; box[1] = (cuuint32_t)box_rows; }
    else           { dims[0] = (cuuint64_t)rows; dims[1] = (cuuint64_t)K; box[0] = 64; box[1] = (cuuint32_t)box_k; }
    CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<__nv_bfloat16*>(p), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { icd_set_error("gemm_tc: cuTensorMapEncodeTiled failed (%d)", (int)r); return -3; }
    return 0;
}

}  // namespace
