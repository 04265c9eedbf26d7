// bf16 tensor-core tier of the dense contraction:  C[M,N] = A[M,K] * B[N,K]^T + epilogue   (fp32 accumulate/out)
//
// Blackwell-native (sm_100a) kernel, hand-written PTX:
//   * operands: bf16, K-major, staged by TMA (cp.async.bulk.tensor.2d, 128B swizzle) into a 4-stage shared-memory ring;
//   * math: tcgen05.mma.cta_group::1.kind::f16, UMMA 128 x BN x 16 (BN = 256 or 128), issued by ONE elected thread,
//     accumulators in TMEM (2 stages x BN fp32 columns, so the epilogue of tile i overlaps the MMAs of tile i+1);
//   * sync: mbarrier full/empty ring (TMA <-> MMA), tcgen05.commit -> mbarrier (MMA -> TMA slot release and
//     MMA -> epilogue), tmem_empty barrier (epilogue -> MMA);
//   * warp roles (192 threads): warp 0 TMA producer, warp 1 TMEM allocator + MMA issuer, warps 2-5 epilogue
//     (tcgen05.ld 32x32b.x32 -> registers -> padded smem transpose -> coalesced fp32 stores with the fused epilogue:
//     bias1 + bias2 + add1 + add2, row mask, beta);
//   * persistent: grid = min(#tiles, #SMs), static round-robin tile schedule.
// fp32 sources are converted to K-major bf16 (optionally transposing) by convert_bf16_kernel into caller-provided
// workspace, so the kernel only ever sees K-major operands (the weight-gradient contractions dW = dY^T X are fed
// transposed copies, which the conversion pass produces at no extra traffic).
#include "common.cuh"
#include "gemm_tc.cuh"
#include <cuda.h>
#include <cuda_bf16.h>
#include <mutex>

namespace {

constexpr int BM = 128, BK = 64, STAGES = 4, UMMA_K = 16;
constexpr int A_BYTES = BM * BK * 2;                 // 16 KB
constexpr int EPI_LD = 33;                           // padded row of the per-warp 32x32 transpose buffer
constexpr int NUM_THREADS = 192;

struct EpiArgs {
    float* C; long long ldc;
    int M, N, K;
    const float* bias1; const float* bias2;
    const float* add1; long long ld1;
    const float* add2; long long ld2;
    const unsigned char* row_mask;
    float beta;
    __nv_bfloat16* C16; long long ldc16;      // optional bf16 copy of the result (C may then be NULL)
};

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
        if (++spins > (1ull << 22)) __trap();        // a protocol bug must fault, never hang the GPU
    }
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* tm, int c0, int c1, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 :: "r"(dst), "l"(tm), "r"(bar), "r"(c0), "r"(c1) : "memory");
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

// K-major, 128B-swizzled shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout):
//   [0,14) start address >> 4 | [16,30) leading byte offset >> 4 (ignored for swizzled K-major; 1)
//   [32,46) stride byte offset >> 4 = 1024 B between 8-row groups | [46,48) version = 1 | [61,64) layout = SWIZZLE_128B (2)
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t addr) {
    return (uint64_t)((addr & 0x3FFFF) >> 4) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}

template <int BN>
struct Cfg {
    static constexpr int B_BYTES = BN * BK * 2;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int EPI_BYTES = 4 * 32 * EPI_LD * 4;
    static constexpr int BAR_BYTES = 256;
    static constexpr int SMEM = 1024 /*align slack*/ + STAGES * STAGE_BYTES + EPI_BYTES + BAR_BYTES;
    static constexpr int TMEM_COLS = 2 * BN;         // two accumulator stages (power of two: 256 or 512)
    // cute::UMMA::InstrDescriptor: c_format F32 (1<<4), a/b format BF16 (1<<7, 1<<10), K-major both,
    // n_dim = N>>3 at [17,23), m_dim = M>>4 at [24,29)
    static constexpr uint32_t IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
};

template <int BN>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const EpiArgs e) {
    using C_ = Cfg<BN>;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;       // SWIZZLE_128B tiles need 1024 B alignment
    uint8_t* gbase = smem_raw + (base - smem_u32(smem_raw));
    const uint32_t sA = base, sB = base + STAGES * A_BYTES;
    float* sEpi = reinterpret_cast<float*>(gbase + STAGES * C_::STAGE_BYTES);
    const uint32_t bars = base + STAGES * C_::STAGE_BYTES + C_::EPI_BYTES;
    const uint32_t full0 = bars, empty0 = bars + 8 * STAGES, tfull0 = bars + 16 * STAGES, tempty0 = tfull0 + 16;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(gbase + STAGES * C_::STAGE_BYTES + C_::EPI_BYTES + 16 * STAGES + 32);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tiles_m = (e.M + BM - 1) / BM, tiles_n = (e.N + BN - 1) / BN;
    const int num_tiles = tiles_m * tiles_n;
    const int nkb = (e.K + BK - 1) / BK;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" :: "l"(&tmA) : "memory");
        asm volatile("prefetch.tensormap [%0];" :: "l"(&tmB) : "memory");
        for (int i = 0; i < STAGES; ++i) { mbar_init(full0 + 8 * i, 1); mbar_init(empty0 + 8 * i, 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(tfull0 + 8 * i, 1); mbar_init(tempty0 + 8 * i, 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     :: "r"(smem_u32(tmem_slot)), "r"((uint32_t)C_::TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // =================================== TMA producer ===================================
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                const int m0 = (tile % tiles_m) * BM, n0 = (tile / tiles_m) * BN;
                for (int kb = 0; kb < nkb; ++kb) {
                    mbar_wait(empty0 + 8 * stage, phase ^ 1);
                    mbar_arrive_expect_tx(full0 + 8 * stage, C_::STAGE_BYTES);
                    tma_load_2d(sA + stage * A_BYTES, &tmA, kb * BK, m0, full0 + 8 * stage);
                    tma_load_2d(sB + stage * C_::B_BYTES, &tmB, kb * BK, n0, full0 + 8 * stage);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // =================================== MMA issuer ===================================
        int stage = 0; uint32_t phase = 0;
        int acc = 0; uint32_t acc_phase = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            mbar_wait(tempty0 + 8 * acc, acc_phase ^ 1);              // epilogue has drained this accumulator
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
            for (int kb = 0; kb < nkb; ++kb) {
                mbar_wait(full0 + 8 * stage, phase);                  // TMA bytes have landed
                tc_fence_after();
                if (lane == 0) {
                    const uint64_t adesc = make_smem_desc(sA + stage * A_BYTES);
                    const uint64_t bdesc = make_smem_desc(sB + stage * C_::B_BYTES);
#pragma unroll
                    for (int k = 0; k < BK / UMMA_K; ++k)              // +32 B per K slice inside the 128 B swizzle span
                        tc_mma_f16(d_tmem, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), C_::IDESC,
                                   (kb > 0 || k > 0) ? 1u : 0u);
                    tc_commit(empty0 + 8 * stage);                    // smem slot reusable once these MMAs retire
                    if (kb == nkb - 1) tc_commit(tfull0 + 8 * acc);   // accumulator complete -> epilogue
                }
                __syncwarp();
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
    } else {
        // =================================== epilogue warps 2..5 ===================================
        const int q = warp & 3;                                       // TMEM lane quarter this warp may access
        float* sE = sEpi + (warp - 2) * 32 * EPI_LD;
        int acc = 0; uint32_t acc_phase = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            const int m0 = (tile % tiles_m) * BM, n0 = (tile / tiles_m) * BN;
            mbar_wait(tfull0 + 8 * acc, acc_phase);
            tc_fence_after();
            const int mrow0 = m0 + q * 32;
#pragma unroll 1
            for (int c = 0; c < BN / 32; ++c) {
                const int nb = n0 + c * 32;
                if (nb >= e.N) break;                                 // warp-uniform
                uint32_t v[32];
                tc_ld_32x32b_x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN + c * 32), v);
                tc_wait_ld();
#pragma unroll
                for (int j = 0; j < 32; ++j) sE[lane * EPI_LD + j] = __uint_as_float(v[j]);
                __syncwarp();
                const int n = nb + lane;
                const bool nok = n < e.N;
                float bsum = 0.f;
                if (nok) { if (e.bias1) bsum += e.bias1[n]; if (e.bias2) bsum += e.bias2[n]; }
#pragma unroll 4
                for (int r = 0; r < 32; ++r) {
                    const int m = mrow0 + r;
                    if (m >= e.M) break;                              // warp-uniform
                    if (nok) {
                        float x = sE[r * EPI_LD + lane] + bsum;
                        if (e.add1) x += e.add1[(long long)m * e.ld1 + n];
                        if (e.add2) x += e.add2[(long long)m * e.ld2 + n];
                        if (e.row_mask && !e.row_mask[m]) x = 0.f;
                        if (e.C) {
                            float* cp = e.C + (long long)m * e.ldc + n;
                            if (e.beta != 0.f) x += e.beta * (*cp);
                            *cp = x;
                        }
                        if (e.C16) e.C16[(long long)m * e.ldc16 + n] = __float2bfloat16_rn(x);
                    }
                }
                __syncwarp();
            }
            tc_fence_before();
            if (lane == 0) mbar_arrive(tempty0 + 8 * acc);
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;"
                     :: "r"(tmem_base), "r"((uint32_t)C_::TMEM_COLS) : "memory");
    }
}

// ---------------------------------------------------------------------------------------------- fp32 -> bf16 staging
// dst[r*ldd + c] = bf16(src[r*s_r + c*s_c]), r < rows, c < cols.  One of s_r / s_c is 1.
__global__ void convert_rows_kernel(const float* __restrict__ src, long long s_r, int rows, int cols,
                                    __nv_bfloat16* __restrict__ dst, long long ldd) {
    const long long n2 = ((long long)cols + 1) / 2;
    const long long total = (long long)rows * n2;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / n2; const int c = (int)(i % n2) * 2;
        const float* s = src + r * s_r + c;
        const float x = s[0], y = (c + 1 < cols) ? s[1] : 0.f;
        if (c + 1 < cols || (ldd & 1) == 0)          // ldd is even (multiple of 8): the pair store stays inside the row
            *reinterpret_cast<__nv_bfloat162*>(dst + r * ldd + c) = __floats2bfloat162_rn(x, y);
        else dst[r * ldd + c] = __float2bfloat16_rn(x);
    }
}

// transposing variant: src is contiguous along r (s_r == 1, s_c = source leading dimension).  32x32 smem tiles.
__global__ void convert_transpose_kernel(const float* __restrict__ src, long long s_c, int rows, int cols,
                                         __nv_bfloat16* __restrict__ dst, long long ldd) {
    __shared__ float t[32][33];
    const int r0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {          // read: lanes along r (contiguous in src)
        const int c = c0 + j, r = r0 + threadIdx.x;
        t[j][threadIdx.x] = (c < cols && r < rows) ? src[(long long)c * s_c + r] : 0.f;
    }
    __syncthreads();
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {          // write: lanes along c (contiguous in dst)
        const int r = r0 + j, c = c0 + threadIdx.x;
        if (r < rows && c < cols) dst[(long long)r * ldd + c] = __float2bfloat16_rn(t[threadIdx.x][j]);
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn g_encode = nullptr;
std::once_flag g_encode_once;

int get_encode(EncodeTiledFn* out) {
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

int make_tmap(CUtensorMap* tm, const __nv_bfloat16* p, long long ld, int rows, int K, int box_rows) {
    EncodeTiledFn enc;
    ICD_TRY(get_encode(&enc));
    ICD_CHECK_ARG((reinterpret_cast<uintptr_t>(p) & 15) == 0 && (ld % 8) == 0,
                  "gemm_tc: bf16 operand must be 16-byte aligned with a leading dimension multiple of 8 (ld=%lld)", ld);
    cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
    cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<__nv_bfloat16*>(p), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { icd_set_error("gemm_tc: cuTensorMapEncodeTiled failed (%d)", (int)r); return -3; }
    return 0;
}

template <int BN>
int launch(const CUtensorMap& tmA, const CUtensorMap& tmB, const EpiArgs& e, cudaStream_t s) {
    static bool attr_set = false;
    if (!attr_set) {
        ICD_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<BN>::SMEM));
        attr_set = true;
    }
    const int tiles = ((e.M + BM - 1) / BM) * ((e.N + BN - 1) / BN);
    const int grid = tiles < ICD_NUM_SMS ? tiles : ICD_NUM_SMS;
    gemm_tc_kernel<BN><<<grid, NUM_THREADS, Cfg<BN>::SMEM, s>>>(tmA, tmB, e);
    ICD_LAUNCH_CHECK();
    return 0;
}

}  // namespace

extern "C" int icd_has_tensor_core_gemm(void) { return 1; }

int icd_convert_bf16(const float* src, int64_t s_r, int64_t s_c, int rows, int cols, void* dst, int64_t ldd,
                     cudaStream_t s) {
    if (rows == 0 || cols == 0) return 0;
    ICD_CHECK_ARG(s_r == 1 || s_c == 1, "convert_bf16: source needs a unit stride");
    ICD_CHECK_ARG(ldd % 8 == 0 && ldd >= cols, "convert_bf16: ldd=%lld must be a multiple of 8 and >= cols", (long long)ldd);
    __nv_bfloat16* d = reinterpret_cast<__nv_bfloat16*>(dst);
    if (s_c == 1) {
        const long long total = (long long)rows * ((cols + 1) / 2);
        long long blocks = (total + 255) / 256;
        if (blocks > ICD_NUM_SMS * 32) blocks = ICD_NUM_SMS * 32;
        convert_rows_kernel<<<(unsigned)blocks, 256, 0, s>>>(src, s_r, rows, cols, d, ldd);
    } else {
        dim3 grid((rows + 31) / 32, (cols + 31) / 32);
        ICD_CHECK_ARG(grid.y <= 65535, "convert_bf16: too many columns for the transposing path");
        convert_transpose_kernel<<<grid, dim3(32, 8), 0, s>>>(src, s_c, rows, cols, d, ldd);
    }
    ICD_LAUNCH_CHECK();
    return 0;
}

int icd_gemm_bf16(const void* A16, int64_t lda, const void* B16, int64_t ldb, float* C, int64_t ldc,
                  int M, int N, int K, const float* bias1, const float* bias2, const float* add1, int64_t ld1,
                  const float* add2, int64_t ld2, const uint8_t* row_mask, float beta, cudaStream_t s,
                  void* C16, int64_t ldc16) {
    if (M == 0 || N == 0) return 0;
    ICD_CHECK_ARG(K > 0, "gemm_tc: K must be positive");
    ICD_CHECK_ARG(C != nullptr || C16 != nullptr, "gemm_tc: no output");
    const int tiles256 = ((M + BM - 1) / BM) * ((N + 255) / 256);
    const bool use256 = (N > 128) && tiles256 >= ICD_NUM_SMS;
    const int BN = use256 ? 256 : 128;
    CUtensorMap tmA, tmB;
    ICD_TRY(make_tmap(&tmA, reinterpret_cast<const __nv_bfloat16*>(A16), lda, M, K, BM));
    ICD_TRY(make_tmap(&tmB, reinterpret_cast<const __nv_bfloat16*>(B16), ldb, N, K, BN));
    EpiArgs e;
    e.C = C; e.ldc = ldc; e.M = M; e.N = N; e.K = K; e.bias1 = bias1; e.bias2 = bias2;
    e.add1 = add1; e.ld1 = ld1; e.add2 = add2; e.ld2 = ld2; e.row_mask = row_mask; e.beta = beta;
    e.C16 = reinterpret_cast<__nv_bfloat16*>(C16); e.ldc16 = ldc16;
    return use256 ? launch<256>(tmA, tmB, e, s) : launch<128>(tmA, tmB, e, s);
}

int64_t icd_gemm_tc_ws_bytes(int M, int N, int K) {
    const int64_t ldk = ((int64_t)K + 7) / 8 * 8;
    return ((int64_t)M * ldk * 2 + 255) / 256 * 256 + ((int64_t)N * ldk * 2 + 255) / 256 * 256;
}

int icd_gemm_tc_launch(const icd_gemm_desc_t* d, cudaStream_t s) {
    if (d->M == 0 || d->N == 0) return 0;
    ICD_CHECK_ARG(d->sak == 1 || d->sam == 1, "gemm: A needs a unit stride");
    ICD_CHECK_ARG(d->sbk == 1 || d->sbn == 1, "gemm: B needs a unit stride");
    const int64_t need = icd_gemm_tc_ws_bytes(d->M, d->N, d->K);
    ICD_CHECK_ARG(d->ws && d->ws_bytes >= need, "gemm: ICD_PREC_BF16 needs %lld bytes of workspace (icd_gemm_ws_bytes), got %lld",
                  (long long)need, (long long)d->ws_bytes);
    const int64_t ldk = ((int64_t)d->K + 7) / 8 * 8;
    char* a16 = reinterpret_cast<char*>(d->ws);
    char* b16 = a16 + ((int64_t)d->M * ldk * 2 + 255) / 256 * 256;
    ICD_TRY(icd_convert_bf16(d->A, d->sam, d->sak, d->M, d->K, a16, ldk, s));
    ICD_TRY(icd_convert_bf16(d->B, d->sbn, d->sbk, d->N, d->K, b16, ldk, s));
    return icd_gemm_bf16(a16, ldk, b16, ldk, d->C, d->ldc, d->M, d->N, d->K, d->bias1, d->bias2, d->add1, d->ld1,
                         d->add2, d->ld2, d->row_mask, d->beta, s, nullptr, 0);
}
