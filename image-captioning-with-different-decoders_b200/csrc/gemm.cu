// Dispatcher for the dense contraction entry point icd_gemm (include/icd_b200.h).
#include "common.cuh"
#include <stdlib.h>
#include "gemm_tc.cuh"

int icd_gemm_f32_launch(const icd_gemm_desc_t* d, cudaStream_t s);   // gemm_f32.cu
int icd_gemm_tc_launch(const icd_gemm_desc_t* d, cudaStream_t s);    // gemm_tc.cu (tcgen05 / TMA, bf16)
int icd_gemm_x3_launch(const icd_gemm_desc_t* d, cudaStream_t s);    // gemm_tc.cu (3-term bf16 split, fp32-grade)

extern "C" int64_t icd_gemm_ws_bytes(int32_t M, int32_t N, int32_t K, int32_t precision) {
    if (precision == ICD_PREC_FP32X3) return icd_gemm_x3_ws_bytes(M, N, K);
    return precision == ICD_PREC_BF16 ? icd_gemm_tc_ws_bytes(M, N, K) : 0;
}

extern "C" int icd_gemm(const icd_gemm_desc_t* d, void* stream) {
    ICD_CHECK_ARG(d != nullptr, "gemm: null descriptor");
    cudaStream_t s = icd_stream(stream);
    if (d->precision == ICD_PREC_FP32) return icd_gemm_f32_launch(d, s);
    if (d->precision == ICD_PREC_BF16) return icd_gemm_tc_launch(d, s);
    if (d->precision == ICD_PREC_FP32X3) return icd_gemm_x3_launch(d, s);
    icd_set_error("gemm: unknown precision %d", d->precision);
    return -1;
}

// workspace the internal helper hands to the tensor-core tiers (set by the decoder entry points from desc->tc_ws for the
// duration of one call; thread-local, so concurrent callers on different host threads do not interfere)
static thread_local void* g_simple_ws = nullptr;
static thread_local int64_t g_simple_ws_bytes = 0;
void icd_gemm_simple_set_ws(void* ws, int64_t bytes) { g_simple_ws = ws; g_simple_ws_bytes = bytes; }

// ---- stationary-operand split cache of the fp32-grade tier (see common.cuh)
namespace {
struct X3Entry { const float* p; int64_t stride; int mn, K, mn_major, seg, which; void* dst; };
struct X3Cache {
    char* base = nullptr; int64_t cap = 0, used = 0;
    const float* marked[8]; int n_marked = 0;
    X3Entry e[16]; int n = 0;
};
thread_local X3Cache g_x3;
}  // namespace
void icd_x3_cache_begin(void* mem, int64_t bytes) {
    g_x3 = X3Cache();
    const char* off = getenv("ICD_X3_CACHE");                 // ICD_X3_CACHE=0: every contraction re-splits both operands (test / A-B hook)
    if (off && atoi(off) == 0) return;
    const uintptr_t a = (reinterpret_cast<uintptr_t>(mem) + 255) & ~uintptr_t(255);
    g_x3.base = reinterpret_cast<char*>(a);
    g_x3.cap = bytes - (int64_t)(a - reinterpret_cast<uintptr_t>(mem));
}
void icd_x3_cache_end() { g_x3 = X3Cache(); }
void icd_x3_cache_mark(const float* base) { if (g_x3.base && g_x3.n_marked < 8) g_x3.marked[g_x3.n_marked++] = base; }
void* icd_x3_cache_lookup(const float* p, int64_t stride, int mn, int K, int mn_major, int seg, int which, int64_t bytes, bool* fresh) {
    *fresh = true;
    if (!g_x3.base) return nullptr;
    bool ok = false;
    for (int i = 0; i < g_x3.n_marked; ++i) ok |= (g_x3.marked[i] == p);
    if (!ok) return nullptr;
    for (int i = 0; i < g_x3.n; ++i) {
        const X3Entry& x = g_x3.e[i];
        if (x.p == p && x.stride == stride && x.mn == mn && x.K == K && x.mn_major == mn_major && x.seg == seg && x.which == which) {
            *fresh = false;
            return x.dst;
        }
    }
    if (g_x3.n >= 16 || g_x3.used + bytes > g_x3.cap) return nullptr;
    void* dst = g_x3.base + g_x3.used;
    g_x3.used += bytes;
    g_x3.e[g_x3.n++] = X3Entry{p, stride, mn, K, mn_major, seg, which, dst};
    return dst;
}

int icd_gemm_simple(int prec, const float* A, int64_t sam, int64_t sak, const float* B, int64_t sbn, int64_t sbk,
                    float* C, int64_t ldc, int M, int N, int K, const float* bias1, const float* bias2,
                    const float* add1, int64_t ld1, const float* add2, int64_t ld2, const uint8_t* row_mask,
                    float beta, cudaStream_t s, int flags) {
    icd_gemm_desc_t d;
    d.flags = flags; d.ws = g_simple_ws; d.ws_bytes = g_simple_ws_bytes;
    d.A = A; d.sam = sam; d.sak = sak; d.B = B; d.sbn = sbn; d.sbk = sbk; d.C = C; d.ldc = ldc;
    d.M = M; d.N = N; d.K = K; d.bias1 = bias1; d.bias2 = bias2; d.add1 = add1; d.ld1 = ld1;
    d.add2 = add2; d.ld2 = ld2; d.row_mask = row_mask; d.beta = beta; d.precision = prec;
    return icd_gemm(&d, (void*)s);
}

extern "C" int64_t icd_gemm_bf16_splitk_ws_floats(int32_t M, int32_t N, int32_t K) {
    return icd_gemm_bf16_splitk_floats(M, N, K);
}

extern "C" int icd_gemm_bf16_operands(const void* A16, int64_t lda, int32_t a_mn, const void* B16, int64_t ldb, int32_t b_mn,
                                      float* C, int64_t ldc, void* C16, int64_t ldc16, int32_t M, int32_t N, int32_t K,
                                      const float* bias1, const float* add1, int64_t ld1, const uint8_t* row_mask,
                                      float* splitk_ws, int64_t splitk_ws_floats, void* stream) {
    ICD_CHECK_ARG(A16 && B16, "gemm_bf16_operands: null operand");
    ICD_CHECK_ARG(M >= 0 && N >= 0 && K > 0, "gemm_bf16_operands: bad dims");
    return icd_gemm_bf16_ex(A16, lda, a_mn, B16, ldb, b_mn, C, ldc, M, N, K, bias1, nullptr, add1, ld1, nullptr, 0, row_mask,
                            0.f, icd_stream(stream), C16, ldc16, splitk_ws, splitk_ws_floats);
}
