// Dispatcher for the dense contraction entry point icd_gemm (include/icd_b200.h).
#include "common.cuh"
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
