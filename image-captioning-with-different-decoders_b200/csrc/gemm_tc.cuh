// Internal interface of the bf16 tensor-core tier (gemm_tc.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

// dst[r*ldd + c] = bf16(src[r*s_r + c*s_c]);  one of s_r / s_c must be 1; ldd multiple of 8.
int icd_convert_bf16(const float* src, int64_t s_r, int64_t s_c, int rows, int cols, void* dst, int64_t ldd,
                     cudaStream_t s);
// C[M,N] = A[M,K] * B[N,K]^T + epilogue; A16/B16 bf16 K-major with leading dimensions lda/ldb (multiples of 8).
int icd_gemm_bf16(const void* A16, int64_t lda, const void* B16, int64_t ldb, float* C, int64_t ldc,
                  int M, int N, int K, const float* bias1, const float* bias2, const float* add1, int64_t ld1,
                  const float* add2, int64_t ld2, const uint8_t* row_mask, float beta, cudaStream_t s,
                  void* C16 = nullptr, int64_t ldc16 = 0);   // optional bf16 copy of the result; C may be NULL then
// General form.  a_mn / b_mn = 1: the operand is MN-major in memory, i.e. stored as [K][M] (resp. [K][N]) with lda
// (ldb) elements between consecutive k rows — the layout of every dW = dY^T X and dX = dY W operand, consumed without
// a transpose.  splitk_ws (optional, icd_gemm_bf16_splitk_floats(M,N,K) floats) enables deterministic split-K.
int icd_gemm_bf16_ex(const void* A16, int64_t lda, int a_mn, const void* B16, int64_t ldb, int b_mn,
                     float* C, int64_t ldc, int M, int N, int K,
                     const float* bias1, const float* bias2, const float* add1, int64_t ld1,
                     const float* add2, int64_t ld2, const uint8_t* row_mask, float beta, cudaStream_t s,
                     void* C16, int64_t ldc16, float* splitk_ws, int64_t splitk_ws_floats,
                     int* deferred_splits = nullptr, const int* m_live = nullptr);
// deferred_splits != NULL: when the plan splits K, the reduce pass is LEFT TO THE CONSUMER — the raw K-slice partials stay in
// splitk_ws as [*deferred_splits][M][N] fp32 planes and C is not written (*deferred_splits = 0: C was written normally,
// epilogue included).  Only for calls whose epilogue is empty (no bias / add / mask / beta / C16).
// m_live != NULL: DEVICE-side row count — only rows < min(M, *m_live) are computed and written (tile loop bounds are
// derived in the kernel, after its dependency wait); the plan then never splits K.
// LSTM weight block [4D][cols] -> bf16 with the rows gate-permuted (row ug*32 + g*8 + j = source row g*D + ug*8 + j)
// several fp32 -> bf16 row conversions in ONE launch (the weight copies at the top of a decoder step: nine launches of a few
// microseconds each otherwise).  Segments that the 128-bit path cannot take are converted by icd_convert_bf16 instead.
constexpr int ICD_CVT_MAX_SEGS = 12;
struct IcdCvtSeg { const float* src; long long s_r; int rows, cols; void* dst; long long ldd; };
int icd_convert_bf16_batch(const IcdCvtSeg* segs, int n, cudaStream_t s);
int icd_convert_bf16_gateperm(const float* src, int64_t s_r, int D, int cols, void* dst, int64_t ldd, cudaStream_t s);
// gates = A W^T (W gate-permuted) + xg + z_hh followed by the LSTMCell, in ONE kernel: writes gates_act / c_new / h_new /
// dropout(h) (+ bf16 copies of h and dropout(h)); the arithmetic equals the contraction + icd_lstm_pointwise_fwd pair bit for bit
int icd_gemm_bf16_lstm_cell(const void* A16, int64_t lda, const void* Wp16, int64_t ldb, int rows, int D, int K,
                            const float* xg, int64_t ld_xg, const float* zhh, int64_t ld_zhh, const float* c_prev,
                            float* gates_act, float* c_new, float* h_new, float* hdrop, int64_t hdrop_stride,
                            const uint8_t* mask, float scale, void* h16, void* hdrop16, cudaStream_t s);
int icd_splitk_finish(const float* splitk_ws, int splits, int M, int N, float* C, int64_t ldc, void* C16, int64_t ldc16,
                      cudaStream_t s);
int64_t icd_gemm_bf16_splitk_floats(int M, int N, int K);
int64_t icd_gemm_tc_ws_bytes(int M, int N, int K);
// fp32-grade tier (ICD_PREC_FP32X3): 3-term bf16 split laid out along K (A pattern which = 0, B pattern which = 1):
// K-major fp32 [rows][cols] (row stride s_r) -> bf16 [rows][6 * up8(cols)]
int icd_split3_bf16(const float* src, int64_t s_r, int rows, int cols, void* dst, int which, cudaStream_t s,
                    const int* m_live = nullptr);   // m_live: optional device-side row count (rows beyond it are skipped)
int64_t icd_gemm_x3_ws_bytes(int M, int N, int K);
// which = 2: THREE stored planes [t1 | t2 | t3] (row of 3 * up8(cols)) instead of the six-segment patterns; the contraction then walks
// them through icd_gemm_x3_planes(segment k-blocks of A, of B): applies to the icd_gemm_bf16_ex calls of this thread until reset with
// (0, 0).  Needs up8(cols) % 64 == 0.  Same MMAs on the same values as the six-segment layout: bit-identical results, half the split bytes.
void icd_gemm_x3_planes(int seg_kb_a, int seg_kb_b);
inline bool icd_x3_three_planes(long long K) {          // K-major operand of K columns: three stored planes possible?
    return ((K + 7) / 8 * 8) % 64 == 0 && getenv("ICD_X3_PLANES6") == nullptr;
}

// Persistent recurrent LSTM kernels (lstm_persistent.cu): the whole h -> gates -> (c, h) recurrence of nn.LSTM, resp. its
// adjoint, as ONE cooperative launch.  icd_lstm_seq_persistent_ok: 1 if the shape is covered (else keep a launch chain).
// h16: ((L+1)*B, H) bf16 with row block 0 zeroed (h_0), c_all / h_all row block 0 zeroed by the caller; bar: 4 bytes.
int icd_lstm_seq_persistent_ok(int B, int L, int H);
int icd_lstm_seq_fwd_persistent(int B, int L, int H, const float* w_hh, const float* xg, float* gates_act, float* c_all,
                                float* h_all, float* hout, void* h16, void* hout16, unsigned int* bar, cudaStream_t s);
int icd_lstm_seq_bwd_persistent(int B, int L, int H, const float* w_hh, const float* d_hout, const float* gates_act,
                                const float* c_all, float* dc, float* dg, void* dg16, unsigned int* bar, cudaStream_t s);
