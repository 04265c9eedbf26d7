// BaselineDecoder.forward / backward (models/baseline.py:81-111): embedding of captions[:, :-1] (:93,97), image
// feature prepended as LSTM step 0 (:101), 1-layer nn.LSTM from a zero state (:106), vocabulary linear (:109).
// The input contraction x * W_ih^T + b_ih + b_hh is hoisted over all L steps, the recurrence runs as
// h * W_hh^T (+ fused gate pointwise) per step, and the vocabulary contraction runs once over all (b, t).
#include "common.cuh"
#include "gemm_tc.cuh"
#include <algorithm>

namespace {
int check(const icd_base_desc_t* d) {
    ICD_CHECK_ARG(d != nullptr, "baseline_decoder: null descriptor");
    ICD_CHECK_ARG(d->B > 0 && d->L > 0 && d->L <= ICD_MAX_STEPS, "baseline_decoder: B=%d L=%d out of range", d->B, d->L);
    ICD_CHECK_ARG(d->E % 4 == 0 && d->H % 4 == 0, "baseline_decoder: E and H must be multiples of 4");
    return 0;
}
}  // namespace

// ---------------------------------------------------------------------------------------------------------------
// bf16 tensor-core tier: same dataflow, every contraction on the tcgen05 kernel (bf16 operands, fp32 accumulate), LSTM gate
// math / saved activations / gradients fp32.  Row-major activations and weights are consumed as MN-major operands by the
// dW = dY^T X and dX = dY W contractions (no transposed copies).
// ---------------------------------------------------------------------------------------------------------------
namespace {
inline int64_t up8(int64_t x) { return (x + 7) / 8 * 8; }
struct Arena16 {
    char* base; int64_t cap, off; bool ok;
    void* take_bytes(int64_t bytes) {
        bytes = (bytes + 255) / 256 * 256;
        void* p = base ? base + off : nullptr;
        off += bytes;
        if (base && off > cap) ok = false;
        return p;
    }
    void* take(int64_t rows, int64_t ld) { return take_bytes(rows * ld * 2); }
};
struct BaseBufs {
    void *Wih, *Whh, *Wlin, *x, *h, *hout, *dY, *dg;
    float *splitk, *colsum_ws; int64_t splitk_floats, ldE, ldV;
    unsigned int* bar;                                   // grid-barrier counter of the persistent recurrent kernels
};
void carve_base(const icd_base_desc_t* d, Arena16& a, BaseBufs& b) {
    const int64_t B = d->B, L = d->L, E = d->E, H = d->H, V = d->V, LB = L * B;
    b.ldE = up8(E); b.ldV = up8(V);
    b.Wih = a.take(4 * H, b.ldE); b.Whh = a.take(4 * H, H); b.Wlin = a.take(V, H);
    b.x = a.take(LB, b.ldE); b.h = a.take(LB + B, H); b.hout = a.take(LB, H);
    b.dY = a.take(LB, b.ldV); b.dg = a.take(LB, 4 * H);
    const int shapes[][3] = {{(int)LB, (int)(4 * H), (int)E}, {(int)B, (int)(4 * H), (int)H}, {(int)LB, (int)V, (int)H},
                             {(int)LB, (int)H, (int)V}, {(int)V, (int)H, (int)LB}, {(int)B, (int)H, (int)(4 * H)},
                             {(int)(4 * H), (int)H, (int)LB}, {(int)(4 * H), (int)E, (int)LB}, {(int)LB, (int)E, (int)(4 * H)}};
    int64_t f = 0;
    for (const auto& sh : shapes) { const int64_t n = icd_gemm_bf16_splitk_floats(sh[0], sh[1], sh[2]); if (n > f) f = n; }
    b.splitk_floats = f;
    b.splitk = reinterpret_cast<float*>(a.take_bytes(f * 4));
    b.colsum_ws = reinterpret_cast<float*>(a.take_bytes(icd_colsum_bf16_ws_floats(LB, (int)V) * 4));
    b.bar = reinterpret_cast<unsigned int*>(a.take_bytes(256));
}
int check_base16(const icd_base_desc_t* d, Arena16& a, BaseBufs& b) {
    ICD_CHECK_ARG(d->H % 8 == 0 && d->E % 4 == 0, "baseline_decoder(bf16): H must be a multiple of 8 and E of 4");
    a.base = reinterpret_cast<char*>(d->tc_ws); a.cap = d->tc_ws_bytes; a.off = 0; a.ok = true;
    carve_base(d, a, b);
    ICD_CHECK_ARG(d->tc_ws != nullptr && a.ok, "baseline_decoder(bf16): tc_ws too small (%lld given, %lld needed)",
                  (long long)d->tc_ws_bytes, (long long)a.off);
    return 0;
}
inline char* at16(void* p, int64_t elem_off) { return reinterpret_cast<char*>(p) + elem_off * 2; }
#define BCVT(src, sr, rows, cols, dst, ld) ICD_TRY(icd_convert_bf16((src), (sr), 1, (rows), (cols), (dst), (ld), s))
#define BMM(A16, lda, amn, B16, ldb, bmn, Cp, ldc, M, N, K, b1, b2, a1, l1, C16, ldc16) \
    ICD_TRY(icd_gemm_bf16_ex((A16), (lda), (amn), (B16), (ldb), (bmn), (Cp), (ldc), (M), (N), (K), (b1), (b2), (a1), (l1), \
                             nullptr, 0, nullptr, 0.f, s, (C16), (ldc16), u.splitk, u.splitk_floats))

int baseline_fwd_bf16(const icd_base_desc_t* d, cudaStream_t s) {
    Arena16 ar; BaseBufs u;
    ICD_TRY(check_base16(d, ar, u));
    const int B = d->B, L = d->L, E = d->E, H = d->H, V = d->V;
    const size_t BH = (size_t)B * H;
    const int LB = L * B;
    ICD_CUDA(cudaMemcpyAsync(d->x, d->img_features, sizeof(float) * (size_t)B * E, cudaMemcpyDeviceToDevice, s));      // :101
    if (L > 1) ICD_TRY(icd_embed_gather(d->emb_w, d->emb_is_f64, d->captions, B, L, L - 1, E, V, d->x + (size_t)B * E, s));
    // K8: the recurrence as ONE persistent kernel (W_hh slices resident in shared memory, grid barrier per step, gate math
    // in the tcgen05 epilogue; lstm_persistent.cu) whenever the shape allows it, else the per-step launch chain below
    const bool persistent = icd_lstm_seq_persistent_ok(B, L, H) != 0;
    BCVT(d->w_ih, E, 4 * H, E, u.Wih, u.ldE);
    if (!persistent) BCVT(d->w_hh, H, 4 * H, H, u.Whh, H);
    BCVT(d->lin_w, H, V, H, u.Wlin, H);
    BCVT(d->x, E, LB, E, u.x, u.ldE);
    BMM(u.x, u.ldE, 0, u.Wih, u.ldE, 0, d->xg, 4 * H, LB, 4 * H, E, d->b_ih, d->b_hh, nullptr, 0, nullptr, 0);
    ICD_CUDA(cudaMemsetAsync(d->h_all, 0, sizeof(float) * BH, s));                           // zero (h0, c0) (:106)
    ICD_CUDA(cudaMemsetAsync(d->c_all, 0, sizeof(float) * BH, s));
    ICD_CUDA(cudaMemsetAsync(u.h, 0, BH * 2, s));
    if (persistent)
        ICD_TRY(icd_lstm_seq_fwd_persistent(B, L, H, d->w_hh, d->xg, d->gates_act, d->c_all, d->h_all, d->hout, u.h, u.hout,
                                            u.bar, s));
    for (int t = 0; t < L && !persistent; ++t) {
        BMM(at16(u.h, (int64_t)t * BH), H, 0, u.Whh, H, 0, d->gates_pre, 4 * H, B, 4 * H, H, nullptr, nullptr,
            d->xg + (size_t)t * B * 4 * H, 4 * H, nullptr, 0);
        ICD_TRY(icd_lstm_pointwise_fwd(B, H, d->gates_pre, d->c_all + t * BH, d->gates_act + (size_t)t * B * 4 * H,
                                       d->c_all + (t + 1) * BH, d->h_all + (t + 1) * BH,
                                       d->hout + (size_t)t * H, (int64_t)L * H, nullptr, 1.f, s,
                                       at16(u.h, (int64_t)(t + 1) * BH), at16(u.hout, (int64_t)t * H)));
    }
    BMM(u.hout, H, 0, u.Wlin, H, 0, d->outputs, V, B * L, V, H, d->lin_b, nullptr, nullptr, 0, nullptr, 0);      // :109
    return 0;
}

int baseline_bwd_bf16(const icd_base_desc_t* d, cudaStream_t s) {
    Arena16 ar; BaseBufs u;
    ICD_TRY(check_base16(d, ar, u));
    const int B = d->B, L = d->L, E = d->E, H = d->H, V = d->V;
    const size_t BH = (size_t)B * H;
    const int LB = L * B;
    const float* NF = nullptr;
    BCVT(d->d_outputs, V, LB, V, u.dY, u.ldV);
    BMM(u.dY, u.ldV, 0, u.Wlin, H, 1, d->d_hout, H, LB, H, V, NF, NF, NF, 0, nullptr, 0);              // d_hout = dY W_lin
    BMM(u.dY, u.ldV, 1, u.hout, H, 1, d->d_lin_w, H, V, H, LB, NF, NF, NF, 0, nullptr, 0);             // dW_lin = dY^T hout
    ICD_TRY(icd_colsum_bf16(u.dY, u.ldV, (int64_t)LB, V, nullptr, d->d_lin_b, u.colsum_ws, s));
    ICD_CUDA(cudaMemsetAsync(d->dh, 0, sizeof(float) * BH, s));
    ICD_CUDA(cudaMemsetAsync(d->dc, 0, sizeof(float) * BH, s));
    const bool persistent = icd_lstm_seq_persistent_ok(B, L, H) != 0;
    if (persistent)
        ICD_TRY(icd_lstm_seq_bwd_persistent(B, L, H, d->w_hh, d->d_hout, d->gates_act, d->c_all, d->dc, d->dg, u.dg, u.bar, s));
    else BCVT(d->w_hh, H, 4 * H, H, u.Whh, H);       // (the forward of a persistent step did not stage it)
    for (int t = L - 1; t >= 0 && !persistent; --t) {
        float* dgt = d->dg + (size_t)t * B * 4 * H;
        char* dg16 = at16(u.dg, (int64_t)t * B * 4 * H);
        ICD_TRY(icd_lstm_pointwise_bwd(B, H, d->dh, d->d_hout + (size_t)t * H, (int64_t)L * H, nullptr, 1.f, d->dc,
                                       d->gates_act + (size_t)t * B * 4 * H, d->c_all + t * BH, d->c_all + (t + 1) * BH,
                                       dgt, 4 * H, s, dg16, 4 * H));
        BMM(dg16, 4 * H, 0, u.Whh, H, 1, d->dh, H, B, H, 4 * H, NF, NF, NF, 0, nullptr, 0);            // dh = dg W_hh
    }
    BMM(u.dg, 4 * H, 1, u.h, H, 1, d->d_w_hh, H, 4 * H, H, LB, NF, NF, NF, 0, nullptr, 0);             // dW_hh = dg^T h_prev
    BMM(u.dg, 4 * H, 1, u.x, u.ldE, 1, d->d_w_ih, E, 4 * H, E, LB, NF, NF, NF, 0, nullptr, 0);         // dW_ih = dg^T x
    ICD_TRY(icd_colsum(d->dg, 4 * H, LB, 4 * H, nullptr, d->d_b, s));
    if (d->d_img_features || d->d_emb_w) {
        BMM(u.dg, 4 * H, 0, u.Wih, u.ldE, 1, d->d_x, E, LB, E, 4 * H, NF, NF, NF, 0, nullptr, 0);      // d_x = dg W_ih
        if (d->d_img_features)
            ICD_CUDA(cudaMemcpyAsync(d->d_img_features, d->d_x, sizeof(float) * (size_t)B * E, cudaMemcpyDeviceToDevice, s));
        if (d->d_emb_w) {
            ICD_CUDA(cudaMemsetAsync(d->d_emb_w, 0, (d->emb_is_f64 ? sizeof(double) : sizeof(float)) * (size_t)V * E, s));
            if (L > 1) {
                int32_t bt[ICD_MAX_STEPS];
                for (int t = 0; t < ICD_MAX_STEPS; ++t) bt[t] = B;
                ICD_TRY(icd_embed_scatter_add(d->d_emb_w, d->emb_is_f64, d->captions, B, L, L - 1, E, V, bt,
                                              d->d_x + (size_t)B * E, s, d->d_hout, (int64_t)sizeof(float) * B * L * H));   // d_hout: free after BPTT
            }
        }
    }
    return 0;
}
}  // namespace

// fp32-grade tier: the split of W_hh (read by every step of the recurrence) is made once per call, kept at the tail of tc_ws
static int64_t base_x3_cache_bytes(const icd_base_desc_t* d) {
    return d->precision == ICD_PREC_FP32X3 ? icd_x3_split_bytes(4 * (int64_t)d->H, d->H) + 512 : 0;
}

extern "C" int64_t icd_baseline_decoder_ws_bytes(const icd_base_desc_t* d) {
    if (d && d->precision == ICD_PREC_FP32X3) {
        const int64_t B = d->B, L = d->L, E = d->E, H = d->H, V = d->V, LB = L * B;
        const int64_t shapes[][3] = {{LB, 4 * H, E}, {B, 4 * H, H}, {LB, V, H}, {LB, H, V}, {V, H, LB}, {B, H, 4 * H},
                                     {4 * H, H, LB}, {4 * H, E, LB}, {LB, E, 4 * H}};
        int64_t need = 0;
        for (const auto& sh : shapes) need = std::max(need, icd_gemm_ws_bytes((int)sh[0], (int)sh[1], (int)sh[2], ICD_PREC_FP32X3));
        return need + base_x3_cache_bytes(d);
    }
    if (!d || d->precision != ICD_PREC_BF16) return 0;
    Arena16 a; a.base = nullptr; a.cap = 0; a.off = 0; a.ok = true;
    BaseBufs b;
    carve_base(d, a, b);
    return a.off;
}

extern "C" int icd_baseline_decoder_fwd(const icd_base_desc_t* d, void* stream) {
    ICD_TRY(check(d));
    cudaStream_t s = icd_stream(stream);
    if (d->precision == ICD_PREC_BF16) return baseline_fwd_bf16(d, s);
    ICD_CHECK_ARG(d->precision == ICD_PREC_FP32 || d->precision == ICD_PREC_FP32X3, "baseline_decoder: unknown precision %d", d->precision);
    IcdSimpleWsScope ws_scope(d->tc_ws, d->tc_ws_bytes, base_x3_cache_bytes(d));
    icd_x3_cache_mark(d->w_hh);
    const int B = d->B, L = d->L, E = d->E, H = d->H, V = d->V, prec = d->precision;
    const size_t BH = (size_t)B * H;
    // x[0] = img_features, x[t] = embedding(captions[:, t-1])   (:93-101)
    ICD_CUDA(cudaMemcpyAsync(d->x, d->img_features, sizeof(float) * (size_t)B * E, cudaMemcpyDeviceToDevice, s));
    if (L > 1) ICD_TRY(icd_embed_gather(d->emb_w, d->emb_is_f64, d->captions, B, L, L - 1, E, V, d->x + (size_t)B * E, s));
    ICD_TRY(icd_gemm_simple(prec, d->x, E, 1, d->w_ih, E, 1, d->xg, 4 * H, L * B, 4 * H, E, d->b_ih, d->b_hh,
                            nullptr, 0, nullptr, 0, nullptr, 0.f, s));
    ICD_CUDA(cudaMemsetAsync(d->h_all, 0, sizeof(float) * BH, s));                           // zero (h0, c0) (:106)
    ICD_CUDA(cudaMemsetAsync(d->c_all, 0, sizeof(float) * BH, s));
    for (int t = 0; t < L; ++t) {
        ICD_TRY(icd_gemm_simple(prec, d->h_all + t * BH, H, 1, d->w_hh, H, 1, d->gates_pre, 4 * H, B, 4 * H, H,
                                nullptr, nullptr, d->xg + (size_t)t * B * 4 * H, 4 * H, nullptr, 0, nullptr, 0.f, s));
        ICD_TRY(icd_lstm_pointwise_fwd(B, H, d->gates_pre, d->c_all + t * BH, d->gates_act + (size_t)t * B * 4 * H,
                                       d->c_all + (t + 1) * BH, d->h_all + (t + 1) * BH,
                                       d->hout + (size_t)t * H, (int64_t)L * H, nullptr, 1.f, s));
    }
    ICD_TRY(icd_gemm_simple(prec, d->hout, H, 1, d->lin_w, H, 1, d->outputs, V, B * L, V, H, d->lin_b, nullptr,
                            nullptr, 0, nullptr, 0, nullptr, 0.f, s));                       // :109
    return 0;
}

extern "C" int icd_baseline_decoder_bwd(const icd_base_desc_t* d, void* stream) {
    ICD_TRY(check(d));
    cudaStream_t s = icd_stream(stream);
    if (d->precision == ICD_PREC_BF16) return baseline_bwd_bf16(d, s);
    IcdSimpleWsScope ws_scope(d->tc_ws, d->tc_ws_bytes, base_x3_cache_bytes(d));
    icd_x3_cache_mark(d->w_hh);
    const int B = d->B, L = d->L, E = d->E, H = d->H, V = d->V, prec = d->precision;
    const size_t BH = (size_t)B * H;
    const int LB = L * B;
    // linear (:109)
    ICD_TRY(icd_gemm_simple(prec, d->d_outputs, V, 1, d->lin_w, 1, H, d->d_hout, H, B * L, H, V,
                            nullptr, nullptr, nullptr, 0, nullptr, 0, nullptr, 0.f, s));
    ICD_TRY(icd_gemm_simple(prec, d->d_outputs, 1, V, d->hout, 1, H, d->d_lin_w, H, V, H, B * L,
                            nullptr, nullptr, nullptr, 0, nullptr, 0, nullptr, 0.f, s, ICD_GEMM_ALLOW_SPLITK));
    ICD_TRY(icd_colsum(d->d_outputs, V, (int64_t)B * L, V, nullptr, d->d_lin_b, s));
    // BPTT (:106)
    ICD_CUDA(cudaMemsetAsync(d->dh, 0, sizeof(float) * BH, s));
    ICD_CUDA(cudaMemsetAsync(d->dc, 0, sizeof(float) * BH, s));
    for (int t = L - 1; t >= 0; --t) {
        float* dgt = d->dg + (size_t)t * B * 4 * H;
        ICD_TRY(icd_lstm_pointwise_bwd(B, H, d->dh, d->d_hout + (size_t)t * H, (int64_t)L * H, nullptr, 1.f, d->dc,
                                       d->gates_act + (size_t)t * B * 4 * H, d->c_all + t * BH, d->c_all + (t + 1) * BH,
                                       dgt, 4 * H, s));
        ICD_TRY(icd_gemm_simple(prec, dgt, 4 * H, 1, d->w_hh, 1, H, d->dh, H, B, H, 4 * H,
                                nullptr, nullptr, nullptr, 0, nullptr, 0, nullptr, 0.f, s));
    }
    // hoisted weight gradients
    ICD_TRY(icd_gemm_simple(prec, d->dg, 1, 4 * H, d->h_all, 1, H, d->d_w_hh, H, 4 * H, H, LB,
                            nullptr, nullptr, nullptr, 0, nullptr, 0, nullptr, 0.f, s, ICD_GEMM_ALLOW_SPLITK));
    ICD_TRY(icd_gemm_simple(prec, d->dg, 1, 4 * H, d->x, 1, E, d->d_w_ih, E, 4 * H, E, LB,
                            nullptr, nullptr, nullptr, 0, nullptr, 0, nullptr, 0.f, s, ICD_GEMM_ALLOW_SPLITK));
    ICD_TRY(icd_colsum(d->dg, 4 * H, LB, 4 * H, nullptr, d->d_b, s));
    if (d->d_img_features || d->d_emb_w) {
        ICD_TRY(icd_gemm_simple(prec, d->dg, 4 * H, 1, d->w_ih, 1, E, d->d_x, E, LB, E, 4 * H,
                                nullptr, nullptr, nullptr, 0, nullptr, 0, nullptr, 0.f, s));
        if (d->d_img_features)
            ICD_CUDA(cudaMemcpyAsync(d->d_img_features, d->d_x, sizeof(float) * (size_t)B * E, cudaMemcpyDeviceToDevice, s));
        if (d->d_emb_w) {
            ICD_CUDA(cudaMemsetAsync(d->d_emb_w, 0, (d->emb_is_f64 ? sizeof(double) : sizeof(float)) * (size_t)V * E, s));
            if (L > 1) {
                int32_t bt[ICD_MAX_STEPS];
                for (int t = 0; t < ICD_MAX_STEPS; ++t) bt[t] = B;
                ICD_TRY(icd_embed_scatter_add(d->d_emb_w, d->emb_is_f64, d->captions, B, L, L - 1, E, V, bt,
                                              d->d_x + (size_t)B * E, s, d->d_hout, (int64_t)sizeof(float) * B * L * H));   // d_hout: free after BPTT
            }
        }
    }
    return 0;
}
