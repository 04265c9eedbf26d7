// BaselineDecoder.forward / backward (models/baseline.py:81-111): embedding of captions[:, :-1] (:93,97), image
// feature prepended as LSTM step 0 (:101), 1-layer nn.LSTM from a zero state (:106), vocabulary linear (:109).
// The input contraction x * W_ih^T + b_ih + b_hh is hoisted over all L steps, the recurrence runs as
// h * W_hh^T (+ fused gate pointwise) per step, and the vocabulary contraction runs once over all (b, t).
#include "common.cuh"

namespace {
int check(const icd_base_desc_t* d) {
    ICD_CHECK_ARG(d != nullptr, "baseline_decoder: null descriptor");
    ICD_CHECK_ARG(d->B > 0 && d->L > 0 && d->L <= ICD_MAX_STEPS, "baseline_decoder: B=%d L=%d out of range", d->B, d->L);
    ICD_CHECK_ARG(d->E % 4 == 0 && d->H % 4 == 0, "baseline_decoder: E and H must be multiples of 4");
    return 0;
}
}  // namespace

extern "C" int icd_baseline_decoder_fwd(const icd_base_desc_t* d, void* stream) {
    ICD_TRY(check(d));
    cudaStream_t s = icd_stream(stream);
    const int B = d->B, L = d->L, E = d->E, H = d->H, V = d->V, prec = d->precision;
    const size_t BH = (size_t)B * H;
    // x[0] = img_features, x[t] = embedding(captions[:, t-1])   (:93-101)
    ICD_CUDA(cudaMemcpyAsync(d->x, d->img_features, sizeof(float) * (size_t)B * E, cudaMemcpyDeviceToDevice, s));
    if (L > 1) ICD_TRY(icd_embed_gather(d->emb_w, d->emb_is_f64, d->captions, B, L, L - 1, E, d->x + (size_t)B * E, s));
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
                ICD_TRY(icd_embed_scatter_add(d->d_emb_w, d->emb_is_f64, d->captions, B, L, L - 1, E, bt,
                                              d->d_x + (size_t)B * E, s));
            }
        }
    }
    return 0;
}
