// bf16 tensor-core tier of AttentionDecoder.forward / backward (BASELINE.json configs[2]: "bf16 GEMMs").
//
// Same dataflow as decoder_loop.cu (the fp32 parity tier); every dense contraction runs on the tcgen05/TMA kernel
// of gemm_tc.cu with bf16 operands and fp32 accumulation, everything else (attention step, softmax, LSTM gate math,
// reductions, all saved activations and all gradients) stays fp32.  bf16 K-major operand copies live in a
// caller-provided arena (desc->tc_ws):
//   forward, kept for backward : weights (W_e, [W_dec;W_fb;W_hh], W_ih split at E, W_h, W_c, W_fc), enc, mean, emb_x,
//                                h_t, gated_t, dropout(h)
//   backward                   : transposed weight copies for the dX contractions, dY and dY^T, per-step dz,
//                                and the transposed activation stacks feeding the hoisted dW contractions.
// The transposes are produced by the fp32->bf16 conversion pass itself (convert_transpose_kernel), so the tensor-core
// kernel only ever sees K-major operands.
#include "common.cuh"
#include "gemm_tc.cuh"

namespace {

struct BtPack { int v[ICD_MAX_STEPS]; };

__global__ void row_valid_kernel16(int B, int T, const BtPack bt, unsigned char* __restrict__ valid) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * T) return;
    const int b = i / T, t = i % T;
    valid[i] = (b < bt.v[t]) ? 1 : 0;
}

inline int64_t up8(int64_t x) { return (x + 7) / 8 * 8; }

// bump allocator over the caller's arena; element type is bf16 (2 bytes)
struct Arena {
    char* base; int64_t cap, off; bool ok;
    void* take(int64_t rows, int64_t ld) {
        const int64_t bytes = (rows * ld * 2 + 255) / 256 * 256;
        void* p = base ? base + off : nullptr;
        off += bytes;
        if (base && off > cap) ok = false;
        return p;
    }
};

// every bf16 buffer of the tier, carved in a fixed order so forward and backward agree on the layout
struct Bufs {
    void *We, *Wcat, *WihE, *WihC, *Wh, *Wc, *Wfc;                 // weights, K-major
    void *enc, *att_enc, *mean, *embx, *h, *gated, *hdrop;         // forward activations, K-major
    void *WcatT, *WihCT, *WihET, *WfcT;                            // backward: transposed weights
    void *dY, *dYT, *hdropT, *dz, *dzT, *hT, *embxT, *gatedT, *dhT, *dcT, *meanT, *daeT, *encT;
    int64_t ldE, ldV, ldTB, ldB, ldBP, ldBT;
};

void carve(const icd_att_desc_t* d, Arena& a, Bufs& b) {
    const int64_t B = d->B, T = d->T, P = d->P, C = d->C, A = d->A, D = d->D, E = d->E, V = d->V;
    const int64_t NZ = A + C + 4 * D, TB = T * B;
    b.ldE = up8(E); b.ldV = up8(V); b.ldTB = up8(TB); b.ldB = up8(B); b.ldBP = up8(B * P); b.ldBT = up8(B * T);
    b.We = a.take(A, C); b.Wcat = a.take(NZ, D); b.WihE = a.take(4 * D, b.ldE); b.WihC = a.take(4 * D, C);
    b.Wh = a.take(D, C); b.Wc = a.take(D, C); b.Wfc = a.take(V, D);
    b.enc = a.take(B * P, C); b.att_enc = a.take(B * P, A); b.mean = a.take(B, C); b.embx = a.take(TB, b.ldE);
    b.h = a.take(TB + B, D);
    b.gated = a.take(TB, C); b.hdrop = a.take(B * T, D);
    b.WcatT = a.take(D, NZ); b.WihCT = a.take(C, 4 * D); b.WihET = a.take(E, 4 * D); b.WfcT = a.take(D, b.ldV);
    b.dY = a.take(B * T, b.ldV); b.dYT = a.take(V, b.ldBT); b.hdropT = a.take(D, b.ldBT);
    b.dz = a.take(TB, NZ); b.dzT = a.take(NZ, b.ldTB); b.hT = a.take(D, b.ldTB); b.embxT = a.take(E, b.ldTB);
    b.gatedT = a.take(C, b.ldTB); b.dhT = a.take(D, b.ldB); b.dcT = a.take(D, b.ldB); b.meanT = a.take(C, b.ldB);
    b.daeT = a.take(A, b.ldBP); b.encT = a.take(C, b.ldBP);
}

int check16(const icd_att_desc_t* d, Arena& a, Bufs& b) {
    ICD_CHECK_ARG(d->A % 8 == 0 && d->C % 8 == 0 && d->D % 8 == 0 && d->E % 4 == 0,
                  "attention_decoder(bf16): A, C, D must be multiples of 8 and E of 4 (A=%d C=%d D=%d E=%d)", d->A, d->C, d->D, d->E);
    a.base = reinterpret_cast<char*>(d->tc_ws); a.cap = d->tc_ws_bytes; a.off = 0; a.ok = true;
    carve(d, a, b);
    ICD_CHECK_ARG(d->tc_ws != nullptr && a.ok, "attention_decoder(bf16): tc_ws too small (%lld bytes given, %lld needed)",
                  (long long)d->tc_ws_bytes, (long long)a.off);
    return 0;
}

#define CVT(src, sr, sc, rows, cols, dst, ld) ICD_TRY(icd_convert_bf16((src), (sr), (sc), (rows), (cols), (dst), (ld), s))
#define MM(A16, lda, B16, ldb, Cp, ldc, M, N, K, b1, b2, a1, l1, a2, l2, mask, beta) \
    ICD_TRY(icd_gemm_bf16((A16), (lda), (B16), (ldb), (Cp), (ldc), (M), (N), (K), (b1), (b2), (a1), (l1), (a2), (l2), (mask), (beta), s))

inline char* at16(void* p, int64_t elem_off) { return reinterpret_cast<char*>(p) + elem_off * 2; }

}  // namespace

int64_t icd_att_tc_ws_bytes(const icd_att_desc_t* d) {
    Arena a; a.base = nullptr; a.cap = 0; a.off = 0; a.ok = true;
    Bufs b;
    carve(d, a, b);
    return a.off;
}

int icd_attention_decoder_fwd_bf16(const icd_att_desc_t* d, cudaStream_t s) {
    Arena ar; Bufs u;
    ICD_TRY(check16(d, ar, u));
    const int B = d->B, T = d->T, P = d->P, C = d->C, A = d->A, D = d->D, E = d->E, V = d->V;
    const int NZ = A + C + 4 * D, TB = T * B;
    const long long BD = (long long)B * D;

    BtPack pack;
    for (int t = 0; t < ICD_MAX_STEPS; ++t) pack.v[t] = t < T ? d->bt_host[t] : 0;
    row_valid_kernel16<<<(B * T + 255) / 256, 256, 0, s>>>(B, T, pack, d->row_valid);
    ICD_LAUNCH_CHECK();
    ICD_CUDA(cudaMemsetAsync(d->alphas, 0, sizeof(float) * (size_t)B * T * P, s));
    ICD_CUDA(cudaMemsetAsync(d->hdrop, 0, sizeof(float) * (size_t)B * T * D, s));
    ICD_CUDA(cudaMemsetAsync(u.hdrop, 0, (size_t)B * T * D * 2, s));
    if (d->bt_host[T - 1] < B) {
        ICD_CUDA(cudaMemsetAsync(d->h_all, 0, sizeof(float) * (size_t)(T + 1) * BD, s));
        ICD_CUDA(cudaMemsetAsync(d->c_all, 0, sizeof(float) * (size_t)(T + 1) * BD, s));
        ICD_CUDA(cudaMemsetAsync(d->gated, 0, sizeof(float) * (size_t)T * B * C, s));
        ICD_CUDA(cudaMemsetAsync(d->z, 0, sizeof(float) * (size_t)T * B * NZ, s));
        ICD_CUDA(cudaMemsetAsync(u.h, 0, (size_t)(TB + B) * D * 2, s));
        ICD_CUDA(cudaMemsetAsync(u.gated, 0, (size_t)TB * C * 2, s));
    }
    ICD_CUDA(cudaMemcpyAsync(d->w_cat, d->dec_att_w, sizeof(float) * (size_t)A * D, cudaMemcpyDeviceToDevice, s));
    ICD_CUDA(cudaMemcpyAsync(d->w_cat + (size_t)A * D, d->f_beta_w, sizeof(float) * (size_t)C * D, cudaMemcpyDeviceToDevice, s));
    ICD_CUDA(cudaMemcpyAsync(d->w_cat + (size_t)(A + C) * D, d->w_hh, sizeof(float) * (size_t)4 * D * D, cudaMemcpyDeviceToDevice, s));
    ICD_CUDA(cudaMemcpyAsync(d->b_cat, d->dec_att_b, sizeof(float) * A, cudaMemcpyDeviceToDevice, s));
    ICD_CUDA(cudaMemcpyAsync(d->b_cat + A, d->f_beta_b, sizeof(float) * C, cudaMemcpyDeviceToDevice, s));
    ICD_CUDA(cudaMemsetAsync(d->b_cat + A + C, 0, sizeof(float) * 4 * D, s));

    // ---- bf16 K-major copies of the weights (they change every optimiser step) and of the features ----
    CVT(d->enc_att_w, C, 1, A, C, u.We, C);
    CVT(d->w_cat, D, 1, NZ, D, u.Wcat, D);
    CVT(d->w_ih, E + C, 1, 4 * D, E, u.WihE, u.ldE);
    CVT(d->w_ih + E, E + C, 1, 4 * D, C, u.WihC, C);
    CVT(d->h_lin_w, C, 1, D, C, u.Wh, C);
    CVT(d->c_lin_w, C, 1, D, C, u.Wc, C);
    CVT(d->fc_w, D, 1, V, D, u.Wfc, D);
    CVT(d->enc, C, 1, B * P, C, u.enc, C);

    // K1: att_enc = enc_att(encoder_out), once per batch (models/attention.py:54)
    ICD_TRY(icd_gemm_bf16(u.enc, C, u.We, C, nullptr, 0, B * P, A, C, d->enc_att_b, nullptr, nullptr, 0, nullptr, 0,
                          nullptr, 0.f, s, u.att_enc, A));
    // K7: init_hidden_state (:161-163)
    ICD_TRY(icd_weighted_pixel_sum(B, P, C, nullptr, d->enc, nullptr, 0, nullptr, 0, d->mean_enc, nullptr, nullptr, s));
    CVT(d->mean_enc, C, 1, B, C, u.mean, C);
    MM(u.mean, C, u.Wh, C, d->h_all, D, B, D, C, d->h_lin_b, nullptr, nullptr, 0, nullptr, 0, nullptr, 0.f);
    MM(u.mean, C, u.Wc, C, d->c_all, D, B, D, C, d->c_lin_b, nullptr, nullptr, 0, nullptr, 0, nullptr, 0.f);
    CVT(d->h_all, D, 1, B, D, u.h, D);                              // h_0 (bf16); h_t, t >= 1, is emitted by the gate kernel
    // K5: embedding lookup (:247) + hoisted input contraction
    ICD_TRY(icd_embed_gather(d->emb_w, d->emb_is_f64, d->captions, B, d->L, T, E, d->emb_x, s));
    CVT(d->emb_x, E, 1, TB, E, u.embx, u.ldE);
    MM(u.embx, u.ldE, u.WihE, u.ldE, d->xg, 4 * D, TB, 4 * D, E, d->b_ih, d->b_hh, nullptr, 0, nullptr, 0, nullptr, 0.f);

    for (int t = 0; t < T; ++t) {
        const int bt = d->bt_host[t];
        if (bt == 0) break;
        const float* h_prev = d->h_all + (size_t)t * BD;
        const float* c_prev = d->c_all + (size_t)t * BD;
        float* zt = d->z + (size_t)t * B * NZ;
        char* h16 = at16(u.h, (int64_t)t * B * D);
        char* g16 = at16(u.gated, (int64_t)t * B * C);
        MM(h16, D, u.Wcat, D, zt, NZ, bt, NZ, D, d->b_cat, nullptr, nullptr, 0, nullptr, 0, nullptr, 0.f);       // K2
        ICD_TRY(icd_attention_step_fwd_bf16(bt, P, C, A, nullptr, u.enc, u.att_enc, zt, NZ, d->full_att_w, d->full_att_b,
                                            zt + A, NZ, d->alphas + (size_t)t * P, (int64_t)T * P,
                                            d->awe_raw + (size_t)t * B * C, d->gate + (size_t)t * B * C,
                                            d->gated + (size_t)t * B * C, g16, (void*)s));                         // K3
        MM(g16, C, u.WihC, C, d->gates_pre, 4 * D, bt, 4 * D, C, nullptr, nullptr,
           d->xg + (size_t)t * B * 4 * D, 4 * D, zt + A + C, NZ, nullptr, 0.f);                                   // K4
        ICD_TRY(icd_lstm_pointwise_fwd(bt, D, d->gates_pre, c_prev, d->gates_act + (size_t)t * B * 4 * D,
                                       d->c_all + (size_t)(t + 1) * BD, d->h_all + (size_t)(t + 1) * BD,
                                       d->hdrop + (size_t)t * D, (int64_t)T * D,
                                       d->drop_mask ? d->drop_mask + (size_t)t * BD : nullptr, d->drop_scale, s,
                                       at16(u.h, (int64_t)(t + 1) * B * D), at16(u.hdrop, (int64_t)t * D)));
    }
    // K6: predictions = fc(dropout(h)) for every (b,t) at once (:279-280); inactive rows exactly 0 (:253)
    MM(u.hdrop, D, u.Wfc, D, d->predictions, V, B * T, V, D, d->fc_b, nullptr, nullptr, 0, nullptr, 0, d->row_valid, 0.f);
    return 0;
}

int icd_attention_decoder_bwd_bf16(const icd_att_desc_t* d, cudaStream_t s) {
    Arena ar; Bufs u;
    ICD_TRY(check16(d, ar, u));
    const int B = d->B, T = d->T, P = d->P, C = d->C, A = d->A, D = d->D, E = d->E, V = d->V;
    const int NZ = A + C + 4 * D, TB = T * B, BT = B * T, BP = B * P;
    const long long BD = (long long)B * D;

    ICD_CUDA(cudaMemsetAsync(d->dz, 0, sizeof(float) * (size_t)TB * NZ, s));
    ICD_CUDA(cudaMemsetAsync(d->d_e, 0, sizeof(float) * (size_t)B * T * P, s));
    ICD_CUDA(cudaMemsetAsync(d->dh, 0, sizeof(float) * (size_t)BD, s));
    ICD_CUDA(cudaMemsetAsync(d->dc, 0, sizeof(float) * (size_t)BD, s));
    ICD_CUDA(cudaMemsetAsync(u.dz, 0, (size_t)TB * NZ * 2, s));

    // transposed bf16 weights for the dX contractions (B operand must be K-major: B[n, k] = W[k, n])
    CVT(d->w_cat, 1, D, D, NZ, u.WcatT, NZ);                       // [D x NZ]
    CVT(d->w_ih + E, 1, E + C, C, 4 * D, u.WihCT, 4 * D);          // [C x 4D]
    CVT(d->fc_w, 1, D, D, V, u.WfcT, u.ldV);                       // [D x V]

    // ---- fc (:279): d_hdrop = dY W_fc ; dW_fc = dY^T hdrop ; db_fc = masked column sum of dY ----
    CVT(d->d_predictions, V, 1, BT, V, u.dY, u.ldV);
    CVT(d->d_predictions, 1, V, V, BT, u.dYT, u.ldBT);
    CVT(d->hdrop, 1, D, D, BT, u.hdropT, u.ldBT);
    MM(u.dY, u.ldV, u.WfcT, u.ldV, d->d_hdrop, D, BT, D, V, nullptr, nullptr, nullptr, 0, nullptr, 0, nullptr, 0.f);
    MM(u.dYT, u.ldBT, u.hdropT, u.ldBT, d->d_fc_w, D, V, D, BT, nullptr, nullptr, nullptr, 0, nullptr, 0, nullptr, 0.f);
    ICD_TRY(icd_colsum(d->d_predictions, V, (int64_t)BT, V, d->row_valid, d->d_fc_b, s));

    // ---- BPTT ----
    for (int t = T - 1; t >= 0; --t) {
        const int bt = d->bt_host[t];
        if (bt == 0) continue;
        float* dzt = d->dz + (size_t)t * B * NZ;
        const float* zt = d->z + (size_t)t * B * NZ;
        char* dz16 = at16(u.dz, (int64_t)t * B * NZ);
        ICD_TRY(icd_lstm_pointwise_bwd(bt, D, d->dh, d->d_hdrop + (size_t)t * D, (int64_t)T * D,
                                       d->drop_mask ? d->drop_mask + (size_t)t * BD : nullptr, d->drop_scale,
                                       d->dc, d->gates_act + (size_t)t * B * 4 * D,
                                       d->c_all + (size_t)t * BD, d->c_all + (size_t)(t + 1) * BD,
                                       dzt + A + C, NZ, s, at16(dz16, A + C), NZ));     // dG also emitted as bf16
        MM(at16(dz16, A + C), NZ, u.WihCT, 4 * D, d->d_gated, C, bt, C, 4 * D,
           nullptr, nullptr, nullptr, 0, nullptr, 0, nullptr, 0.f);                // d_gated = dG W_ih[:, E:]
        ICD_TRY(icd_attention_step_bwd_bf16(bt, P, C, A, u.enc, u.att_enc, zt, NZ, d->full_att_w,
                                            d->alphas + (size_t)t * P, (int64_t)T * P,
                                            d->d_alphas ? d->d_alphas + (size_t)t * P : nullptr, (int64_t)T * P,
                                            d->gate + (size_t)t * B * C, d->awe_raw + (size_t)t * B * C, d->d_gated,
                                            dzt, NZ, dzt + A, NZ, d->d_e + (size_t)t * P, (int64_t)T * P,
                                            dz16, NZ, (void*)s));                   // d att_dec | d fbeta_pre (+ bf16)
        MM(dz16, NZ, u.WcatT, NZ, d->dh, D, bt, D, NZ, nullptr, nullptr, nullptr, 0, nullptr, 0, nullptr, 0.f);
    }

    // ---- init_hidden_state (:161-163): dh, dc now hold d h0, d c0 ----
    CVT(d->dh, 1, D, D, B, u.dhT, u.ldB);
    CVT(d->dc, 1, D, D, B, u.dcT, u.ldB);
    CVT(d->mean_enc, 1, C, C, B, u.meanT, u.ldB);
    MM(u.dhT, u.ldB, u.meanT, u.ldB, d->d_h_lin_w, C, D, C, B, nullptr, nullptr, nullptr, 0, nullptr, 0, nullptr, 0.f);
    ICD_TRY(icd_colsum(d->dh, D, B, D, nullptr, d->d_h_lin_b, s));
    MM(u.dcT, u.ldB, u.meanT, u.ldB, d->d_c_lin_w, C, D, C, B, nullptr, nullptr, nullptr, 0, nullptr, 0, nullptr, 0.f);
    ICD_TRY(icd_colsum(d->dc, D, B, D, nullptr, d->d_c_lin_b, s));

    // ---- hoisted weight gradients over the T*B stacked rows ----
    CVT(d->dz, 1, NZ, NZ, TB, u.dzT, u.ldTB);                      // [NZ x TB]
    CVT(d->h_all, 1, D, D, TB, u.hT, u.ldTB);                      // [D x TB]   (h_0 .. h_{T-1})
    CVT(d->emb_x, 1, E, E, TB, u.embxT, u.ldTB);                   // [E x TB]
    CVT(d->gated, 1, C, C, TB, u.gatedT, u.ldTB);                  // [C x TB]
    MM(u.dzT, u.ldTB, u.hT, u.ldTB, d->d_w_cat, D, NZ, D, TB, nullptr, nullptr, nullptr, 0, nullptr, 0, nullptr, 0.f);
    ICD_TRY(icd_colsum(d->dz, NZ, TB, NZ, nullptr, d->d_b_cat, s));
    char* dGT = at16(u.dzT, (int64_t)(A + C) * u.ldTB);           // rows [A+C, NZ) of dz^T = dG^T  [4D x TB]
    MM(dGT, u.ldTB, u.embxT, u.ldTB, d->d_w_ih, E + C, 4 * D, E, TB, nullptr, nullptr, nullptr, 0, nullptr, 0, nullptr, 0.f);
    MM(dGT, u.ldTB, u.gatedT, u.ldTB, d->d_w_ih + E, E + C, 4 * D, C, TB, nullptr, nullptr, nullptr, 0, nullptr, 0, nullptr, 0.f);
    if (d->d_emb_w) {
        CVT(d->w_ih, 1, E + C, E, 4 * D, u.WihET, 4 * D);          // [E x 4D]
        MM(at16(u.dz, A + C), NZ, u.WihET, 4 * D, d->d_emb_x, E, TB, E, 4 * D,
           nullptr, nullptr, nullptr, 0, nullptr, 0, nullptr, 0.f);
        ICD_CUDA(cudaMemsetAsync(d->d_emb_w, 0, (d->emb_is_f64 ? sizeof(double) : sizeof(float)) * (size_t)V * E, s));
        ICD_TRY(icd_embed_scatter_add(d->d_emb_w, d->emb_is_f64, d->captions, B, d->L, T, E, d->bt_host, d->d_emb_x, s));
    }
    // attention projections: d_att_enc for all steps at once, full_att grads, then enc_att grads (:54)
    ICD_TRY(icd_attention_proj_bwd_bf16(B, T, P, A, d->bt_host, u.att_enc, d->z, NZ, d->full_att_w, d->d_e,
                                        d->d_att_enc, d->d_full_att_w, d->d_full_att_b, d->proj_partial, (void*)s));
    ICD_TRY(icd_colsum(d->d_att_enc, A, (int64_t)BP, A, nullptr, d->d_enc_att_b, s));
    CVT(d->d_att_enc, 1, A, A, BP, u.daeT, u.ldBP);                // [A x BP]
    CVT(d->enc, 1, C, C, BP, u.encT, u.ldBP);                      // [C x BP]
    MM(u.daeT, u.ldBP, u.encT, u.ldBP, d->d_enc_att_w, C, A, C, BP, nullptr, nullptr, nullptr, 0, nullptr, 0, nullptr, 0.f);
    return 0;
}
