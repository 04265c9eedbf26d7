"""ORACLE — TEST INFRASTRUCTURE ONLY.  Never imported by the product path.

A CPU restatement of the reference's captioning-decoder hot path, written as plain functions
over a ``state_dict``-style dict of weights.  It follows the reference line by line (file:line
cited at every function; all paths relative to the reference root) but is dtype-generic, so it
can be evaluated in float64 — the reference modules cannot (hard-coded ``.float()`` casts at
models/attention.py:277-278 and models/baseline.py:101).  Gradients come from torch autograd
over this restatement.

Floating-point path => the restatement is torch (CPU) rather than numpy/C; there is no integer
arithmetic here besides token ids and top-k indices.

Pinning: the reference ships NO tests, golden vectors or known-answer fixtures for this path
(SURVEY.md §4, §8c).  The oracle is therefore pinned against OUTPUTS OF THE REFERENCE ITSELF,
run in the build container: ``tests/golden/make_golden.py`` imports the unmodified reference,
runs it on seeded inputs, asserts this oracle agrees, and writes the fixtures under
``tests/golden/`` that ``tests/test_oracle_golden.py`` re-checks everywhere.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl
reference`` legs may import this module.
"""
import torch
import torch.nn.functional as F


# --------------------------------------------------------------------------------------
# weights
# --------------------------------------------------------------------------------------
ATT_KEYS = [
    "attention.enc_att.weight", "attention.enc_att.bias",
    "attention.dec_att.weight", "attention.dec_att.bias",
    "attention.full_att.weight", "attention.full_att.bias",
    "decode_step.weight_ih", "decode_step.weight_hh", "decode_step.bias_ih", "decode_step.bias_hh",
    "h_lin.weight", "h_lin.bias", "c_lin.weight", "c_lin.bias",
    "f_beta.weight", "f_beta.bias", "fc.weight", "fc.bias", "embedding.weight",
]
BASE_KEYS = [
    "embedding.weight", "lstm.weight_ih_l0", "lstm.weight_hh_l0", "lstm.bias_ih_l0",
    "lstm.bias_hh_l0", "linear.weight", "linear.bias",
]


def cast_weights(sd, dtype, requires_grad=False, frozen=()):
    """Detached copies of a state_dict in ``dtype`` (leaves for autograd)."""
    out = {}
    for k, v in sd.items():
        t = v.detach().to("cpu").to(dtype).clone()
        if requires_grad and k not in frozen:
            t.requires_grad_(True)
        out[k] = t
    return out


# --------------------------------------------------------------------------------------
# SoftAttention.forward — models/attention.py:43-61
# --------------------------------------------------------------------------------------
def soft_attention(w, encoder_out, decoder_hidden, att_enc=None):
    """(awe, alpha).  ``att_enc`` may be passed to skip the time-invariant projection
    (models/attention.py:54); the reference recomputes it at every call."""
    if att_enc is None:
        att_enc = F.linear(encoder_out, w["attention.enc_att.weight"], w["attention.enc_att.bias"])   # :54
    att_dec = F.linear(decoder_hidden, w["attention.dec_att.weight"], w["attention.dec_att.bias"])     # :55
    att = F.linear(torch.relu(att_enc + att_dec.unsqueeze(1)),                                           # :56-57
                   w["attention.full_att.weight"], w["attention.full_att.bias"]).squeeze(2)
    alpha = torch.softmax(att, dim=1)                                                                    # :58
    awe = (encoder_out * alpha.unsqueeze(2)).sum(dim=1)                                                  # :59-60
    return awe, alpha


def init_hidden_state(w, encoder_out):
    """models/attention.py:151-164 — linear, no tanh."""
    mean_enc = encoder_out.mean(dim=1)
    return (F.linear(mean_enc, w["h_lin.weight"], w["h_lin.bias"]),
            F.linear(mean_enc, w["c_lin.weight"], w["c_lin.bias"]))


def lstm_cell(x, h, c, w_ih, w_hh, b_ih, b_hh):
    """torch.nn.LSTMCell semantics (gate order i, f, g, o) — used at models/attention.py:277."""
    gates = F.linear(x, w_ih, b_ih) + F.linear(h, w_hh, b_hh)
    i, f, g, o = gates.chunk(4, dim=1)
    c2 = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(g)
    h2 = torch.sigmoid(o) * torch.tanh(c2)
    return h2, c2


# --------------------------------------------------------------------------------------
# AttentionDecoder.forward — models/attention.py:218-284
# --------------------------------------------------------------------------------------
def attention_decoder_forward(w, encoder_out, encoded_captions, caption_lengths,
                              dropout_p=0.0, dropout_masks=None, hoist=False, embeddings=None):
    """Returns (predictions, encoded_captions, decode_lengths, attention_weights).

    dropout_masks: optional list (one per step) of 0/1 keep-masks of shape (batch_size_t, D);
      the activations are multiplied by mask/(1-p) exactly like nn.Dropout in train mode
      (models/attention.py:107,279).  None => no dropout (eval mode / p = 0).
    hoist: compute enc_att once instead of once per step (numerically the same op).
    embeddings: optional (B, L, E) pre-computed caption embeddings — what ``_create_bert_embeddings`` returns in the
      reference's use_bert branch (models/attention.py:242-244); they replace the table lookup of :247.
    """
    dtype = w["fc.weight"].dtype
    B = encoder_out.size(0)
    C = encoder_out.size(-1)
    V = w["fc.weight"].size(0)
    enc = encoder_out.reshape(B, -1, C).to(dtype)                             # :230
    P = enc.size(1)
    decode_lengths = [l - 1 for l in caption_lengths]                          # :236
    T = max(decode_lengths)
    if embeddings is not None:
        emb = embeddings                                                       # :244 (use_bert)
    else:
        emb = F.embedding(encoded_captions, w["embedding.weight"])             # :247 (may be fp64 table)
    h, c = init_hidden_state(w, enc)                                           # :250
    predictions = torch.zeros(B, T, V, dtype=dtype)                            # :253
    alphas = torch.zeros(B, T, P, dtype=dtype)                                 # :257
    att_enc_all = None
    if hoist:
        att_enc_all = F.linear(enc, w["attention.enc_att.weight"], w["attention.enc_att.bias"])
    for t in range(T):                                                         # :260
        bt = sum(l > t for l in decode_lengths)                                # :261
        enc_t, h_t, c_t = enc[:bt], h[:bt], c[:bt]                             # :263-265
        awe, alpha = soft_attention(w, enc_t, h_t,
                                    None if att_enc_all is None else att_enc_all[:bt])   # :267
        gate = torch.sigmoid(F.linear(h_t, w["f_beta.weight"], w["f_beta.bias"]))        # :270
        awe = gate * awe                                                       # :271
        x = torch.cat([emb[:bt, t, :].double(), awe.double()], dim=1).to(dtype)          # :273-277
        h, c = lstm_cell(x, h_t, c_t, w["decode_step.weight_ih"], w["decode_step.weight_hh"],
                         w["decode_step.bias_ih"], w["decode_step.bias_hh"])   # :277-278
        hd = h
        if dropout_masks is not None and dropout_p > 0.0:
            hd = h * dropout_masks[t].to(dtype) / (1.0 - dropout_p)            # :279 (nn.Dropout, train)
        predictions[:bt, t, :] = F.linear(hd, w["fc.weight"], w["fc.bias"])    # :279-280
        alphas[:bt, t, :] = alpha                                              # :281
    return predictions, encoded_captions, decode_lengths, alphas


def attention_loss(predictions, encoded_captions, decode_lengths, alphas, alpha_c=1.0):
    """Loss glue of the reference train loop — models/attention.py:401-414.

    pack_padded_sequence(batch_first=True) => time-major concatenation of the first
    batch_size_t rows of each step; CrossEntropyLoss() (mean, NO ignore_index); plus the
    doubly-stochastic regulariser ((alpha_c - sum_t alpha)^2).mean()."""
    targets = encoded_captions[:, 1:]                                          # :401
    T = max(decode_lengths)
    rows_s, rows_t = [], []
    for t in range(T):
        bt = sum(l > t for l in decode_lengths)
        rows_s.append(predictions[:bt, t, :])
        rows_t.append(targets[:bt, t])
    scores = torch.cat(rows_s, 0)                                              # :405-406
    tgt = torch.cat(rows_t, 0)                                                 # :407-408
    loss = F.cross_entropy(scores, tgt)                                        # :411
    loss = loss + ((alpha_c - alphas.sum(dim=1)) ** 2).mean()                  # :414
    return loss


def teacher_forced_ids(predictions, decode_lengths):
    """"Greedy" ids of the reference = argmax over teacher-forced logits, truncated to
    decode_lengths — models/attention.py:544-548."""
    _, ids = torch.max(predictions, dim=2)
    return [ids[j, :decode_lengths[j]].tolist() for j in range(ids.size(0))]


# --------------------------------------------------------------------------------------
# BaselineDecoder.forward — models/baseline.py:81-111
# --------------------------------------------------------------------------------------
def baseline_decoder_forward(w, img_features, captions):
    dtype = w["linear.weight"].dtype
    cap = captions[:, :-1]                                                     # :93
    emb = F.embedding(cap, w["embedding.weight"])                              # :97
    x = torch.cat((img_features.unsqueeze(1).to(dtype), emb.to(dtype)), dim=1)  # :101
    B, L, _ = x.shape
    H = w["lstm.weight_hh_l0"].size(1)
    h = torch.zeros(B, H, dtype=dtype)                                         # :106 zero (h0, c0)
    c = torch.zeros(B, H, dtype=dtype)
    outs = []
    for t in range(L):
        h, c = lstm_cell(x[:, t], h, c, w["lstm.weight_ih_l0"], w["lstm.weight_hh_l0"],
                         w["lstm.bias_ih_l0"], w["lstm.bias_hh_l0"])
        outs.append(h)
    lstm_out = torch.stack(outs, dim=1)
    return F.linear(lstm_out, w["linear.weight"], w["linear.bias"])           # :109


def baseline_loss(outputs, captions, pad_id=0):
    """models/baseline.py:194-195, 224-225 — CE with ignore_index=PAD over all L positions."""
    return F.cross_entropy(outputs.reshape(-1, outputs.shape[2]), captions.reshape(-1),
                           ignore_index=pad_id)


# --------------------------------------------------------------------------------------
# Beam search — gen_captions.py:16-131 (state machine of SURVEY.md Appendix C)
# --------------------------------------------------------------------------------------
def beam_search(w, encoder_out, beam_size, start_id, end_id, max_steps=50, trace=None):
    """One image.  Returns (seq, alphas, Caption_End) exactly like the reference:
    seq includes <start> and <end>; alphas is a nested list (len(seq) x 14 x 14) whose first
    frame is all ones; on failure ([start, end], [], False).

    ``max_steps`` mirrors the hard-coded ``step > 50`` break (gen_captions.py:119): the loop
    body runs for step = 1 .. max_steps+1.  ``trace`` (list) receives next-word ids per step
    — the reference prints them (gen_captions.py:91)."""
    dtype = w["fc.weight"].dtype
    k = beam_size
    V = w["fc.weight"].size(0)
    C = encoder_out.size(-1)
    enc_size = encoder_out.size(1)
    enc = encoder_out.reshape(1, -1, C).to(dtype)                              # :41
    P = enc.size(1)
    enc = enc.expand(k, P, C)                                                  # :44
    k_prev_words = torch.full((k, 1), start_id, dtype=torch.long)             # :47
    seqs = torch.full((k, 1), start_id, dtype=torch.long)                     # :50
    top_k_scores = torch.zeros(k, 1, dtype=dtype)                              # :52
    seqs_alpha = torch.ones(k, 1, enc_size, enc_size, dtype=dtype)             # :54
    complete_seqs, complete_alpha, complete_scores = [], [], []
    caption_end = False
    step = 1
    h, c = init_hidden_state(w, enc)                                           # :62
    while True:
        emb = F.embedding(k_prev_words, w["embedding.weight"]).squeeze(1)      # :65
        awe, alpha = soft_attention(w, enc, h)                                 # :66
        alpha = alpha.view(-1, enc_size, enc_size).unsqueeze(1)                # :67
        gate = torch.sigmoid(F.linear(h, w["f_beta.weight"], w["f_beta.bias"]))  # :68
        awe = gate * awe                                                       # :69
        x = torch.cat([emb.double(), awe.double()], dim=1).to(dtype)           # :70
        h, c = lstm_cell(x, h, c, w["decode_step.weight_ih"], w["decode_step.weight_hh"],
                         w["decode_step.bias_ih"], w["decode_step.bias_hh"])   # :71
        scores = F.linear(h, w["fc.weight"], w["fc.bias"])                     # :72 (no dropout)
        scores = F.log_softmax(scores, dim=1)                                  # :74
        scores = top_k_scores.expand_as(scores) + scores                       # :76
        if step == 1:
            top_k_scores, top_k_words = scores[0].topk(k, 0, True, True)       # :78-79
        else:
            top_k_scores, top_k_words = scores.view(-1).topk(k, 0, True, True)  # :82
        prev_word_inds = top_k_words // V                                      # :85
        next_word_inds = top_k_words % V                                       # :86
        seqs = torch.cat([seqs[prev_word_inds], next_word_inds.unsqueeze(1)], dim=1)         # :88
        seqs_alpha = torch.cat([seqs_alpha[prev_word_inds], alpha[prev_word_inds]], dim=1)   # :89
        if trace is not None:
            trace.append(next_word_inds.tolist())                              # :91 (print)
        incomplete = [i for i, nw in enumerate(next_word_inds.tolist()) if nw != end_id]     # :93-94
        complete = sorted(set(range(len(next_word_inds))) - set(incomplete))   # :96
        if len(complete) > 0:                                                  # :99-103
            caption_end = True
            complete_seqs.extend(seqs[complete].tolist())
            complete_alpha.extend(seqs_alpha[complete].tolist())
            complete_scores.extend(top_k_scores[complete])
        k -= len(complete)                                                     # :104
        if k == 0:                                                             # :107
            break
        seqs = seqs[incomplete]                                                # :109-116
        seqs_alpha = seqs_alpha[incomplete]
        enc = enc[prev_word_inds[incomplete]]
        top_k_scores = top_k_scores[incomplete].unsqueeze(1)
        h = h[prev_word_inds[incomplete]]
        c = c[prev_word_inds[incomplete]]
        k_prev_words = next_word_inds[incomplete].unsqueeze(1)
        if step > max_steps:                                                   # :119
            break
        step += 1
    if not caption_end:                                                        # :123-125
        return [start_id, end_id], [], caption_end
    idx = complete_scores.index(max(complete_scores))                          # :127 first max
    return complete_seqs[idx], complete_alpha[idx], caption_end


# --------------------------------------------------------------------------------------
# evaluate() — models/attention.py:516-553, the reference's batch-size-1 validation loop
# --------------------------------------------------------------------------------------
def evaluate_reference_style(w, encoder_out, captions, caption_lengths, special_ids):
    """One decoder call PER IMAGE like the reference's val_loader (batch_size=1, :489-494): -> (losses, hypotheses,
    references) exactly as the loop appends them (:533, :536-553)."""
    losses, hypotheses, references = [], [], []
    for j in range(encoder_out.size(0)):
        L = caption_lengths[j]
        caps = captions[j:j + 1, :L]                      # pad_sequence over a single caption adds no padding (:481-483)
        scores, caps_sorted, decode_lengths, alphas = attention_decoder_forward(w, encoder_out[j:j + 1], caps, [L])
        loss = attention_loss(scores, caps_sorted, decode_lengths, alphas)                          # :526-531
        losses.append(float(loss))
        targets = caps_sorted[:, 1:]
        img_captions = targets[0].tolist()
        cleaned = [x for x in img_captions if x not in special_ids]                                 # :539
        references.append([cleaned for _ in img_captions])                                          # :540
        _, preds = torch.max(scores, dim=2)                                                         # :544
        pred = preds.tolist()[0][:decode_lengths[0]]                                                # :548
        hypotheses.append([x for x in pred if x not in special_ids])                                # :550
    return losses, hypotheses, references
