"""GPU tier, module level: the drop-in modules (CUDA kernels through the C ABI) against
  (a) the golden vectors produced by the unmodified reference (tests/golden/*.npz), and
  (b) the oracle (oracle/decoders.py) run on the CPU on the same seeded inputs, fp32 and fp64.

Tolerances, fp32 tier (BASELINE.json north_star: 1e-3 relative on logits, alphas, loss, gradients):
  logits / alphas / loss          1e-4 norm-wise (measured ~1e-6; pure re-association of fp32 sums)
  gradients                       1e-3 norm-wise per tensor, with the SURVEY.md Appendix B caveat: the four
                                  attention-projection gradients are ill-conditioned (the reference itself is ~5e-4
                                  from fp64 truth), so they are judged by distance to the fp64 oracle as well
  attention.full_att.bias.grad    absolute 1e-5 (true value identically 0)
  greedy ids                      identical wherever the reference's top-2 logit margin exceeds 1e-4
"""
import os

import numpy as np
import pytest
import torch

import helpers as H
from oracle import decoders as O

pytestmark = pytest.mark.gpu

ILL_CONDITIONED = ("attention.enc_att.weight", "attention.enc_att.bias", "attention.dec_att.weight",
                   "attention.dec_att.bias")


def load(name):
    return np.load(os.path.join(H.GOLDEN_DIR, name + ".npz"), allow_pickle=False)


def golden_tensor(case, g, name):
    return torch.from_numpy(g[name]) if case["store"] == "full" else None


def compare(case, g, name, t, tol, atol=0.0):
    if case["store"] == "full":
        H.assert_close_norm(t, torch.from_numpy(g[name]), tol, name, atol)
    else:
        H.assert_digest_close(name, t, dict(norm=g[name + "@norm"], samples=g[name + "@samples"]), tol, atol)


def run_attention(case, cuda, dtype_oracle=None, precision="fp32"):
    import icd_b200.models.attention as my_att
    from icd_b200.vocabulary import synthetic_vocab
    dec = H.build_attention_module(case, my_att.AttentionDecoder, my_att.AttentionDecoderParams,
                                   synthetic_vocab(case["V"]))
    sd = {k: v.detach().clone() for k, v in dec.state_dict().items()}
    dec = dec.to(cuda)
    dec.precision = precision
    enc, caps, lens = H.att_inputs(case)
    dl = [l - 1 for l in lens]
    masks = None
    if case["train"]:
        masks = H.dropout_masks_like_reference(case, dl, case["D"])
        dec._dropout_mask_override = H.stack_masks(masks, case["B"], case["D"])
    caps_dev = caps.to(cuda)
    preds, caps_out, odl, alphas = dec(enc.to(cuda), caps_dev, lens)
    assert caps_out is caps_dev and odl == dl
    if case.get("loss", True):
        loss = O.attention_loss(preds, caps_dev, odl, alphas)       # the reference's loss glue, torch ops on GPU
    else:
        gen = torch.Generator().manual_seed(case["iseed"] + 1000)
        g1 = torch.randn(preds.shape, generator=gen).to(cuda)
        g2 = torch.randn(alphas.shape, generator=gen).to(cuda)
        loss = (preds * g1).sum() + (alphas * g2).sum()
    loss.backward()
    grads = {k: p.grad for k, p in dec.named_parameters()}
    return dec, sd, (enc, caps, lens, dl, masks), preds, alphas, loss, grads


@pytest.mark.parametrize("precision", ["fp32", "fp32x3"])
@pytest.mark.parametrize("name", list(H.ATT_CASES))
def test_attention_decoder_matches_reference_golden(cuda, name, precision):
    """Both fp32-class tiers against the same bars: "fp32" (fp32 FMA contractions) and "fp32x3" (3-term bf16 split on the
    tcgen05 tensor cores, ~6e-6 per contraction — inside BASELINE.json's 1e-3 bar for the fp32/TF32 path)."""
    case = H.ATT_CASES[name]
    g = load(name)
    dec, sd, (enc, caps, lens, dl, masks), preds, alphas, loss, grads = run_attention(case, cuda, precision=precision)
    cs = {k: H.checksum(v) for k, v in sd.items()}
    assert list(cs.values()) == [str(v) for v in g["weight_checksums"]], "seeded init differs from the reference"
    assert list(g["decode_lengths"]) == dl
    compare(case, g, "predictions", preds, 1e-4)
    compare(case, g, "alphas", alphas, 1e-4)
    # rows >= batch_size_t are exactly zero (models/attention.py:253-258)
    # (the first batch_size_t rows are active at step t, whatever their own length — :261-265)
    for t in range(max(dl)):
        bt = sum(l > t for l in dl)
        assert torch.all(preds[bt:, t] == 0) and torch.all(alphas[bt:, t] == 0)
    if case.get("loss", True):
        assert abs(loss.item() - float(g["loss"])) <= 1e-4 * abs(float(g["loss"]))
        ids = O.teacher_forced_ids(preds.cpu(), dl)
        gold, margin = g["greedy_ids"], g["greedy_margin"]
        for j, row in enumerate(ids):
            for t, tok in enumerate(row):
                if margin[j, t] > 1e-4:
                    assert tok == gold[j, t], "greedy id differs at (%d,%d) with margin %.2e" % (j, t, margin[j, t])
    # fp64 oracle as the arbiter for the ill-conditioned tensors
    frozen = () if case["fine_tune_embedding"] else ("embedding.weight",)
    w64 = {k: v.double().requires_grad_(k not in frozen) for k, v in sd.items()}
    p64, _, dl64, a64 = O.attention_decoder_forward(w64, enc.double(), caps, lens, dropout_p=case["dropout"],
                                                    dropout_masks=masks)
    if case.get("loss", True):
        l64 = O.attention_loss(p64, caps, dl64, a64)
    else:
        gen = torch.Generator().manual_seed(case["iseed"] + 1000)
        g1 = torch.randn(p64.shape, generator=gen).double()
        g2 = torch.randn(a64.shape, generator=gen).double()
        l64 = (p64 * g1).sum() + (a64 * g2).sum()
    l64.backward()
    names = [str(x) for x in g["grad_names"]]
    for k in names:
        gr = grads[k]
        assert gr is not None, k
        assert gr.dtype == sd[k].dtype, "gradient dtype of %s" % k
        if k == "attention.full_att.bias":
            assert float(gr.abs().max()) < 1e-5
            continue
        # the four attention-projection gradients are ill-conditioned (the reference's own fp32 result is ~5e-4 from the
        # fp64 truth): fp32 FMA tier 1e-3 / 2e-3, tensor-core fp32x3 tier 3e-3 on those four only
        ill = k in ILL_CONDITIONED
        H.assert_close_norm(gr, w64[k].grad, 3e-3 if (ill and precision == "fp32x3") else 1e-3, "grad vs fp64 oracle: " + k)
        tol = (3e-3 if precision == "fp32x3" else 2e-3) if ill else 1e-3
        compare(case, g, "grad:" + k, gr, tol)
    for k, gr in grads.items():
        if k not in names:
            assert gr is None, "unexpected gradient for " + k


def test_attention_decoder_eval_has_no_dropout_and_is_deterministic(cuda):
    import icd_b200.models.attention as my_att
    from icd_b200.vocabulary import synthetic_vocab
    case = dict(H.ATT_CASES["att_small_dropout"], train=False)
    dec = H.build_attention_module(case, my_att.AttentionDecoder, my_att.AttentionDecoderParams,
                                   synthetic_vocab(case["V"])).to(cuda)
    enc, caps, lens = H.att_inputs(case)
    with torch.no_grad():
        p1, _, _, a1 = dec(enc.to(cuda), caps.to(cuda), lens)
        p2, _, _, a2 = dec(enc.to(cuda), caps.to(cuda), lens)
    assert torch.equal(p1, p2) and torch.equal(a1, a2)
    w = O.cast_weights(dec.state_dict(), torch.float32)
    po, _, _, ao = O.attention_decoder_forward(w, enc, caps, lens)
    H.assert_close_norm(p1, po, 1e-4, "eval predictions")
    # train mode with the built-in Philox mask: seeded by torch's generator, drops ~p of fc inputs
    dec.train()
    torch.manual_seed(3)
    with torch.no_grad():
        t1, _, _, _ = dec(enc.to(cuda), caps.to(cuda), lens)
        torch.manual_seed(3)
        t2, _, _, _ = dec(enc.to(cuda), caps.to(cuda), lens)
        t3, _, _, _ = dec(enc.to(cuda), caps.to(cuda), lens)
    assert torch.equal(t1, t2) and not torch.equal(t1, t3) and not torch.equal(t1, p1)


def test_attention_decoder_accepts_permuted_encoder_view(cuda):
    """EncoderAttention returns an NCHW tensor permuted to (B,14,14,2048) (models/encoder.py:107-110)."""
    import icd_b200.models.attention as my_att
    from icd_b200.vocabulary import synthetic_vocab
    case = H.ATT_CASES["att_small_ragged"]
    dec = H.build_attention_module(case, my_att.AttentionDecoder, my_att.AttentionDecoderParams,
                                   synthetic_vocab(case["V"])).to(cuda)
    enc, caps, lens = H.att_inputs(case)
    nchw = enc.permute(0, 3, 1, 2).contiguous().to(cuda)
    view = nchw.permute(0, 2, 3, 1)
    assert not view.is_contiguous()
    with torch.no_grad():
        p1, _, _, a1 = dec(enc.to(cuda), caps.to(cuda), lens)
        p2, _, _, a2 = dec(view, caps.to(cuda), lens)
    assert torch.equal(p1, p2) and torch.equal(a1, a2)


def test_soft_attention_module_encoder_gradient(cuda):
    """SoftAttention used standalone (beam search / custom loops) also back-propagates into the features."""
    import icd_b200.models.attention as my_att
    torch.manual_seed(3)
    att = my_att.SoftAttention(2048, 32, 48)
    w = {"attention." + k: v.detach().clone().double() for k, v in att.state_dict().items()}
    att = att.to(cuda)
    g = torch.Generator().manual_seed(5)
    enc = torch.randn(3, 196, 2048, generator=g).clamp_min_(0)
    h = torch.randn(3, 32, generator=g)
    ga, gb = torch.randn(3, 2048, generator=g), torch.randn(3, 196, generator=g)
    enc_dev = enc.to(cuda).requires_grad_(True)
    awe, alpha = att(enc_dev, h.to(cuda))
    ((awe * ga.to(cuda)).sum() + (alpha * gb.to(cuda)).sum()).backward()
    e64 = enc.double().requires_grad_(True)
    awe64, alpha64 = O.soft_attention(w, e64, h.double())
    ((awe64 * ga.double()).sum() + (alpha64 * gb.double()).sum()).backward()
    H.assert_close_norm(enc_dev.grad, e64.grad, 1e-4, "SoftAttention d encoder_out")


def test_soft_attention_module_matches_oracle(cuda):
    import icd_b200.models.attention as my_att
    torch.manual_seed(0)
    att = my_att.SoftAttention(2048, 64, 48)
    w = {"attention." + k: v.detach().clone().double().requires_grad_(True) for k, v in att.state_dict().items()}
    att = att.to(cuda)
    g = torch.Generator().manual_seed(1)
    enc = torch.randn(4, 196, 2048, generator=g).clamp_min_(0)
    h = torch.randn(4, 64, generator=g)
    h_dev = h.to(cuda).requires_grad_(True)
    awe, alpha = att(enc.to(cuda), h_dev)
    h64 = h.double().requires_grad_(True)
    awe64, alpha64 = O.soft_attention(w, enc.double(), h64)
    H.assert_close_norm(awe, awe64, 1e-5, "awe")
    H.assert_close_norm(alpha, alpha64, 1e-5, "alpha")
    ga = torch.randn(awe.shape, generator=g)
    gb = torch.randn(alpha.shape, generator=g)
    ((awe * ga.to(cuda)).sum() + (alpha * gb.to(cuda)).sum()).backward()
    ((awe64 * ga.double()).sum() + (alpha64 * gb.double()).sum()).backward()
    H.assert_close_norm(h_dev.grad, h64.grad, 1e-4, "d hidden")
    for k, p in att.named_parameters():
        if k == "full_att.bias":
            assert float(p.grad.abs().max()) < 1e-4
            continue
        H.assert_close_norm(p.grad, w["attention." + k].grad, 1e-4, "grad " + k)


@pytest.mark.parametrize("precision", ["fp32", "fp32x3"])
@pytest.mark.parametrize("name", list(H.BASE_CASES))
def test_baseline_decoder_matches_reference_golden(cuda, name, precision):
    import icd_b200.models.baseline as my_base
    case = H.BASE_CASES[name]
    g = load(name)
    dec = H.build_baseline_module(case, my_base.BaselineDecoder, my_base.BaselineDecoderParams)
    cs = H.state_checksums(dec)
    assert list(cs.values()) == [str(v) for v in g["weight_checksums"]]
    dec = dec.to(cuda)
    dec.precision = precision
    img, caps, lens = H.base_inputs(case)
    img_dev = img.to(cuda).requires_grad_(True)
    outs = dec(img_dev, caps.to(cuda))
    assert outs.shape == (case["B"], case["L"], case["V"])
    compare(case, g, "outputs", outs, 1e-4)
    loss = O.baseline_loss(outs, caps.to(cuda))
    assert abs(loss.item() - float(g["loss"])) <= 1e-4 * abs(float(g["loss"]))
    loss.backward()
    for k in [str(x) for x in g["grad_names"]]:
        gr = img_dev.grad if k == "img_features" else dict(dec.named_parameters())[k].grad
        compare(case, g, "grad:" + k, gr, 1e-3)
    ids = outs.argmax(dim=2).cpu().numpy()
    margin = g["greedy_margin"]
    assert np.all((ids == g["greedy_ids"]) | (margin <= 1e-4))


@pytest.mark.parametrize("precision", ["fp32x3", "fp32"])
@pytest.mark.parametrize("name", list(H.BEAM_CASES))
def test_beam_search_matches_reference_golden(cuda, name, precision, capsys):
    import icd_b200.models.attention as my_att
    from icd_b200.gen_captions import attention_caption_image_beam_search, beam_search_batched
    from icd_b200.vocabulary import synthetic_vocab
    case = H.BEAM_CASES[name]
    g = load(name)
    vocab = synthetic_vocab(case["V"])
    acase = dict(case, dropout=0.5, train=False, fine_tune_embedding=True)
    dec = H.build_attention_module(acase, my_att.AttentionDecoder, my_att.AttentionDecoderParams, vocab)
    H.apply_beam_recipe(dec, case)
    assert list(H.state_checksums(dec).values()) == [str(v) for v in g["weight_checksums"]]
    dec = dec.to(cuda)
    feats = H.beam_features(case).to(cuda)
    V, k = case["V"], case["k"]
    with torch.no_grad():
        res = beam_search_batched(dec, feats, k, V - 3, V - 2, max_steps=50, want_alphas=True, want_trace=True,
                                  precision=precision)
    lens = res["len"].cpu().tolist()
    for i in range(case["n_img"]):
        if not bool(g["stable_%d" % i]):
            continue        # caption of the reference flips under fp64 re-evaluation: not a meaningful target
        gold = list(g["seq_%d" % i])
        if bool(g["ended_%d" % i]):
            assert lens[i] == len(gold)
            assert res["seq"][i, :lens[i]].cpu().tolist() == gold
            a = res["alpha"][i, :lens[i]].cpu().numpy().reshape(lens[i], 14, 14)
            assert np.abs(a - g["alphas_%d" % i]).max() < 1e-4
        else:
            assert lens[i] == 0
        gt = g["trace_%d" % i]
        tr = res["trace"][:gt.shape[0], i].cpu().numpy()
        assert np.array_equal(tr, gt), "per-step next-word trace differs for image %d" % i
        if gt.shape[0] < 51:
            assert np.all(res["trace"][gt.shape[0]:, i].cpu().numpy() == -1)

    # the drop-in single-image function: same signature / return tuple / per-step print as gen_captions.py:16-131
    class Args:
        beam_size = k

    class Identity(torch.nn.Module):
        def forward(self, x):
            return x
    i = [j for j in range(case["n_img"]) if bool(g["ended_%d" % j])][0]
    capsys.readouterr()
    seq, alphas, ended = attention_caption_image_beam_search(cuda, Args, feats[i:i + 1], Identity(), dec, vocab)
    printed = [ln for ln in capsys.readouterr().out.strip().split("\n") if ln]
    assert ended is True and seq == list(g["seq_%d" % i])
    gt = g["trace_%d" % i]
    assert printed == [str([vocab.i2w[int(x)] for x in row if x >= 0]) for row in gt]
    assert isinstance(alphas, list) and len(alphas) == len(seq) and len(alphas[0]) == 14 and len(alphas[0][0]) == 14
    assert alphas[0][0][0] == 1.0
    assert np.abs(np.asarray(alphas, dtype=np.float32) - g["alphas_%d" % i]).max() < 1e-4


def test_beam_search_failure_tuple_when_no_end(cuda):
    """Default random init never emits <end> within 51 steps => ([start, end], [], False) (gen_captions.py:123-125)."""
    import icd_b200.models.attention as my_att
    from icd_b200.gen_captions import attention_caption_image_beam_search
    from icd_b200.vocabulary import synthetic_vocab
    case = dict(H.BEAM_CASES["beam_small"], dropout=0.5, train=False, fine_tune_embedding=True)
    vocab = synthetic_vocab(case["V"])
    dec = H.build_attention_module(case, my_att.AttentionDecoder, my_att.AttentionDecoderParams, vocab)
    with torch.no_grad():
        dec.fc.bias[case["V"] - 2] = -50.0
    dec = dec.to(cuda)

    class Args:
        beam_size = 3

    class Identity(torch.nn.Module):
        def forward(self, x):
            return x
    feats = H.beam_features(case)[:1].to(cuda)
    seq, alphas, ended = attention_caption_image_beam_search(cuda, Args, feats, Identity(), dec, vocab)
    assert (seq, alphas, ended) == ([case["V"] - 3, case["V"] - 2], [], False)


def test_full_size_properties_config3_shape(cuda):
    """BASELINE config-3 shape (B=64 slice of the 512 batch to bound test time), properties that do not need the
    oracle: softmax rows sum to 1, masked rows are zero, batch rows are independent (a row's outputs do not depend on
    which other rows share the batch), the gradient of a sum over a split batch adds up."""
    import icd_b200.models.attention as my_att
    from icd_b200 import synthetic
    from icd_b200.vocabulary import synthetic_vocab
    V = 9490
    p = my_att.AttentionDecoderParams()
    p.vocab = synthetic_vocab(V)
    p.dropout = 0.0
    torch.manual_seed(0)
    dec = my_att.AttentionDecoder(cuda, p).to(cuda)
    dec.fine_tune_embeddings(False)
    B = 64
    enc = synthetic.features(B).to(cuda)
    caps, lens = synthetic.captions(B, V, max_len=25, lengths="ragged")
    caps = caps.to(cuda)
    preds, _, dl, alphas = dec(enc, caps, lens)
    assert preds.shape == (B, 24, V) and alphas.shape == (B, 24, 196)
    for b, l in enumerate(dl):
        assert torch.all(preds[b, l:] == 0) and torch.all(alphas[b, l:] == 0)
        assert torch.allclose(alphas[b, :l].sum(-1), torch.ones(l, device=cuda), atol=1e-5)
    loss = O.attention_loss(preds, caps, dl, alphas)
    loss.backward()
    g_full = {k: v.grad.clone() for k, v in dec.named_parameters() if v.grad is not None}
    # row independence: first 16 rows alone give the same logits
    with torch.no_grad():
        p16, _, dl16, a16 = dec(enc[:16], caps[:16], lens[:16])
    H.assert_close_norm(p16, preds[:16, :max(dl16)], 1e-5, "row independence (predictions)")
    H.assert_close_norm(a16, alphas[:16, :max(dl16)], 1e-5, "row independence (alphas)")
    # additivity over a batch split (what data parallelism relies on): sum-reduced losses
    dec.zero_grad()
    parts = []
    for sl in (slice(0, 32), slice(32, 64)):
        pp, _, dd, aa = dec(enc[sl], caps[sl], lens[sl])
        parts.append((pp * pp).sum() + (aa * aa).sum())
        parts[-1].backward()
    g_split = {k: v.grad.clone() for k, v in dec.named_parameters() if v.grad is not None}
    dec.zero_grad()
    pp, _, dd, aa = dec(enc, caps, lens)
    ((pp * pp).sum() + (aa * aa).sum()).backward()
    for k, v in dec.named_parameters():
        if v.grad is None or k == "attention.full_att.bias":
            continue
        H.assert_close_norm(g_split[k], v.grad, 1e-3, "split-batch additivity " + k)
    assert set(g_full) == set(g_split)


@pytest.mark.parametrize("name", ["att_small_ragged", "att_small_dropout", "att_cfg1"])
def test_attention_decoder_bf16_tier_matches_reference_golden(cuda, name):
    """bf16 tensor-core tier (tcgen05 GEMMs, fp32 everything else).  Stated tolerances (measured
    bf16-operand emulation; the small-dimension goldens are noisier than the full-size case): logits / alphas 5e-3,
    loss 1e-3, gradients 1e-2, the four ill-conditioned attention-projection gradients 1.5e-1; greedy ids must agree wherever the reference's top-2 margin exceeds 2e-2."""
    case = dict(H.ATT_CASES[name])
    g = load(name)
    import icd_b200.models.attention as my_att
    from icd_b200.vocabulary import synthetic_vocab
    dec = H.build_attention_module(case, my_att.AttentionDecoder, my_att.AttentionDecoderParams,
                                   synthetic_vocab(case["V"])).to(cuda)
    dec.precision = "bf16"
    enc, caps, lens = H.att_inputs(case)
    dl = [l - 1 for l in lens]
    if case["train"]:
        masks = H.dropout_masks_like_reference(case, dl, case["D"])
        dec._dropout_mask_override = H.stack_masks(masks, case["B"], case["D"])
    caps_dev = caps.to(cuda)
    preds, _, odl, alphas = dec(enc.to(cuda), caps_dev, lens)
    loss = O.attention_loss(preds, caps_dev, odl, alphas)
    loss.backward()
    compare(case, g, "predictions", preds, 5e-3)
    compare(case, g, "alphas", alphas, 5e-3)
    assert abs(loss.item() - float(g["loss"])) <= 1e-3 * abs(float(g["loss"]))
    for t in range(max(dl)):
        bt = sum(l > t for l in dl)
        assert torch.all(preds[bt:, t] == 0) and torch.all(alphas[bt:, t] == 0)
    ids = O.teacher_forced_ids(preds.cpu(), dl)
    gold, margin = g["greedy_ids"], g["greedy_margin"]
    for j, row in enumerate(ids):
        for t, tok in enumerate(row):
            if margin[j, t] > 2e-2:
                assert tok == gold[j, t]
    grads = {k: p.grad for k, p in dec.named_parameters()}
    for k in [str(x) for x in g["grad_names"]]:
        if k == "attention.full_att.bias":
            assert float(grads[k].abs().max()) < 1e-4
            continue
        compare(case, g, "grad:" + k, grads[k], 1.5e-1 if k in ILL_CONDITIONED else 1e-2)


def test_bf16_tier_accepts_bf16_stored_features_in_place(cuda):
    """Features handed over in bf16 (encoder under autocast / bf16 feature store) are consumed without an fp32 round
    trip; the result must equal the fp32-input path fed the same (bf16-representable) values."""
    import icd_b200.models.attention as my_att
    from icd_b200.vocabulary import synthetic_vocab
    case = dict(H.ATT_CASES["att_small_ragged"])
    dec = H.build_attention_module(case, my_att.AttentionDecoder, my_att.AttentionDecoderParams,
                                   synthetic_vocab(case["V"])).to(cuda)
    dec.precision = "bf16"
    enc, caps, lens = H.att_inputs(case)
    enc16 = enc.bfloat16()
    outs = []
    for feats in (enc16.float().to(cuda), enc16.to(cuda)):
        dec.zero_grad()
        preds, _, dl, alphas = dec(feats, caps.to(cuda), lens)
        loss = O.attention_loss(preds, caps.to(cuda), dl, alphas)
        loss.backward()
        outs.append((preds.detach().clone(), alphas.detach().clone(),
                     {k: p.grad.detach().clone() for k, p in dec.named_parameters()}))
    H.assert_close_norm(outs[1][0], outs[0][0], 1e-5, "predictions (bf16-stored features)")
    H.assert_close_norm(outs[1][1], outs[0][1], 1e-5, "alphas (bf16-stored features)")
    for k in outs[0][2]:
        H.assert_close_norm(outs[1][2][k], outs[0][2][k], 1e-4, "grad " + k, atol=1e-7)


@pytest.mark.parametrize("name", list(H.BASE_CASES))
def test_baseline_decoder_bf16_tier_matches_reference_golden(cuda, name):
    """bf16 tensor-core tier of the baseline decoder (BASELINE.json configs[1] on B200): stated tolerances — logits 5e-3,
    loss 1e-3, gradients 2e-2 norm-wise; greedy ids equal wherever the reference's top-2 margin exceeds 2e-2."""
    import icd_b200.models.baseline as my_base
    case = H.BASE_CASES[name]
    g = load(name)
    dec = H.build_baseline_module(case, my_base.BaselineDecoder, my_base.BaselineDecoderParams).to(cuda)
    dec.precision = "bf16"
    img, caps, lens = H.base_inputs(case)
    img_dev = img.to(cuda).requires_grad_(True)
    outs = dec(img_dev, caps.to(cuda))
    compare(case, g, "outputs", outs, 5e-3)
    loss = O.baseline_loss(outs, caps.to(cuda))
    assert abs(loss.item() - float(g["loss"])) <= 1e-3 * abs(float(g["loss"]))
    loss.backward()
    for k in [str(x) for x in g["grad_names"]]:
        gr = img_dev.grad if k == "img_features" else dict(dec.named_parameters())[k].grad
        compare(case, g, "grad:" + k, gr, 2e-2)
    ids = outs.argmax(dim=2).cpu().numpy()
    assert np.all((ids == g["greedy_ids"]) | (g["greedy_margin"] <= 2e-2))


@pytest.mark.parametrize("precision,tol", [("fp32", 1e-3), ("bf16", 3e-2)])
def test_encoder_feature_gradient_matches_oracle(cuda, precision, tol):
    """encoder_out.requires_grad (the reference reaches this with --fine_tune_encoder, train.py:39): the gradient w.r.t.
    the features — attention-weighted sum (models/attention.py:59-60), enc_att projection (:54) and initial-state mean
    (:161) paths — against autograd over the fp64 oracle, for a ragged batch fed as the encoder's permuted NCHW view."""
    import icd_b200.models.attention as my_att
    from icd_b200.vocabulary import synthetic_vocab
    case = dict(H.ATT_CASES["att_small_ragged"])
    dec = H.build_attention_module(case, my_att.AttentionDecoder, my_att.AttentionDecoderParams,
                                   synthetic_vocab(case["V"]))
    sd = {k: v.detach().clone() for k, v in dec.state_dict().items()}
    dec = dec.to(cuda)
    dec.precision = precision
    enc, caps, lens = H.att_inputs(case)
    nchw = enc.permute(0, 3, 1, 2).contiguous().to(cuda).requires_grad_(True)       # what ResNet produces
    preds, _, dl, alphas = dec(nchw.permute(0, 2, 3, 1), caps.to(cuda), lens)       # models/encoder.py:109 view
    loss = O.attention_loss(preds, caps.to(cuda), dl, alphas)
    loss.backward()
    w64 = {k: v.double() for k, v in sd.items()}
    e64 = enc.double().requires_grad_(True)
    p64, _, dl64, a64 = O.attention_decoder_forward(w64, e64, caps, lens)
    O.attention_loss(p64, caps, dl64, a64).backward()
    assert nchw.grad is not None and nchw.grad.shape == nchw.shape and nchw.grad.dtype == torch.float32
    H.assert_close_norm(nchw.grad.permute(0, 2, 3, 1), e64.grad, tol, "d loss / d encoder_out (%s)" % precision)
    # rows are independent: a caption's feature gradient only depends on its own row
    assert float(nchw.grad[0].abs().max()) > 0


@pytest.mark.parametrize("precision,tol", [("fp32", 1e-4), ("bf16", 8e-3)])
@pytest.mark.parametrize("B,lengths", [(1, [2]), (1, [9]), (3, [6, 1, 4]), (4, [2, 2, 2, 2]), (2, [40, 37])])
def test_attention_decoder_edge_shapes(cuda, precision, tol, B, lengths):
    """Edges of the teacher-forced loop against the oracle: a single caption, a single decode step (length-2 captions),
    a caption of length 1 (zero decode steps: its rows stay exactly 0), long captions (T = 39 steps)."""
    import icd_b200.models.attention as my_att
    from icd_b200.vocabulary import synthetic_vocab
    case = dict(B=B, V=61, A=48, D=32, E=24, max_len=max(lengths), lengths=lengths, wseed=9, iseed=31 + B,
                dropout=0.0, train=False, fine_tune_embedding=True)
    dec = H.build_attention_module(case, my_att.AttentionDecoder, my_att.AttentionDecoderParams,
                                   synthetic_vocab(case["V"]))
    w = {k: v.detach().clone().requires_grad_(True) for k, v in dec.state_dict().items()}
    dec = dec.to(cuda)
    dec.precision = precision
    enc = synthetic_features(B, case["iseed"])
    caps, lens = synthetic_caps(case)
    preds, _, dl, alphas = dec(enc.to(cuda), caps.to(cuda), lens)
    o_preds, _, o_dl, o_alphas = O.attention_decoder_forward(w, enc, caps, lens)
    assert dl == o_dl and preds.shape == o_preds.shape and alphas.shape == o_alphas.shape
    H.assert_close_norm(preds, o_preds, tol, "predictions")
    H.assert_close_norm(alphas, o_alphas, tol, "alphas")
    for t in range(max(dl)):                      # rows beyond batch_size_t stay exactly 0 (models/attention.py:253-258)
        bt = sum(l > t for l in dl)
        assert torch.all(preds[bt:, t] == 0) and torch.all(alphas[bt:, t] == 0)
    g = torch.Generator().manual_seed(5)
    gp, ga = torch.randn(preds.shape, generator=g), torch.randn(alphas.shape, generator=g)
    ((preds * gp.to(cuda)).sum() + (alphas * ga.to(cuda)).sum()).backward()
    ((o_preds * gp).sum() + (o_alphas * ga).sum()).backward()
    gtol = 2e-3 if precision == "fp32" else 1.5e-1
    for k, p in dec.named_parameters():
        if k == "attention.full_att.bias":
            continue
        H.assert_close_norm(p.grad, w[k].grad, gtol, "grad " + k, atol=1e-6)


def synthetic_features(B, seed):
    from icd_b200 import synthetic
    return synthetic.features(B, seed=seed)


def synthetic_caps(case):
    from icd_b200 import synthetic
    return synthetic.captions(case["B"], case["V"], max_len=case["max_len"], seed=case["iseed"], lengths=case["lengths"])


@pytest.mark.parametrize("k,n_img", [(1, 3), (8, 2), (5, 1)])
def test_beam_search_edge_beam_widths_match_oracle(cuda, k, n_img):
    """Beam widths 1 and 8 (the kernel's maximum) and a single image, against the oracle's per-image state machine."""
    import icd_b200.models.attention as my_att
    from icd_b200.gen_captions import beam_search_batched
    from icd_b200.vocabulary import synthetic_vocab
    case = dict(H.BEAM_CASES["beam_small"], dropout=0.5, train=False, fine_tune_embedding=True, k=k, n_img=n_img)
    vocab = synthetic_vocab(case["V"])
    dec = H.build_attention_module(case, my_att.AttentionDecoder, my_att.AttentionDecoderParams, vocab)
    H.apply_beam_recipe(dec, case)
    w = {kk: v.detach().clone() for kk, v in dec.state_dict().items()}
    w64 = {kk: v.double() for kk, v in w.items()}
    dec = dec.to(cuda)
    feats = H.beam_features(case)[:n_img]
    V = case["V"]
    with torch.no_grad():
        res = beam_search_batched(dec, feats.to(cuda), k, V - 3, V - 2, max_steps=50, want_alphas=True, want_trace=True)
    lens = res["len"].cpu().tolist()
    checked = 0
    for i in range(n_img):
        tr32, tr64 = [], []
        seq, alphas, ended = O.beam_search(w, feats[i:i + 1], k, V - 3, V - 2, trace=tr32)
        seq64, _, ended64 = O.beam_search(w64, feats[i:i + 1].double(), k, V - 3, V - 2, trace=tr64)
        if (seq, ended) != (seq64, ended64) or tr32 != tr64:
            continue            # the reference's own outcome is not stable under fp64 re-evaluation: not a target
        checked += 1
        if ended:
            assert lens[i] == len(seq) and res["seq"][i, :lens[i]].cpu().tolist() == seq
            a = res["alpha"][i, :lens[i]].cpu().numpy().reshape(lens[i], 14, 14)
            assert np.abs(a - np.asarray(alphas, dtype=np.float32)).max() < 1e-4
        else:
            assert lens[i] == 0
        got = [[x for x in row if x >= 0] for row in res["trace"][:len(tr32), i].cpu().tolist()]
        assert got == tr32
    assert checked >= 1


@pytest.mark.gpu
@pytest.mark.parametrize("precision", ["fp32", "fp32x3"])
def test_beam_search_batch_equals_single_images_under_compaction(cuda, precision):
    """Finished beams / images leave the working set (slot and row compaction on the device).  An image's caption,
    score, alpha frames and per-step candidate words must not depend on which other images share the batch: the batched
    call equals one call per image, for a mix of images that finish at different steps (<end> bias varied per run)."""
    import icd_b200.models.attention as my_att
    from icd_b200.gen_captions import beam_search_batched
    from icd_b200.vocabulary import synthetic_vocab
    case = dict(H.BEAM_CASES["beam_small"], dropout=0.5, train=False, fine_tune_embedding=True, k=4)
    vocab = synthetic_vocab(case["V"])
    dec = H.build_attention_module(case, my_att.AttentionDecoder, my_att.AttentionDecoderParams, vocab)
    H.apply_beam_recipe(dec, case)
    dec = dec.to(cuda)
    V, k = case["V"], 4
    base = H.beam_features(case)
    g = torch.Generator().manual_seed(11)
    feats = torch.cat([base, base.flip(0) * 0.7 + 0.1, torch.rand(base.shape, generator=g) * base.max()], dim=0).to(cuda)
    n = feats.shape[0]
    with torch.no_grad():
        full = beam_search_batched(dec, feats, k, V - 3, V - 2, max_steps=30, want_alphas=True, want_trace=True,
                                   precision=precision)
        lens = full["len"].cpu().tolist()
        assert len(set(lens)) > 1, "the test needs images that finish at different steps"
        for i in range(n):
            one = beam_search_batched(dec, feats[i:i + 1], k, V - 3, V - 2, max_steps=30, want_alphas=True,
                                      want_trace=True, precision=precision)
            assert one["len"].cpu().tolist()[0] == lens[i]
            L = lens[i]
            assert torch.equal(one["seq"][0, :L], full["seq"][i, :L])
            assert torch.equal(one["score"][0], full["score"][i])
            assert torch.equal(one["alpha"][0, :L], full["alpha"][i, :L])
            assert torch.equal(one["trace"][:, 0], full["trace"][:, i])


@pytest.mark.gpu
@pytest.mark.parametrize("precision,tol,gtol", [("fp32", 1e-4, 2e-3), ("fp32x3", 1e-4, 3e-3), ("bf16", 8e-3, 1.5e-1)])
def test_attention_decoder_real_vocabulary_size_odd_v(cuda, precision, tol, gtol):
    """V = 8111 (the reference's real COCO vocabulary): an ODD row length, so the logits rows take every
    4-byte alignment and the vocabulary-layer kernels run their scalar / shifted paths, at the full model dimensions."""
    import icd_b200.models.attention as my_att
    from icd_b200.vocabulary import synthetic_vocab
    case = dict(B=3, V=8111, A=512, D=512, E=512, max_len=6, lengths=[6, 5, 3], wseed=5, iseed=77,
                dropout=0.0, train=False, fine_tune_embedding=True)
    dec = H.build_attention_module(case, my_att.AttentionDecoder, my_att.AttentionDecoderParams,
                                   synthetic_vocab(case["V"]))
    w = {k: v.detach().clone().requires_grad_(True) for k, v in dec.state_dict().items()}
    dec = dec.to(cuda)
    dec.precision = precision
    enc = synthetic_features(case["B"], case["iseed"])
    caps, lens = synthetic_caps(case)
    preds, cs, dl, alphas = dec(enc.to(cuda), caps.to(cuda), lens)
    o_preds, _, o_dl, o_alphas = O.attention_decoder_forward(w, enc, caps, lens)
    assert dl == o_dl
    H.assert_close_norm(preds, o_preds, tol, "predictions")
    H.assert_close_norm(alphas, o_alphas, tol, "alphas")
    from icd_b200.losses import attention_caption_loss
    loss = attention_caption_loss(preds, cs, dl, alphas)
    o_loss = O.attention_loss(o_preds, caps, o_dl, o_alphas)
    assert abs(loss.item() - o_loss.item()) < max(tol, 1e-5) * abs(o_loss.item())
    loss.backward()
    o_loss.backward()
    for k, p in dec.named_parameters():
        if k == "attention.full_att.bias":
            continue
        H.assert_close_norm(p.grad, w[k].grad, gtol, "grad " + k, atol=1e-6)


@pytest.mark.gpu
def test_bf16_grad_only_loss_is_identical_and_refuses_misuse(cuda):
    """attention_caption_loss(bf16_grad_only=True) skips the fp32 copy of d(loss)/d(logits): the decoder gradients are
    bit-identical (the bf16 tier consumes the bf16 copy either way), and a graph in which ``predictions`` also feeds
    something else is refused with an error instead of silently reading the hollow fp32 tensor."""
    import icd_b200.models.attention as my_att
    from icd_b200.losses import attention_caption_loss
    from icd_b200.vocabulary import synthetic_vocab
    case = dict(B=6, V=310, A=64, D=64, E=64, max_len=9, lengths=[9, 8, 8, 5, 3, 2], wseed=2, iseed=41,
                dropout=0.0, train=False, fine_tune_embedding=True)
    dec = H.build_attention_module(case, my_att.AttentionDecoder, my_att.AttentionDecoderParams,
                                   synthetic_vocab(case["V"])).to(cuda)
    dec.precision = "bf16"
    enc = synthetic_features(case["B"], case["iseed"]).to(cuda)
    caps, lens = synthetic_caps(case)
    caps = caps.to(cuda)
    grads = []
    for only in (False, True):
        dec.zero_grad()
        preds, cs, dl, alphas = dec(enc, caps, lens)
        attention_caption_loss(preds, cs, dl, alphas, bf16_grad_only=only).backward()
        grads.append({k: p.grad.clone() for k, p in dec.named_parameters()})
    for k in grads[0]:
        if k == "embedding.weight":           # scatter-add with atomics: run-to-run summation order differs
            H.assert_close_norm(grads[1][k], grads[0][k], 1e-5, "grad " + k)
        else:
            assert torch.equal(grads[0][k], grads[1][k]), k
    preds, cs, dl, alphas = dec(enc, caps, lens)
    loss = attention_caption_loss(preds, cs, dl, alphas, bf16_grad_only=True) + 1e-3 * preds.sum()
    with pytest.raises(RuntimeError, match="bf16_grad_only"):
        loss.backward()



@pytest.mark.gpu
@pytest.mark.parametrize("precision,tol,gtol", [("fp32", 1e-4, 2e-3), ("bf16", 8e-3, 1.5e-1)])
def test_precomputed_caption_embeddings_use_bert_branch(cuda, precision, tol, gtol):
    """use_bert=True (models/attention.py:96-100, 242-244): the decoder takes (B, L, 768) pre-computed caption embeddings from
    ``bert_embedder`` (or ``forward(embeddings=...)``) and runs the same kernels with E = 768; the embedding table is bypassed
    and receives no gradient.  Against the fp64 oracle's ``embeddings=`` path (pinned to the reference's branch on the CPU)."""
    import icd_b200.models.attention as my_att
    from icd_b200.vocabulary import synthetic_vocab
    V, B, L = 131, 5, 9
    p = my_att.AttentionDecoderParams()
    p.attention_dim, p.decoder_dim, p.embed_size, p.dropout, p.use_bert = 48, 32, 768, 0.0, True
    p.vocab = synthetic_vocab(V)
    torch.manual_seed(7)
    dec = my_att.AttentionDecoder(cuda, p)
    w64 = {k: v.detach().clone().double().requires_grad_(True) for k, v in dec.state_dict().items()}
    dec = dec.to(cuda)
    dec.precision = precision
    g = torch.Generator().manual_seed(8)
    emb = torch.randn(B, L, 768, generator=g) * 0.3
    enc = synthetic_features(B, 17)
    caps, lens = synthetic_caps(dict(B=B, V=V, max_len=L, iseed=17, lengths=[9, 9, 6, 4, 2]))
    with pytest.raises(NotImplementedError, match="bert_embedder"):
        dec(enc.to(cuda), caps.to(cuda), lens)
    calls = []
    dec.bert_embedder = lambda c: (calls.append(c.shape), emb.to(cuda))[1]
    preds, _, dl, alphas = dec(enc.to(cuda), caps.to(cuda), lens)
    assert calls == [caps.shape]
    o_preds, _, o_dl, o_alphas = O.attention_decoder_forward(w64, enc.double(), caps, lens, embeddings=emb.double())
    assert dl == o_dl
    H.assert_close_norm(preds, o_preds, tol, "predictions")
    H.assert_close_norm(alphas, o_alphas, tol, "alphas")
    O.attention_loss(preds, caps.to(cuda), dl, alphas).backward()
    O.attention_loss(o_preds, caps, o_dl, o_alphas).backward()
    for k, q in dec.named_parameters():
        if k == "embedding.weight":
            assert q.grad is None
        elif k != "attention.full_att.bias":
            H.assert_close_norm(q.grad, w64[k].grad, gtol, "grad " + k, atol=1e-7)
    # the explicit keyword gives the same result as the embedder hook
    with torch.no_grad():
        p2, _, _, a2 = dec(enc.to(cuda), caps.to(cuda), lens, embeddings=emb.to(cuda))
    assert torch.equal(p2, preds) and torch.equal(a2, alphas)


@pytest.mark.gpu
def test_batched_evaluate_equals_reference_per_image_loop(cuda):
    """icd_b200.evaluation (batched port of evaluate(), models/attention.py:454-567): one call over a ragged batch returns the
    per-image losses, hypotheses and references the reference's batch-size-1 loop produces (oracle restatement of :516-553)."""
    import icd_b200.models.attention as my_att
    from icd_b200.evaluation import evaluate, evaluate_batch
    from icd_b200.vocabulary import END_TOKEN, PAD_TOKEN, START_TOKEN, synthetic_vocab
    case = dict(H.ATT_CASES["att_small_ragged"])
    vocab = synthetic_vocab(case["V"])
    dec = H.build_attention_module(case, my_att.AttentionDecoder, my_att.AttentionDecoderParams, vocab)
    w = O.cast_weights(dec.state_dict(), torch.float64)
    dec = dec.to(cuda).train()                      # evaluate() must switch it to eval itself (:503)
    enc, caps, lens = H.att_inputs(case)
    special = [vocab(START_TOKEN), vocab(END_TOKEN), vocab(PAD_TOKEN)]
    o_losses, o_hyp, o_ref = O.evaluate_reference_style(w, enc.double(), caps, lens, special)
    dec.eval()
    losses, hyp, ref, n = evaluate_batch(dec, enc.to(cuda), caps.to(cuda), lens, vocab)
    assert n == [l - 1 for l in lens] and ref == o_ref
    for a, b in zip(losses, o_losses):
        assert abs(a - b) < 1e-4 * abs(b)
    # hypotheses: identical wherever the oracle's own top-2 margin is not razor thin
    p64, _, dl64, _ = O.attention_decoder_forward(w, enc.double(), caps, lens)
    top2 = p64.topk(2, dim=2).values
    margin_ok = [(top2[j, :dl64[j], 0] - top2[j, :dl64[j], 1]).min().item() > 1e-4 for j in range(case["B"])]
    assert any(margin_ok)
    for j in range(case["B"]):
        if margin_ok[j]:
            assert hyp[j] == o_hyp[j]

    class Identity(torch.nn.Module):
        def forward(self, x):
            return x

    class Args:
        print_freq = 1
    dec.train()
    loader = [(enc[:3], caps[:3], lens[:3]), (enc[3:], caps[3:, :lens[3]], lens[3:])]
    m = evaluate(cuda, Args, Identity(), dec, loader, vocab, score_fn=lambda r, h: {"n": len(h)})
    assert not dec.training and m["n"] == case["B"] and len(m["losses"]) == case["B"]
    for a, b in zip(m["losses"], o_losses):
        assert abs(a - b) < 1e-4 * abs(b)
    tok = [l - 1 for l in lens]
    assert abs(m["avg_loss"] - sum(a * b for a, b in zip(o_losses, tok)) / sum(tok)) < 1e-4



@pytest.mark.gpu
@pytest.mark.parametrize("glove", [False, True])
def test_embedding_gradient_is_run_to_run_deterministic(cuda, glove):
    """The embedding-table gradient (models/attention.py:247 backward; fp64 for the GloVe table) is a scatter-add over
    repeated token ids: rows are sorted by (token, row) and each token's rows are added in row order by one CTA, so two runs
    give bit-identical gradients (atomics would not), and the values match the fp64 oracle."""
    import icd_b200.models.attention as my_att
    from icd_b200.vocabulary import synthetic_vocab
    V = 23                                                   # few tokens: every id repeats many times
    case = dict(B=40, V=V, A=48, D=32, E=(300 if glove else 24), max_len=12, lengths=[12] * 40, wseed=4, iseed=9,
                dropout=0.0, train=False, fine_tune_embedding=True, glove=glove)
    dec = H.build_attention_module(case, my_att.AttentionDecoder, my_att.AttentionDecoderParams, synthetic_vocab(V))
    w64 = {k: v.detach().clone().double().requires_grad_(True) for k, v in dec.state_dict().items()}
    dec = dec.to(cuda)
    enc = synthetic_features(case["B"], case["iseed"])
    caps, lens = synthetic_caps(case)
    grads = []
    for _ in range(3):
        dec.zero_grad()
        preds, _, dl, alphas = dec(enc.to(cuda), caps.to(cuda), lens)
        O.attention_loss(preds, caps.to(cuda), dl, alphas).backward()
        grads.append(dec.embedding.weight.grad.clone())
    assert grads[0].dtype == (torch.float64 if glove else torch.float32)
    assert torch.equal(grads[0], grads[1]) and torch.equal(grads[0], grads[2])
    p64, _, dl64, a64 = O.attention_decoder_forward(w64, enc.double(), caps, lens)
    O.attention_loss(p64, caps, dl64, a64).backward()
    H.assert_close_norm(grads[0], w64["embedding.weight"].grad, 1e-4, "embedding gradient")


@pytest.mark.gpu
@pytest.mark.parametrize("precision", ["fp32x3", "bf16"])
def test_lstm_gate_adjoint_vectorised_kernel_is_bit_identical_to_scalar(cuda, precision, monkeypatch):
    """(The tensor-core tiers: their contractions are run-to-run deterministic; the fp32 FMA tier's split-K weight gradients use atomics.)
    The LSTMCell adjoint of the BPTT (models/attention.py:277-278 backward) runs four hidden units per thread; the
    one-unit-per-thread kernel (ICD_LSTM_BWD_SCALAR, also the fallback for D % 4 != 0) must give bit-identical gradients,
    with train-mode dropout, ragged lengths and the deferred split-K planes of the dh contraction in play."""
    import icd_b200.models.attention as my_att
    from icd_b200.vocabulary import synthetic_vocab
    case = dict(B=37, V=97, A=64, D=64, E=32, max_len=11, lengths=[11] * 9 + [9] * 8 + [6] * 10 + [3] * 10, wseed=6, iseed=41,
                dropout=0.5, train=True, fine_tune_embedding=True)
    dec = H.build_attention_module(case, my_att.AttentionDecoder, my_att.AttentionDecoderParams, synthetic_vocab(case["V"]))
    dec = dec.to(cuda)
    dec.train()
    dec.precision = precision
    enc = synthetic_features(case["B"], case["iseed"]).to(cuda)
    caps, lens = synthetic_caps(case)
    caps = caps.to(cuda)
    keep = (torch.rand(max(lens) - 1, case["B"], case["D"], generator=torch.Generator().manual_seed(3)) < 0.5).to(cuda)

    def grads():
        dec.zero_grad()
        dec._dropout_mask_override = keep
        preds, _, dl, alphas = dec(enc, caps, lens)
        O.attention_loss(preds, caps, dl, alphas).backward()
        return {k: p.grad.clone() for k, p in dec.named_parameters() if p.grad is not None}

    monkeypatch.delenv("ICD_LSTM_BWD_SCALAR", raising=False)
    g_vec = grads()
    monkeypatch.setenv("ICD_LSTM_BWD_SCALAR", "1")
    g_sc = grads()
    monkeypatch.delenv("ICD_LSTM_BWD_SCALAR", raising=False)
    assert g_vec.keys() == g_sc.keys() and len(g_vec) > 8
    for k in g_vec:
        assert torch.equal(g_vec[k], g_sc[k]), k


@pytest.mark.gpu
def test_in_place_weight_update_between_forward_and_backward_raises(cuda):
    """The decoders' autograd Functions read the weights through raw pointers in the backward; like autograd's saved-tensor check
    they must refuse a backward when a parameter was modified in place after the forward (the reference, built from nn modules,
    raises "modified by an inplace operation" in the same situation) — and run normally when nothing was touched."""
    import icd_b200.models.attention as my_att
    import icd_b200.models.baseline as my_base
    from icd_b200.vocabulary import synthetic_vocab
    case = H.ATT_CASES["att_small_ragged"]
    dec = H.build_attention_module(case, my_att.AttentionDecoder, my_att.AttentionDecoderParams, synthetic_vocab(case["V"])).to(cuda)
    enc, caps, lens = H.att_inputs(case)
    preds, cs, dl, alphas = dec(enc.to(cuda), caps.to(cuda), lens)
    loss = O.attention_loss(preds, caps.to(cuda), dl, alphas)
    with torch.no_grad():
        dec.fc.weight.mul_(1.0)
    with pytest.raises(RuntimeError, match="modified by an in-place operation"):
        loss.backward()
    dec.zero_grad()
    preds, cs, dl, alphas = dec(enc.to(cuda), caps.to(cuda), lens)
    O.attention_loss(preds, caps.to(cuda), dl, alphas).backward()
    assert dec.fc.weight.grad is not None and torch.isfinite(dec.fc.weight.grad).all()
    p = my_base.BaselineDecoderParams()
    p.vocab_size, p.embed_size, p.hidden_size = 50, 16, 16
    torch.manual_seed(0)
    base = my_base.BaselineDecoder(p).to(cuda)
    img = torch.randn(3, 16, device=cuda)
    bcaps = torch.randint(1, 47, (3, 7), device=cuda)
    out = base(img, bcaps)
    with torch.no_grad():
        base.lstm.weight_hh_l0.add_(0.0)
    with pytest.raises(RuntimeError, match="modified by an in-place operation"):
        out.sum().backward()


@pytest.mark.gpu
@pytest.mark.parametrize("ties", ["none", "groups", "all_equal"])
def test_beam_topk_filtered_pass_equals_streaming_pass(cuda, ties, monkeypatch):
    """The top-k kernel of the beam search keeps a row in registers and offers only the logits >= tau (a lower bound of the row's
    k-th largest logit) to the candidate lists; rows it cannot take (odd V, V > 9728, more than 512 logits >= tau) go through the
    streaming pass.  Both must select the same candidates in the same order — also under exact ties (lower flat index first,
    gen_captions.py:78-82): "groups" gives every logit one of ~9 values (hundreds of exact ties at the top), "all_equal" makes
    every logit of a row identical (the candidate list overflows and the row falls back to the streaming pass)."""
    import icd_b200.models.attention as my_att
    from icd_b200.gen_captions import beam_search_batched
    from icd_b200.vocabulary import synthetic_vocab
    case = dict(H.BEAM_CASES["beam_small"], V=1000, dropout=0.5, train=False, fine_tune_embedding=True, k=5)
    vocab = synthetic_vocab(case["V"])
    dec = H.build_attention_module(case, my_att.AttentionDecoder, my_att.AttentionDecoderParams, vocab)
    H.apply_beam_recipe(dec, case)
    V = case["V"]
    with torch.no_grad():
        if ties != "none":
            dec.fc.weight.zero_()
            if ties == "groups":
                dec.fc.bias.copy_((torch.randn(V, generator=torch.Generator().manual_seed(3)) * 2).round() / 2)
            else:
                dec.fc.bias.fill_(0.25)
    dec = dec.to(cuda)
    feats = H.beam_features(case).to(cuda)
    outs = []
    for stream in (False, True):
        if stream:
            monkeypatch.setenv("ICD_BEAM_TOPK_STREAM", "1")
        else:
            monkeypatch.delenv("ICD_BEAM_TOPK_STREAM", raising=False)
        with torch.no_grad():
            outs.append(beam_search_batched(dec, feats, 5, V - 3, V - 2, max_steps=12, want_alphas=False, want_trace=True,
                                            precision="fp32x3"))
    monkeypatch.delenv("ICD_BEAM_TOPK_STREAM", raising=False)
    a, b = outs
    assert torch.equal(a["len"], b["len"]) and torch.equal(a["seq"], b["seq"]) and torch.equal(a["trace"], b["trace"])
    assert torch.equal(a["score"], b["score"])


@pytest.mark.gpu
@pytest.mark.parametrize("k,precision", [(1, "fp32x3"), (3, "fp32"), (5, "fp32x3"), (8, "fp32x3")])
def test_beam_grouped_attention_ring_equals_register_staged(cuda, k, precision, monkeypatch):
    """The attention step of caption generation runs as persistent CTAs that stream att_enc / enc through a bulk-async
    shared-memory ring (att_step_fwd_grouped_ring_kernel); ICD_BEAM_ATT_RING=0 selects the register-staged one-CTA-per-image
    kernel.  Same per-row arithmetic in the same order: alphas, scores and captions must agree bit for bit — with more live
    images than resident CTAs (every CTA draws several slot tickets, the ring wraps across images) and while slots die and get
    compacted."""
    import icd_b200.models.attention as my_att
    from icd_b200.gen_captions import beam_search_batched
    from icd_b200.vocabulary import synthetic_vocab
    case = dict(H.BEAM_CASES["beam_small"], dropout=0.5, train=False, fine_tune_embedding=True, k=k, n_img=330)   # > 2 CTAs x 148 SMs
    vocab = synthetic_vocab(case["V"])
    dec = H.build_attention_module(case, my_att.AttentionDecoder, my_att.AttentionDecoderParams, vocab)
    H.apply_beam_recipe(dec, case)
    V = case["V"]
    dec = dec.to(cuda)
    feats = H.beam_features(case).to(cuda)
    outs = []
    for ring in ("1", "0"):
        monkeypatch.setenv("ICD_BEAM_ATT_RING", ring)
        with torch.no_grad():
            outs.append(beam_search_batched(dec, feats, k, V - 3, V - 2, max_steps=20, want_alphas=True, want_trace=True,
                                            precision=precision))
    monkeypatch.delenv("ICD_BEAM_ATT_RING", raising=False)
    a, b = outs
    assert torch.equal(a["len"], b["len"]) and torch.equal(a["seq"], b["seq"]) and torch.equal(a["trace"], b["trace"])
    assert torch.equal(a["score"], b["score"])
    assert torch.equal(a["alpha"], b["alpha"])
    assert len(set(a["len"].tolist())) > 1          # captions of different lengths (0 = no beam completed): slots did die


@pytest.mark.gpu
@pytest.mark.parametrize("which", ["attention", "baseline"])
def test_fp32x3_stationary_weight_split_cache_is_bit_identical(cuda, which, monkeypatch):
    """fp32-grade tensor-core tier: the 3-term bf16 splits of the weights every step of the time loop reads ([W_dec; W_fbeta; W_hh],
    W_ih[:, E:], the baseline decoder's W_hh) are made once per forward / backward call and reused by the following steps
    (ICD_X3_CACHE=0: every contraction re-splits its operands, the round-1 behaviour).  Same split values, same contraction:
    outputs and gradients must agree bit for bit — ragged lengths, so the row count of the contractions changes between steps."""
    import icd_b200.models.attention as my_att
    import icd_b200.models.baseline as my_base
    from icd_b200.vocabulary import synthetic_vocab
    if which == "attention":
        case = dict(B=21, V=97, A=64, D=64, E=32, max_len=9, lengths=[9] * 6 + [7] * 5 + [4] * 10, wseed=6, iseed=41,
                    dropout=0.5, train=False, fine_tune_embedding=True)
        dec = H.build_attention_module(case, my_att.AttentionDecoder, my_att.AttentionDecoderParams, synthetic_vocab(case["V"]))
        enc = synthetic_features(case["B"], case["iseed"]).to(cuda)
    else:
        case = dict(H.BASE_CASES["base_small"])
        dec = H.build_baseline_module(case, my_base.BaselineDecoder, my_base.BaselineDecoderParams)
        enc = None
    dec = dec.to(cuda)
    dec.eval()
    dec.precision = "fp32x3"

    def run():
        dec.zero_grad()
        if which == "attention":
            caps, lens = synthetic_caps(case)
            preds, _, dl, alphas = dec(enc, caps.to(cuda), lens)
            O.attention_loss(preds, caps.to(cuda), dl, alphas).backward()
        else:
            img, caps, lens = H.base_inputs(case)
            preds = dec(img.to(cuda), caps.to(cuda))
            O.baseline_loss(preds, caps.to(cuda)).backward()
        return preds.detach().clone(), {k: p.grad.clone() for k, p in dec.named_parameters() if p.grad is not None}

    monkeypatch.delenv("ICD_X3_CACHE", raising=False)
    p1, g1 = run()
    monkeypatch.setenv("ICD_X3_CACHE", "0")
    p0, g0 = run()
    monkeypatch.delenv("ICD_X3_CACHE", raising=False)
    assert torch.equal(p1, p0)
    assert g1.keys() == g0.keys() and len(g1) >= 4
    for k in g1:
        assert torch.equal(g1[k], g0[k]), k


@pytest.mark.gpu
def test_beam_prologue_fused_split_and_pixel_mean_is_bit_identical(cuda, monkeypatch):
    """fp32-grade caption generation reads the features once in its prologue: one kernel writes the 3-term bf16 split of the
    pixel rows (A operand of the enc_att projection) and the pixel mean of init_hidden_state (gen_captions.py:62).
    ICD_BEAM_FUSED_MEAN_OFF=1 restores the two separate passes; the mean is summed in the same order, so every output must
    agree bit for bit (70 images: the projection runs in two passes of different size)."""
    import icd_b200.models.attention as my_att
    from icd_b200.gen_captions import beam_search_batched
    from icd_b200.vocabulary import synthetic_vocab
    case = dict(H.BEAM_CASES["beam_small"], dropout=0.5, train=False, fine_tune_embedding=True, k=3, n_img=70)
    dec = H.build_attention_module(case, my_att.AttentionDecoder, my_att.AttentionDecoderParams, synthetic_vocab(case["V"]))
    H.apply_beam_recipe(dec, case)
    dec = dec.to(cuda)
    V = case["V"]
    feats = H.beam_features(case).to(cuda)
    outs = []
    for off in (False, True):
        if off:
            monkeypatch.setenv("ICD_BEAM_FUSED_MEAN_OFF", "1")
        else:
            monkeypatch.delenv("ICD_BEAM_FUSED_MEAN_OFF", raising=False)
        with torch.no_grad():
            outs.append(beam_search_batched(dec, feats, 3, V - 3, V - 2, max_steps=20, want_alphas=True, want_trace=True))
    monkeypatch.delenv("ICD_BEAM_FUSED_MEAN_OFF", raising=False)
    a, b = outs
    for key in ("len", "seq", "score", "alpha", "trace"):
        assert torch.equal(a[key], b[key]), key


@pytest.mark.gpu
@pytest.mark.parametrize("which", ["train", "beam"])
def test_fp32x3_three_stored_planes_equal_six_segments(cuda, which, monkeypatch):
    """fp32-grade tier: where a K segment is a whole number of 64-wide k-blocks each operand is stored as THREE bf16 planes
    [t1 | t2 | t3] and the contraction's TMA producer walks them in the order of the six cross terms; ICD_X3_PLANES6=1 restores the
    six-segment K-concatenated copies.  Same MMAs on the same values in the same order: bit-identical outputs and gradients
    (train: D = A = E = 64 and C = 2048 take the three-plane path, K-major and MN-major operands, cached weight splits and
    split-K included; beam: the full-size decoder of the golden case, fused producers of the split activations included)."""
    import icd_b200.models.attention as my_att
    from icd_b200.gen_captions import beam_search_batched
    from icd_b200.vocabulary import synthetic_vocab

    def both(fn):
        monkeypatch.delenv("ICD_X3_PLANES6", raising=False)
        a = fn()
        monkeypatch.setenv("ICD_X3_PLANES6", "1")
        b = fn()
        monkeypatch.delenv("ICD_X3_PLANES6", raising=False)
        return a, b

    if which == "train":
        case = dict(B=21, V=97, A=64, D=64, E=64, max_len=9, lengths=[9] * 6 + [7] * 5 + [4] * 10, wseed=6, iseed=41,
                    dropout=0.5, train=False, fine_tune_embedding=True)
        dec = H.build_attention_module(case, my_att.AttentionDecoder, my_att.AttentionDecoderParams, synthetic_vocab(case["V"]))
        dec = dec.to(cuda)
        dec.eval()
        dec.precision = "fp32x3"
        enc = synthetic_features(case["B"], case["iseed"]).to(cuda).requires_grad_(True)
        caps, lens = synthetic_caps(case)
        caps = caps.to(cuda)

        def run():
            dec.zero_grad()
            enc.grad = None
            preds, _, dl, alphas = dec(enc, caps, lens)
            O.attention_loss(preds, caps, dl, alphas).backward()
            out = {k: p.grad.clone() for k, p in dec.named_parameters() if p.grad is not None}
            out["predictions"], out["alphas"], out["d_enc"] = preds.detach().clone(), alphas.detach().clone(), enc.grad.clone()
            return out
        a, b = both(run)
        assert a.keys() == b.keys() and len(a) > 10
    else:
        case = dict(H.BEAM_CASES["beam_cfg"], dropout=0.5, train=False, fine_tune_embedding=True)
        vocab = synthetic_vocab(case["V"])
        dec = H.build_attention_module(case, my_att.AttentionDecoder, my_att.AttentionDecoderParams, vocab)
        H.apply_beam_recipe(dec, case)
        dec = dec.to(cuda)
        feats = H.beam_features(case).to(cuda)
        V, k = case["V"], case["k"]

        def run():
            with torch.no_grad():
                return beam_search_batched(dec, feats, k, V - 3, V - 2, max_steps=30, want_alphas=True, want_trace=True)
        a, b = both(run)
    for key in a:
        assert torch.equal(a[key], b[key]), key
