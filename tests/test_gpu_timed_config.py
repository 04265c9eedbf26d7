"""GPU tier: parity AT THE CONFIGURATION bench.py TIMES (BASELINE.json configs[2]) — B = 512 captions of 25 tokens,
V = 9490, A = D = E = 512, embedding frozen, train mode with dropout 0.5, the bf16 tier with
``attention_caption_loss(bf16_grad_only=True)`` — and at the full-dimension configs[3] (glove) and configs[1] (baseline)
shapes, against the fp64 oracle on the SAME 512 rows (evaluated in row chunks: rows are independent, only the
parameters are shared; models/attention.py:396-430).

Stated bf16-tier tolerances (norm-wise relative per tensor; the measured values are printed, written to
$ICD_PARITY_TABLE when set and committed as profiles/r02_parity_table.md):
    logits, alphas                         5e-3
    loss                                   1e-3
    gradients                              1e-2
    the four attention-projection grads    5e-2   (ill-conditioned on iid synthetic features)
The fp32-class tiers on the same batch: logits / alphas 1e-4, loss 1e-5, gradients 1e-3 (fp32x3: 3e-3 on the four).
"""
import json
import os

import pytest
import torch

import helpers as H

pytestmark = pytest.mark.gpu

V, A, D, E, MAXLEN, B = 9490, 512, 512, 512, 25, 512
ILL = ("attention.enc_att.weight", "attention.enc_att.bias", "attention.dec_att.weight", "attention.dec_att.bias")
TIERS = ("bf16+bf16_grad_only", "bf16", "fp32x3", "fp32")


def _bench_module(cuda, precision):
    """Exactly bench.py's construction: seed 0, embedding frozen, train mode."""
    import icd_b200.models.attention as my_att
    from icd_b200.vocabulary import synthetic_vocab
    p = my_att.AttentionDecoderParams()
    p.vocab = synthetic_vocab(V)
    torch.manual_seed(0)
    dec = my_att.AttentionDecoder(cuda, p)
    dec.fine_tune_embeddings(False)
    sd = {k: v.detach().clone() for k, v in dec.state_dict().items()}
    dec = dec.to(cuda)
    dec.precision = precision
    dec.train()
    return dec, sd


@pytest.fixture(scope="module")
def timed(cuda):
    from icd_b200 import synthetic
    from icd_b200.losses import attention_caption_loss
    enc = synthetic.features(B, seed=1234)                       # bench.py rank 0 draws exactly these
    caps, lens = synthetic.captions(B, V, max_len=MAXLEN, seed=1234)
    T = MAXLEN - 1
    keep = H.seeded_keep_mask(T, B, D, 0.5, seed=4321)
    enc_d, caps_d = enc.to(cuda), caps.to(cuda)
    out = {}
    sd0 = None
    for tier in TIERS:
        precision = tier.split("+")[0]
        dec, sd = _bench_module(cuda, precision)
        sd0 = sd0 or sd
        dec._dropout_mask_override = keep
        preds, cs, dl, alphas = dec(enc_d, caps_d, lens)
        loss = attention_caption_loss(preds, cs, dl, alphas, alpha_c=1.0, bf16_grad_only=tier.endswith("grad_only"))
        loss.backward()
        out[tier] = dict(preds=preds.detach().cpu(), alphas=alphas.detach().cpu(), loss=float(loss.item()),
                         grads={k: p.grad.detach().cpu() for k, p in dec.named_parameters() if p.grad is not None})
        del dec, preds, alphas, loss
        torch.cuda.empty_cache()
    torch.set_num_threads(os.cpu_count() or 1)
    ref = H.oracle_fp64_chunked(sd0, enc, caps, lens, keep_mask=keep, p=0.5, frozen=("embedding.weight",), chunk=64,
                                cuda_preds={t: out[t]["preds"] for t in TIERS},
                                cuda_alphas={t: out[t]["alphas"] for t in TIERS})
    table = {}
    for tier in TIERS:
        row = {"predictions": ref["err"][tier]["predictions"], "alphas": ref["err"][tier]["alphas"],
               "loss": abs(out[tier]["loss"] - ref["loss"]) / abs(ref["loss"])}
        for k, g in ref["grads"].items():
            row["grad:" + k] = (float(out[tier]["grads"][k].abs().max()) if k == "attention.full_att.bias"
                                else H.rel_err(out[tier]["grads"][k], g))
        ids = out[tier]["preds"].argmax(dim=2)
        flips = (ids != ref["ids"])
        row["greedy_id_flips"] = int(flips.sum())
        row["greedy_id_flips_margin_gt_2e-2"] = int((flips & (ref["margin"] > 2e-2)).sum())
        row["positions"] = int(ids.numel())
        table[tier] = row
    path = os.environ.get("ICD_PARITY_TABLE")
    if path:
        with open(path, "w") as f:
            json.dump({"config": "configs[2] timed configuration: B=512, T=24, V=9490, A=D=E=512, dropout 0.5 (shared keep-mask), "
                                 "embedding frozen; fp64 oracle on the same 512 rows", "oracle_loss": ref["loss"],
                       "table": table}, f, indent=1)
    print("\nmeasured errors vs the fp64 oracle at the timed configuration:")
    for tier in TIERS:
        print(" ", tier, {k: ("%.2e" % v if isinstance(v, float) else v) for k, v in table[tier].items()})
    return dict(out=out, ref=ref, table=table)


@pytest.mark.parametrize("tier", TIERS)
def test_timed_configuration_matches_fp64_oracle(timed, tier):
    row = timed["table"][tier]
    bf16 = tier.startswith("bf16")
    assert row["predictions"] < (5e-3 if bf16 else 1e-4)
    assert row["alphas"] < (5e-3 if bf16 else 1e-4)
    assert row["loss"] < (1e-3 if bf16 else 1e-5)
    for k, v in row.items():
        if not k.startswith("grad:"):
            continue
        name = k[5:]
        if name == "attention.full_att.bias":
            assert v < 1e-4, "full_att.bias gradient (true value 0): max abs %.2e" % v
        elif name in ILL:
            assert v < (5e-2 if bf16 else (3e-3 if tier == "fp32x3" else 1e-3)), "%s: %.2e" % (name, v)
        else:
            assert v < (1e-2 if bf16 else 1e-3), "%s: %.2e" % (name, v)
    assert "grad:embedding.weight" not in row                     # frozen: no gradient
    # greedy ids: identical wherever the reference's top-2 margin is beyond the tier's logit error
    if bf16:
        assert row["greedy_id_flips_margin_gt_2e-2"] == 0
        assert row["greedy_id_flips"] <= 0.02 * row["positions"]
    else:
        ids = timed["out"][tier]["preds"].argmax(dim=2)
        assert not bool(((ids != timed["ref"]["ids"]) & (timed["ref"]["margin"] > 1e-4)).any())


def test_bf16_grad_only_is_bit_identical_at_the_timed_configuration(timed):
    a, b = timed["out"]["bf16"], timed["out"]["bf16+bf16_grad_only"]
    assert torch.equal(a["preds"], b["preds"]) and torch.equal(a["alphas"], b["alphas"]) and a["loss"] == b["loss"]
    for k in a["grads"]:
        assert torch.equal(a["grads"][k], b["grads"][k]), k


@pytest.mark.parametrize("precision,tol,gtol,gill", [("fp32", 1e-4, 1e-3, 2e-3), ("fp32x3", 1e-4, 1e-3, 3e-3),
                                                     ("bf16", 5e-3, 1e-2, 5e-2)])
def test_glove_config_full_dimensions(cuda, precision, tol, gtol, gill):
    """configs[3] at the model's real dimensions: E = 300, fp64 fine-tuned table (fp64 gradient), A = D = 512, V = 9490,
    a 32-row ragged batch (the golden `att_glove` case is A = D = 64)."""
    import icd_b200.models.attention as my_att
    from icd_b200 import synthetic
    from icd_b200.losses import attention_caption_loss
    from icd_b200.vocabulary import synthetic_vocab
    case = dict(B=32, V=V, A=A, D=D, E=300, max_len=MAXLEN, lengths="ragged", wseed=0, iseed=99, dropout=0.0,
                train=False, fine_tune_embedding=True, glove=True)
    dec = H.build_attention_module(case, my_att.AttentionDecoder, my_att.AttentionDecoderParams, synthetic_vocab(V))
    sd = {k: v.detach().clone() for k, v in dec.state_dict().items()}
    assert sd["embedding.weight"].dtype == torch.float64
    dec = dec.to(cuda)
    dec.precision = precision
    enc = synthetic.features(case["B"], seed=case["iseed"])
    caps, lens = synthetic.captions(case["B"], V, max_len=MAXLEN, seed=case["iseed"], lengths="ragged")
    preds, cs, dl, alphas = dec(enc.to(cuda), caps.to(cuda), lens)
    loss = attention_caption_loss(preds, cs, dl, alphas)
    loss.backward()
    ref = H.oracle_fp64_chunked(sd, enc, caps, lens, chunk=32, cuda_preds=preds.detach().cpu(),
                                cuda_alphas=alphas.detach().cpu())
    assert ref["err"]["predictions"] < tol and ref["err"]["alphas"] < tol
    assert abs(loss.item() - ref["loss"]) < max(tol, 1e-5) * abs(ref["loss"])
    g = dict(dec.named_parameters())
    assert g["embedding.weight"].grad.dtype == torch.float64
    for k, want in ref["grads"].items():
        if k == "attention.full_att.bias":
            continue
        H.assert_close_norm(g[k].grad, want, gill if k in ILL else gtol, "glove grad " + k, atol=1e-9)


@pytest.mark.parametrize("precision,tol,gtol", [("fp32", 1e-5, 1e-3), ("bf16", 1e-3, 2e-2)])
def test_baseline_caption_loss_matches_reference_expression(cuda, precision, tol, gtol):
    """icd_b200.losses.baseline_caption_loss == CrossEntropyLoss(ignore_index=<pad>) over scores.reshape(-1, V)
    (models/baseline.py:194-195, 224-225): value, and the gradients it sends through the CUDA baseline decoder, at
    configs[1] dimensions with ragged (padded) captions."""
    import icd_b200.models.baseline as my_base
    from icd_b200.losses import baseline_caption_loss
    from oracle import decoders as O
    case = dict(H.BASE_CASES["base_cfg2"], B=32, lengths="ragged")
    dec = H.build_baseline_module(case, my_base.BaselineDecoder, my_base.BaselineDecoderParams)
    w = {k: v.detach().clone().double().requires_grad_(True) for k, v in dec.state_dict().items()}
    dec = dec.to(cuda)
    dec.precision = precision
    img, caps, lens = H.base_inputs(case)
    assert int((caps == 0).sum()) > 0, "the case needs <pad> positions"
    img_dev = img.to(cuda).requires_grad_(True)
    outs = dec(img_dev, caps.to(cuda))
    loss = baseline_caption_loss(outs, caps.to(cuda))
    (loss * 0.7).backward()
    # value of the loss glue alone: the literal reference expression, fp64 torch, on the CUDA decoder's own logits
    lit = torch.nn.CrossEntropyLoss(ignore_index=0)(outs.detach().cpu().double().reshape(-1, V), caps.reshape(-1))
    assert abs(loss.item() - lit.item()) < 1e-6 * abs(lit.item())
    i64 = img.double().requires_grad_(True)
    o64 = O.baseline_decoder_forward(w, i64, caps)
    l64 = O.baseline_loss(o64, caps)
    (l64 * 0.7).backward()
    assert abs(loss.item() - l64.item()) < tol * abs(l64.item())
    H.assert_close_norm(img_dev.grad, i64.grad, gtol, "d img_features")
    for k, p in dec.named_parameters():
        H.assert_close_norm(p.grad, w[k].grad, gtol, "baseline grad " + k)
    # pad positions send no gradient into the logits
    d_out = torch.autograd.grad(baseline_caption_loss(outs2 := dec(img_dev, caps.to(cuda)), caps.to(cuda)), outs2)[0]
    assert torch.all(d_out[caps.to(cuda) == 0] == 0)


def test_data_parallel_clip_adam_single_process_equals_reference_optimizer_glue(cuda):
    """DataParallelClipAdam (world 1) == clip_gradient(opt, 5); opt.step() with torch.optim.Adam(lr=1e-4)
    (models/attention.py:352-355, 417-430; train_utils.py:2-12) over three real decoder steps."""
    import copy
    import icd_b200.models.attention as my_att
    from icd_b200.losses import attention_caption_loss
    from icd_b200.parallel import DataParallelClipAdam
    from icd_b200.train_utils import clip_gradient
    from icd_b200.vocabulary import synthetic_vocab
    case = dict(H.ATT_CASES["att_small_ragged"], fine_tune_embedding=False)
    dec_a = H.build_attention_module(case, my_att.AttentionDecoder, my_att.AttentionDecoderParams,
                                     synthetic_vocab(case["V"])).to(cuda)
    dec_b = copy.deepcopy(dec_a)
    lr = 1e-3
    opt_a = DataParallelClipAdam(dec_a, lr=lr, grad_clip=1e-3)            # a clip small enough to bite
    opt_b = torch.optim.Adam([p for p in dec_b.parameters() if p.requires_grad], lr=lr)
    enc, caps, lens = H.att_inputs(case)
    enc, caps = enc.to(cuda), caps.to(cuda)
    for step in range(3):
        for dec, opt in ((dec_a, opt_a), (dec_b, opt_b)):
            preds, cs, dl, alphas = dec(enc, caps, lens)
            loss = attention_caption_loss(preds, cs, dl, alphas)
            opt.zero_grad()
            loss.backward()
            if opt is opt_b:
                clip_gradient(opt, 1e-3)
            opt.step()
        for (k, pa), (_, pb) in zip(dec_a.named_parameters(), dec_b.named_parameters()):
            if step == 0:       # identical gradients in, one update: only the rounding of the update arithmetic differs
                assert float((pa - pb).abs().max()) <= 1e-3 * lr, "parameter after the first step: " + k
            else:               # Adam's update is ~lr * sign(g) early on: an element whose tiny gradient changes sign between
                d = (pa - pb).abs()       # the two (now slightly different) models moves by up to 2 lr; all others agree
                assert float(d.max()) <= 2.1 * lr * (step + 1), "parameter after step %d: %s" % (step + 1, k)
                if k != "attention.full_att.bias":        # (one element whose true gradient is 0: pure sign noise)
                    assert float(d.mean()) <= 2e-2 * lr, "parameter after step %d: %s" % (step + 1, k)
    assert torch.equal(dec_a.embedding.weight, dec_b.embedding.weight)       # frozen
    assert opt_a.step_count == 3
