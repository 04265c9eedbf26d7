"""Generate the golden fixtures in this directory FROM THE UNMODIFIED REFERENCE.

Runs only in the build container (needs /root/reference, imported through oracle/reference_import.py with
five stubbed third-party modules).  For every case in tests/helpers.py it

  1. builds the reference module under ``torch.manual_seed(wseed)`` and the icd_b200 drop-in the same way, and
     asserts their ``state_dict`` are bit-identical (so the GPU-box tests can rebuild the weights from the seed);
  2. runs the reference forward (+ the reference train-loop loss glue and ``backward()`` where defined);
  3. runs oracle/decoders.py on the same inputs and asserts agreement (this is what pins the oracle);
  4. stores outputs / gradients (whole tensors for the small cases, digests for the full-size ones) plus the
     weight checksums into ``<case>.npz`` and the oracle-vs-reference deltas into PINNING.json.

    python tests/golden/make_golden.py
"""
import contextlib
import io
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import helpers as H  # noqa: E402
from oracle import decoders as O  # noqa: E402
from oracle.reference_import import load_reference, make_reference_vocab  # noqa: E402

import icd_b200.models.attention as my_att  # noqa: E402
import icd_b200.models.baseline as my_base  # noqa: E402
from icd_b200.vocabulary import synthetic_vocab  # noqa: E402

torch.set_num_threads(8)
ns = load_reference()
pinning = {}


def surrogate_weights(case, shape_p, shape_a):
    g = torch.Generator().manual_seed(case["iseed"] + 1000)
    return torch.randn(shape_p, generator=g), torch.randn(shape_a, generator=g)


def store_tensor(out, case, name, t):
    if case["store"] == "full":
        out[name] = t.detach().cpu().numpy()
    else:
        dg = H.digest(name, t)
        out[name + "@norm"] = np.float64(dg["norm"])
        out[name + "@sum"] = np.float64(dg["sum"])
        out[name + "@samples"] = dg["samples"]


def run_attention_case(name, case):
    ref_vocab = make_reference_vocab(ns, case["V"])
    ref = H.build_attention_module(case, ns.AttentionDecoder, ns.AttentionDecoderParams, ref_vocab)
    mine = H.build_attention_module(case, my_att.AttentionDecoder, my_att.AttentionDecoderParams,
                                    synthetic_vocab(case["V"]))
    sd_r, sd_m = ref.state_dict(), mine.state_dict()
    assert list(sd_r.keys()) == list(sd_m.keys()), (list(sd_r.keys()), list(sd_m.keys()))
    for k in sd_r:
        assert sd_r[k].dtype == sd_m[k].dtype and torch.equal(sd_r[k], sd_m[k]), "init differs: " + k
    enc, caps, lens = H.att_inputs(case)
    out = {}
    if case["train"]:
        torch.manual_seed(case["mask_seed"])
    preds, caps_out, dl, alphas = ref(enc, caps, lens)
    assert caps_out is caps
    use_loss = case.get("loss", True)
    if use_loss:
        from torch.nn.utils.rnn import pack_padded_sequence
        targets = caps_out[:, 1:]
        sc = pack_padded_sequence(preds, dl, batch_first=True).data          # models/attention.py:405-408
        tg = pack_padded_sequence(targets, dl, batch_first=True).data
        loss = torch.nn.CrossEntropyLoss()(sc, tg)                           # :411
        loss = loss + ((1.0 - alphas.sum(dim=1)) ** 2).mean()                # :414 (alpha_c = 1)
    else:
        g1, g2 = surrogate_weights(case, preds.shape, alphas.shape)
        loss = (preds * g1).sum() + (alphas * g2).sum()
    ref.zero_grad()
    loss.backward()
    grads = {k: p.grad for k, p in ref.named_parameters()}

    # ---- oracle pinning (fp32, same op order => expect bitwise / ~1e-7) ----
    frozen = () if case["fine_tune_embedding"] else ("embedding.weight",)
    w = {k: v.detach().clone().requires_grad_(k not in frozen) for k, v in ref.state_dict().items()}
    masks = H.dropout_masks_like_reference(case, dl, case["D"]) if case["train"] else None
    o_preds, _, o_dl, o_alphas = O.attention_decoder_forward(w, enc, caps, lens, dropout_p=case["dropout"],
                                                             dropout_masks=masks)
    assert o_dl == dl
    if use_loss:
        o_loss = O.attention_loss(o_preds, caps, o_dl, o_alphas)
    else:
        o_loss = (o_preds * g1).sum() + (o_alphas * g2).sum()
    o_loss.backward()
    pin = dict(pred_maxabs=float((o_preds - preds).abs().max()), alpha_maxabs=float((o_alphas - alphas).abs().max()),
               loss_abs=float((o_loss - loss).abs()))
    for k, gr in grads.items():
        if gr is None:
            assert w[k].grad is None, k
            continue
        pin["grad:" + k] = H.rel_err(w[k].grad, gr) if float(gr.norm()) > 1e-12 else float((w[k].grad - gr).abs().max())
    pinning[name] = pin
    assert pin["pred_maxabs"] < 1e-5 and pin["alpha_maxabs"] < 1e-6, pin
    for k, v in pin.items():
        if k.startswith("grad:") and "full_att.bias" not in k:
            assert v < 2e-4, (k, v)

    store_tensor(out, case, "predictions", preds)
    store_tensor(out, case, "alphas", alphas)
    out["loss"] = np.float64(loss.item())
    out["decode_lengths"] = np.asarray(dl, dtype=np.int64)
    for k, gr in grads.items():
        if gr is not None:
            store_tensor(out, case, "grad:" + k, gr)
    out["grad_names"] = np.asarray([k for k, gr in grads.items() if gr is not None])
    if use_loss:
        ids = O.teacher_forced_ids(preds, dl)
        out["greedy_ids"] = np.asarray([i + [-1] * (max(dl) - len(i)) for i in ids], dtype=np.int64)
        # top-2 logit margins so the GPU tests only require identical argmax where it is meaningful
        top2 = preds.topk(2, dim=2).values
        out["greedy_margin"] = (top2[..., 0] - top2[..., 1]).detach().numpy()
    cs = H.state_checksums(ref)
    out["weight_keys"] = np.asarray(list(cs.keys()))
    out["weight_checksums"] = np.asarray(list(cs.values()))
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print("wrote", name, {k: ("%.2e" % v) for k, v in pin.items() if not k.startswith("grad:")})


def run_baseline_case(name, case):
    ref = H.build_baseline_module(case, ns.BaselineDecoder, ns.BaselineDecoderParams)
    mine = H.build_baseline_module(case, my_base.BaselineDecoder, my_base.BaselineDecoderParams)
    sd_r, sd_m = ref.state_dict(), mine.state_dict()
    assert list(sd_r.keys()) == list(sd_m.keys())
    for k in sd_r:
        assert torch.equal(sd_r[k], sd_m[k]), "init differs: " + k
    img, caps, lens = H.base_inputs(case)
    img = img.clone().requires_grad_(True)
    outs = ref(img, caps)
    loss = torch.nn.CrossEntropyLoss(ignore_index=0)(outs.reshape(-1, outs.shape[2]), caps.reshape(-1))   # baseline.py:194,224
    ref.zero_grad()
    loss.backward()
    grads = {k: p.grad for k, p in ref.named_parameters()}
    grads["img_features"] = img.grad

    w = {k: v.detach().clone().requires_grad_(True) for k, v in ref.state_dict().items()}
    img2 = img.detach().clone().requires_grad_(True)
    o_outs = O.baseline_decoder_forward(w, img2, caps)
    o_loss = O.baseline_loss(o_outs, caps)
    o_loss.backward()
    pin = dict(out_maxabs=float((o_outs - outs).abs().max()), loss_abs=float((o_loss - loss).abs()))
    for k, gr in grads.items():
        og = img2.grad if k == "img_features" else w[k].grad
        pin["grad:" + k] = H.rel_err(og, gr)
    pinning[name] = pin
    assert pin["out_maxabs"] < 1e-5, pin
    for k, v in pin.items():
        if k.startswith("grad:"):
            assert v < 2e-4, (k, v)
    out = {}
    store_tensor(out, case, "outputs", outs)
    out["loss"] = np.float64(loss.item())
    for k, gr in grads.items():
        store_tensor(out, case, "grad:" + k, gr)
    out["grad_names"] = np.asarray(list(grads.keys()))
    ids = outs.argmax(dim=2)
    out["greedy_ids"] = ids.numpy()
    top2 = outs.topk(2, dim=2).values
    out["greedy_margin"] = (top2[..., 0] - top2[..., 1]).detach().numpy()
    cs = H.state_checksums(ref)
    out["weight_keys"] = np.asarray(list(cs.keys()))
    out["weight_checksums"] = np.asarray(list(cs.values()))
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print("wrote", name, {k: ("%.2e" % v) for k, v in pin.items() if not k.startswith("grad:")})


class _Identity(torch.nn.Module):
    def forward(self, x):
        return x


def run_beam_case(name, case):
    ref_vocab = make_reference_vocab(ns, case["V"])
    acase = dict(case, dropout=0.5, train=False, fine_tune_embedding=True)
    ref = H.build_attention_module(acase, ns.AttentionDecoder, ns.AttentionDecoderParams, ref_vocab)
    H.apply_beam_recipe(ref, case)
    feats = H.beam_features(case)
    V = case["V"]

    class Args:
        beam_size = case["k"]
    out = {}
    pin = {}
    lens = []
    with torch.no_grad():
        w = {k: v.detach().clone() for k, v in ref.state_dict().items()}
        w64 = {k: v.detach().double() for k, v in ref.state_dict().items()}
        for i in range(case["n_img"]):
            buf = io.StringIO()
            with contextlib.redirect_stdout(buf):
                seq, alphas, ended = ns.beam_search(torch.device("cpu"), Args, feats[i:i + 1], _Identity(), ref, ref_vocab)
            trace = []
            o_seq, o_alphas, o_ended = O.beam_search(w, feats[i:i + 1], case["k"], V - 3, V - 2, trace=trace)
            assert o_seq == seq and o_ended == ended, (i, seq, o_seq)
            if alphas:
                d = float(np.abs(np.asarray(alphas) - np.asarray(o_alphas)).max())
                assert d < 1e-6, d
            # the printed per-step words are the reference's own trace (gen_captions.py:91)
            printed = [ln for ln in buf.getvalue().strip().split("\n") if ln]
            assert len(printed) == len(trace)
            for ln, words in zip(printed, trace):
                assert ln == str([ref_vocab.i2w[x] for x in words]), (ln, words)
            # fp64 run of the oracle: is the caption stable under arithmetic noise? (SURVEY 8c caution (v))
            s64, _, e64 = O.beam_search(w64, feats[i:i + 1], case["k"], V - 3, V - 2)
            out["seq_%d" % i] = np.asarray(seq, dtype=np.int64)
            out["ended_%d" % i] = np.asarray(ended)
            out["stable_%d" % i] = np.asarray(s64 == seq and e64 == ended)
            out["alphas_%d" % i] = np.asarray(alphas, dtype=np.float32)
            tr = np.full((len(trace), case["k"]), -1, dtype=np.int64)
            for s, words in enumerate(trace):
                tr[s, :len(words)] = words
            out["trace_%d" % i] = tr
            lens.append(len(seq))
    pin["lens"] = lens
    pin["stable"] = [bool(out["stable_%d" % i]) for i in range(case["n_img"])]
    pinning[name] = pin
    cs = H.state_checksums(ref)
    out["weight_keys"] = np.asarray(list(cs.keys()))
    out["weight_checksums"] = np.asarray(list(cs.values()))
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print("wrote", name, pin)


if __name__ == "__main__":
    only = sys.argv[1:]
    for name, case in H.ATT_CASES.items():
        if not only or name in only:
            run_attention_case(name, case)
    for name, case in H.BASE_CASES.items():
        if not only or name in only:
            run_baseline_case(name, case)
    for name, case in H.BEAM_CASES.items():
        if not only or name in only:
            run_beam_case(name, case)
    path = os.path.join(HERE, "PINNING.json")
    old = json.load(open(path)) if os.path.exists(path) else {}
    old.update(pinning)
    json.dump(old, open(path, "w"), indent=1, sort_keys=True)
