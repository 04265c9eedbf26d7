"""CPU tier: the oracle (oracle/decoders.py) against the golden vectors generated from the unmodified reference
(tests/golden/make_golden.py).  Also checks that the icd_b200 module constructors reproduce the reference's
seeded initialisation (weight checksums stored with every golden)."""
import json
import os

import numpy as np
import pytest
import torch

import helpers as H
from oracle import decoders as O
from oracle import reference_import

import icd_b200.models.attention as my_att
import icd_b200.models.baseline as my_base
from icd_b200.vocabulary import synthetic_vocab


def load(name):
    return np.load(os.path.join(H.GOLDEN_DIR, name + ".npz"), allow_pickle=False)


def check_weights(module, g):
    cs = H.state_checksums(module)
    assert list(cs.keys()) == [str(k) for k in g["weight_keys"]]
    assert list(cs.values()) == [str(v) for v in g["weight_checksums"]]


def compare(case, g, name, t, tol, atol=0.0):
    if case["store"] == "full":
        H.assert_close_norm(t, torch.from_numpy(g[name]), tol, name, atol)
    else:
        H.assert_digest_close(name, t, dict(norm=g[name + "@norm"], samples=g[name + "@samples"]), tol, atol)


@pytest.mark.parametrize("name", list(H.ATT_CASES))
def test_attention_oracle_matches_reference_golden(name):
    case = H.ATT_CASES[name]
    if name == "att_cfg1":
        torch.set_num_threads(max(1, os.cpu_count() or 1))
    g = load(name)
    mine = H.build_attention_module(case, my_att.AttentionDecoder, my_att.AttentionDecoderParams,
                                    synthetic_vocab(case["V"]))
    check_weights(mine, g)
    enc, caps, lens = H.att_inputs(case)
    frozen = () if case["fine_tune_embedding"] else ("embedding.weight",)
    w = {k: v.detach().clone().requires_grad_(k not in frozen) for k, v in mine.state_dict().items()}
    masks = None
    dl = [l - 1 for l in lens]
    if case["train"]:
        masks = H.dropout_masks_like_reference(case, dl, case["D"])
    preds, caps_out, odl, alphas = O.attention_decoder_forward(w, enc, caps, lens, dropout_p=case["dropout"],
                                                               dropout_masks=masks)
    assert caps_out is caps and odl == list(g["decode_lengths"])
    compare(case, g, "predictions", preds, 1e-6)
    compare(case, g, "alphas", alphas, 1e-6)
    if case.get("loss", True):
        loss = O.attention_loss(preds, caps, odl, alphas)
        assert abs(loss.item() - float(g["loss"])) <= 1e-6 * abs(float(g["loss"]))
        ids = O.teacher_forced_ids(preds, odl)
        gold = g["greedy_ids"]
        for j, row in enumerate(ids):
            assert row == list(gold[j, :len(row)])
    else:
        gen = torch.Generator().manual_seed(case["iseed"] + 1000)
        g1 = torch.randn(preds.shape, generator=gen)
        g2 = torch.randn(alphas.shape, generator=gen)
        loss = (preds * g1).sum() + (alphas * g2).sum()
    loss.backward()
    for k in [str(x) for x in g["grad_names"]]:
        tol, atol = (1e-5, 0.0)
        if k == "attention.full_att.bias":
            tol, atol = (0.0, 1e-6)          # true value is identically 0
        compare(case, g, "grad:" + k, w[k].grad, tol, atol)
    if not case["fine_tune_embedding"]:
        assert "embedding.weight" not in [str(x) for x in g["grad_names"]]


def test_hoisted_restatement_equals_per_step_recompute():
    """enc_att(encoder_out) hoisted out of the loop (what the kernels do) == per-step recompute (reference)."""
    case = H.ATT_CASES["att_small_ragged"]
    mine = H.build_attention_module(case, my_att.AttentionDecoder, my_att.AttentionDecoderParams,
                                    synthetic_vocab(case["V"]))
    enc, caps, lens = H.att_inputs(case)
    w = O.cast_weights(mine.state_dict(), torch.float32)
    p1, _, _, a1 = O.attention_decoder_forward(w, enc, caps, lens)
    p2, _, _, a2 = O.attention_decoder_forward(w, enc, caps, lens, hoist=True)
    H.assert_close_norm(p2, p1, 1e-6, "hoisted predictions")
    H.assert_close_norm(a2, a1, 1e-6, "hoisted alphas")


def test_fp64_oracle_is_close_to_fp32_reference():
    case = H.ATT_CASES["att_small_ragged"]
    g = load("att_small_ragged")
    mine = H.build_attention_module(case, my_att.AttentionDecoder, my_att.AttentionDecoderParams,
                                    synthetic_vocab(case["V"]))
    enc, caps, lens = H.att_inputs(case)
    w64 = O.cast_weights(mine.state_dict(), torch.float64)
    preds, _, _, alphas = O.attention_decoder_forward(w64, enc.double(), caps, lens)
    H.assert_close_norm(preds, torch.from_numpy(g["predictions"]), 1e-5, "fp64 vs reference fp32 predictions")
    H.assert_close_norm(alphas, torch.from_numpy(g["alphas"]), 1e-5, "fp64 vs reference fp32 alphas")


@pytest.mark.parametrize("name", list(H.BASE_CASES))
def test_baseline_oracle_matches_reference_golden(name):
    case = H.BASE_CASES[name]
    g = load(name)
    mine = H.build_baseline_module(case, my_base.BaselineDecoder, my_base.BaselineDecoderParams)
    check_weights(mine, g)
    img, caps, lens = H.base_inputs(case)
    w = {k: v.detach().clone().requires_grad_(True) for k, v in mine.state_dict().items()}
    img = img.clone().requires_grad_(True)
    outs = O.baseline_decoder_forward(w, img, caps)
    compare(case, g, "outputs", outs, 2e-6)
    loss = O.baseline_loss(outs, caps)
    assert abs(loss.item() - float(g["loss"])) <= 1e-6 * abs(float(g["loss"]))
    loss.backward()
    for k in [str(x) for x in g["grad_names"]]:
        gr = img.grad if k == "img_features" else w[k].grad
        compare(case, g, "grad:" + k, gr, 2e-5)


@pytest.mark.parametrize("name", list(H.BEAM_CASES))
def test_beam_oracle_matches_reference_golden(name):
    case = H.BEAM_CASES[name]
    g = load(name)
    acase = dict(case, dropout=0.5, train=False, fine_tune_embedding=True)
    mine = H.build_attention_module(acase, my_att.AttentionDecoder, my_att.AttentionDecoderParams,
                                    synthetic_vocab(case["V"]))
    H.apply_beam_recipe(mine, case)
    check_weights(mine, g)
    feats = H.beam_features(case)
    V = case["V"]
    w = O.cast_weights(mine.state_dict(), torch.float32)
    n = case["n_img"] if name == "beam_small" else 3          # full-size case: a subset keeps the CPU tier fast
    for i in range(n):
        trace = []
        seq, alphas, ended = O.beam_search(w, feats[i:i + 1], case["k"], V - 3, V - 2, trace=trace)
        assert seq == list(g["seq_%d" % i]) and ended == bool(g["ended_%d" % i])
        gt = g["trace_%d" % i]
        assert len(trace) == gt.shape[0]
        for s, words in enumerate(trace):
            assert words == [x for x in gt[s] if x >= 0]
        if alphas:
            assert np.abs(np.asarray(alphas, dtype=np.float32) - g["alphas_%d" % i]).max() < 1e-6


def test_pinning_report_present():
    rep = json.load(open(os.path.join(H.GOLDEN_DIR, "PINNING.json")))
    for name in list(H.ATT_CASES) + list(H.BASE_CASES) + list(H.BEAM_CASES):
        assert name in rep


@pytest.mark.skipif(not reference_import.reference_available(), reason="needs the reference tree (build container only)")
def test_reference_whole_module_checkpoint_loads_as_drop_in(tmp_path):
    """The reference pickles whole modules (checkpoint.py:51-59).  icd_b200.checkpoint maps the pickled class paths
    (models.attention.*, models.baseline.*, vocabulary.Vocabulary) onto the drop-in classes: same weights, same keys."""
    import sys
    import icd_b200.models.attention as my_att
    import icd_b200.models.baseline as my_base
    from icd_b200 import checkpoint as ckpt
    from icd_b200.vocabulary import Vocabulary as MyVocab
    ns = reference_import.load_reference()
    case = H.ATT_CASES["att_small_ragged"]
    ref_dec = H.build_attention_module(case, ns.AttentionDecoder, ns.AttentionDecoderParams,
                                       reference_import.make_reference_vocab(ns, case["V"]))
    bcase = H.BASE_CASES["base_small"]
    ref_base = H.build_baseline_module(bcase, ns.BaselineDecoder, ns.BaselineDecoderParams)
    opt = torch.optim.Adam(ref_dec.parameters(), lr=1e-4)
    path = tmp_path / "basic_att_0.pth.tar"
    # pickling needs the reference modules importable under their own names (they are parked under _icd_ref.* otherwise)
    names = ["models", "models.attention", "models.baseline", "vocabulary"]
    saved = {k: sys.modules.get(k) for k in names}
    try:
        for k in names:
            sys.modules[k] = sys.modules["_icd_ref." + k]
        torch.save({"epoch": 0, "metrics": {"loss": [1.0]}, "encoder": None, "decoder": ref_dec,
                    "encoder_optimizer": None, "decoder_optimizer": opt, "baseline": ref_base}, str(path))
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    chk = ckpt.load_checkpoint_file(str(path))
    epoch, enc, dec, enc_opt, dec_opt, metrics = ckpt.unpack_checkpoint(chk)
    assert epoch == 0 and metrics == {"loss": [1.0]}
    assert type(dec) is my_att.AttentionDecoder and type(dec.attention) is my_att.SoftAttention
    assert type(dec.vocab) is MyVocab and len(dec.vocab) == case["V"] and dec.vocab("<end>") == case["V"] - 2
    assert dec.precision == "fp32" and dec._dropout_mask_override is None
    sd_ref, sd = ref_dec.state_dict(), dec.state_dict()
    assert list(sd) == list(sd_ref) and all(torch.equal(sd[k], sd_ref[k]) for k in sd)
    base = chk["baseline"]
    assert type(base) is my_base.BaselineDecoder and base.precision == "fp32"
    assert all(torch.equal(v, ref_base.state_dict()[k]) for k, v in base.state_dict().items())
    assert isinstance(dec_opt, torch.optim.Adam)


@pytest.mark.skipif(not reference_import.reference_available(), reason="needs the reference tree (build container only)")
def test_saved_checkpoint_loads_in_the_unmodified_reference(tmp_path, monkeypatch):
    """The other direction: a checkpoint written by icd_b200.checkpoint.save_checkpoint — decoder whose parameters
    live in a FlatParamBuffer, fused DataParallelClipAdam optimiser with two steps of state — is read by a plain
    ``torch.load`` with only the REFERENCE's modules importable, comes back as the reference's own classes, computes the
    reference forward, and its ``decoder_optimizer`` is a torch.optim.Adam that steps."""
    import sys
    import icd_b200.models.attention as my_att
    from icd_b200 import checkpoint as ckpt
    from icd_b200.parallel import DataParallelClipAdam
    from icd_b200.vocabulary import synthetic_vocab
    ns = reference_import.load_reference()
    case = H.ATT_CASES["att_small_ragged"]
    dec = H.build_attention_module(case, my_att.AttentionDecoder, my_att.AttentionDecoderParams, synthetic_vocab(case["V"]))
    opt = DataParallelClipAdam(dec, lr=1e-3, grad_clip=5.0)          # re-points the parameters into one flat buffer
    g = torch.Generator().manual_seed(1)
    opt.step_count = 2                                               # two steps of (synthetic) optimiser state
    opt.exp_avg.copy_(torch.randn(opt.exp_avg.shape, generator=g) * 1e-3)
    opt.exp_avg_sq.copy_(torch.rand(opt.exp_avg_sq.shape, generator=g) * 1e-6)
    monkeypatch.setattr(ckpt, "CHECKPOINTS_DIR", str(tmp_path))

    class Args:
        model_name = "basic_att"
        checkpoint = "basic_att_3.pth.tar"
    ckpt.save_checkpoint(Args, 3, None, dec, None, opt, {"loss": [2.5]}, verbose=False)
    assert my_att.AttentionDecoder.__module__ == my_att.__name__ and my_att.AttentionDecoder.__qualname__ == "AttentionDecoder"
    path = tmp_path / "basic_att_3.pth.tar"
    names = ["models", "models.attention", "models.baseline", "vocabulary"]
    saved = {k: sys.modules.get(k) for k in names}
    try:
        for k in names:
            sys.modules[k] = sys.modules["_icd_ref." + k]
        chk = torch.load(str(path), map_location="cpu", weights_only=False)     # the reference's call (checkpoint.py:18)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    rdec = chk["decoder"]
    assert type(rdec) is ns.AttentionDecoder and type(rdec.attention) is ns.SoftAttention
    assert type(rdec.vocab) is ns.Vocabulary and len(rdec.vocab) == case["V"]
    assert chk["epoch"] == 3 and chk["metrics"] == {"loss": [2.5]}
    sd, rsd = dec.state_dict(), rdec.state_dict()
    assert list(sd) == list(rsd) and all(torch.equal(sd[k], rsd[k]) for k in sd)
    storages = {p_.untyped_storage().data_ptr() for p_ in rdec.parameters()}
    assert len(storages) == len(list(rdec.parameters())), "saved parameters must own their storage (not views of a flat buffer)"
    enc, caps, lens = H.att_inputs(case)
    rdec.eval()
    with torch.no_grad():
        preds, _, dl, alphas = rdec(enc, caps, lens)                             # the reference's own forward
        o_preds, _, o_dl, o_alphas = O.attention_decoder_forward(O.cast_weights(sd, torch.float32), enc, caps, lens)
    assert torch.equal(preds, o_preds) and torch.equal(alphas, o_alphas) and dl == o_dl
    ropt = chk["decoder_optimizer"]
    assert type(ropt) is torch.optim.Adam and ropt.param_groups[0]["lr"] == 1e-3
    rparams = [p_ for p_ in rdec.parameters() if p_.requires_grad]
    assert [id(p_) for p_ in ropt.param_groups[0]["params"]] == [id(p_) for p_ in rparams]
    rnames = [k for k, p_ in rdec.named_parameters() if p_.requires_grad]
    for k, p_ in zip(rnames, rparams):
        st = ropt.state[p_]
        off, n = opt.buf.offsets[k]
        assert int(st["step"]) == 2 and n == p_.numel()
        assert torch.equal(st["exp_avg"].reshape(-1), opt.exp_avg[off:off + n])
    rdec.train()
    preds, cs, dl, alphas = rdec(enc, caps, lens)
    O.attention_loss(preds, cs, dl, alphas).backward()
    ropt.step()                                                                  # resumes like models/attention.py:359-364
    # and back: this package reads its own file as drop-in modules; the Adam state moves into the fused optimiser
    chk2 = ckpt.load_checkpoint(torch.device("cpu"), Args, verbose=False)
    assert type(chk2["decoder"]) is my_att.AttentionDecoder
    opt2 = DataParallelClipAdam(chk2["decoder"], lr=1.0)
    opt2.load_torch_adam(chk2["decoder_optimizer"])
    assert opt2.step_count == 2 and opt2.lr == 1e-3
    for k, (off, n) in opt.buf.offsets.items():
        off2, n2 = opt2.buf.offsets[k]
        assert n == n2 and torch.equal(opt2.exp_avg[off2:off2 + n], opt.exp_avg[off:off + n]), k
        assert torch.equal(opt2.exp_avg_sq[off2:off2 + n], opt.exp_avg_sq[off:off + n]), k


@pytest.mark.skipif(not reference_import.reference_available(), reason="needs the reference tree (build container only)")
def test_oracle_precomputed_embedding_branch_equals_reference_use_bert_branch(monkeypatch):
    """models/attention.py:242-244, :273 — with ``use_bert`` the loop consumes (B, L, 768) vectors from
    ``_create_bert_embeddings`` instead of the table lookup.  BERT itself is unavailable offline, so the reference decoder is
    built with stand-in tokenizer / model classes and its ``_create_bert_embeddings`` is replaced by a function returning
    fixed vectors: everything downstream is the reference's own code.  The oracle's ``embeddings=`` path must agree exactly,
    forward and backward."""
    ns = reference_import.load_reference()

    class _Fake:
        @classmethod
        def from_pretrained(cls, name):
            return cls()

        def to(self, device):
            return self

        def eval(self):
            return self
    monkeypatch.setattr(ns.attention, "BertTokenizer", _Fake)
    monkeypatch.setattr(ns.attention, "BertModel", _Fake)
    V, B, L = 53, 4, 7
    p = ns.AttentionDecoderParams()
    p.attention_dim, p.decoder_dim, p.embed_size, p.dropout, p.use_bert = 24, 16, 768, 0.5, True
    p.vocab = reference_import.make_reference_vocab(ns, V)
    torch.manual_seed(5)
    dec = ns.AttentionDecoder(torch.device("cpu"), p)
    dec.eval()
    g = torch.Generator().manual_seed(6)
    emb = torch.randn(B, L, 768, generator=g)
    dec._create_bert_embeddings = lambda caps: emb
    case = dict(B=B, V=V, max_len=L, iseed=3, lengths=[7, 6, 6, 3])
    enc, caps, lens = H.att_inputs(case)
    preds, _, dl, alphas = dec(enc, caps, lens)
    O.attention_loss(preds, caps, dl, alphas).backward()
    w = {k: v.detach().clone().requires_grad_(True) for k, v in dec.state_dict().items()}
    o_preds, _, o_dl, o_alphas = O.attention_decoder_forward(w, enc, caps, lens, embeddings=emb)
    O.attention_loss(o_preds, caps, o_dl, o_alphas).backward()
    assert dl == o_dl and torch.equal(preds, o_preds) and torch.equal(alphas, o_alphas)
    for k, q in dec.named_parameters():
        if k == "embedding.weight":
            assert q.grad is None and w[k].grad is None          # the table is bypassed
        else:
            assert torch.equal(q.grad, w[k].grad), k


@pytest.mark.skipif(not reference_import.reference_available(), reason="needs the reference tree (build container only)")
def test_oracle_evaluate_restatement_equals_reference_per_image_calls():
    """oracle.evaluate_reference_style restates the body of evaluate() (models/attention.py:516-553; the function itself builds
    a COCO dataset and cannot run offline): checked against the unmodified reference decoder called once per image with the
    literal loss / argmax / cleaning expressions of those lines."""
    from torch.nn.utils.rnn import pack_padded_sequence
    ns = reference_import.load_reference()
    case = H.ATT_CASES["att_small_ragged"]
    dec = H.build_attention_module(case, ns.AttentionDecoder, ns.AttentionDecoderParams,
                                   reference_import.make_reference_vocab(ns, case["V"]))
    dec.eval()
    vocab = dec.vocab
    enc, caps, lens = H.att_inputs(case)
    special = [vocab(ns.vocabulary.START_TOKEN), vocab(ns.vocabulary.END_TOKEN), vocab(ns.vocabulary.PAD_TOKEN)]
    losses, hyps, refs = O.evaluate_reference_style(O.cast_weights(dec.state_dict(), torch.float32), enc, caps, lens, special)
    criterion = torch.nn.CrossEntropyLoss()
    with torch.no_grad():
        for j in range(case["B"]):
            c1 = caps[j:j + 1, :lens[j]]
            scores, cs, dl, aw = dec(enc[j:j + 1], c1, [lens[j]])
            targets = cs[:, 1:]
            sp = pack_padded_sequence(scores, dl, batch_first=True).data
            tp = pack_padded_sequence(targets, dl, batch_first=True).data
            loss = criterion(sp, tp)
            loss += ((1. - aw.sum(dim=1)) ** 2).mean()
            assert abs(loss.item() - losses[j]) < 1e-6
            img_captions = targets[0].tolist()
            cleaned = [x for x in img_captions if x not in special]
            assert refs[j] == list(map(lambda c: cleaned, img_captions))
            _, pr = torch.max(scores, dim=2)
            assert hyps[j] == [x for x in pr.tolist()[0][:dl[0]] if x not in special]


def test_oracle_loss_glue_equals_pack_padded_sequence_expression():
    """The oracle restates models/attention.py:401-414 without pack_padded_sequence (explicit time-major gather of the
    first batch_size_t rows); check it against the reference's literal expression on random ragged, sorted lengths,
    and the teacher-forced ids against torch.max(dim=2) truncated to decode_lengths (:544-553)."""
    from torch.nn.utils.rnn import pack_padded_sequence
    g = torch.Generator().manual_seed(3)
    for trial in range(6):
        B, T, V, P = 2 + trial, 3 + 2 * trial, 37 + trial, 11
        dl = sorted((int(x) for x in torch.randint(1, T + 1, (B,), generator=g)), reverse=True)
        dl[0] = T
        preds = torch.randn(B, T, V, generator=g, dtype=torch.float64)
        alphas = torch.rand(B, T, P, generator=g, dtype=torch.float64)
        caps = torch.randint(0, V, (B, T + 1), generator=g)
        scores = pack_padded_sequence(preds, dl, batch_first=True).data
        targets = pack_padded_sequence(caps[:, 1:], dl, batch_first=True).data
        ref = torch.nn.CrossEntropyLoss()(scores, targets) + ((1.0 - alphas.sum(dim=1)) ** 2).mean()
        got = O.attention_loss(preds, caps, dl, alphas, alpha_c=1.0)
        assert abs(float(ref) - float(got)) < 1e-12 * max(1.0, abs(float(ref)))
        ids = O.teacher_forced_ids(preds, dl)
        _, top = torch.max(preds, dim=2)
        assert [list(map(int, r)) for r in ids] == [top[j, :dl[j]].tolist() for j in range(B)]
