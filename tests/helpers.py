"""Shared test helpers: seeded cases, digests, comparison utilities.

Every golden case is fully determined by a small dict of integers (dims + seeds): weights come from the module
constructors under ``torch.manual_seed`` (bit-identical between the reference and icd_b200 — asserted when the
goldens are generated and re-checked by checksum in the tests), inputs from ``icd_b200.synthetic``.
"""
import hashlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from icd_b200 import synthetic  # noqa: E402
from icd_b200.vocabulary import synthetic_vocab  # noqa: E402

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")

# name -> case description.  "full": whole tensors stored; "digest": norms + sampled entries stored.
ATT_CASES = {
    "att_small_ragged": dict(B=5, V=97, A=48, D=32, E=24, max_len=9, lengths=[9, 9, 7, 4, 2], wseed=0, iseed=11,
                             dropout=0.0, train=False, store="full", fine_tune_embedding=True),
    "att_small_unsorted": dict(B=4, V=97, A=48, D=32, E=24, max_len=8, lengths=[5, 8, 3, 8], wseed=1, iseed=12,
                               dropout=0.0, train=False, store="full", fine_tune_embedding=True, loss=False),
    "att_small_dropout": dict(B=4, V=97, A=48, D=32, E=24, max_len=7, lengths=[7, 7, 7, 7], wseed=2, iseed=13,
                              dropout=0.5, train=True, mask_seed=5, store="full", fine_tune_embedding=False),
    "att_cfg1": dict(B=4, V=9490, A=512, D=512, E=512, max_len=25, lengths=[25] * 4, wseed=0, iseed=1234,
                     dropout=0.0, train=False, store="digest", fine_tune_embedding=False),
    "att_glove": dict(B=3, V=211, A=64, D=64, E=300, max_len=10, lengths=[10, 8, 6], wseed=3, iseed=14,
                      dropout=0.0, train=False, store="digest", fine_tune_embedding=True, glove=True),
}
BASE_CASES = {
    "base_small": dict(B=6, V=97, E=32, H=32, L=9, wseed=0, iseed=21, store="full", lengths="ragged"),
    "base_cfg2": dict(B=128, V=9490, E=512, H=512, L=25, wseed=0, iseed=1234, store="digest", lengths=None),
}
BEAM_CASES = {
    # SURVEY.md 8c beam fixture recipe: embedding *= 30, fc.weight[END] *= gain, fc.bias[END] = bias
    "beam_cfg": dict(V=9490, A=512, D=512, E=512, k=5, wseed=0, iseed=77, n_img=6, emb_scale=30.0,
                     end_gain=30.0, end_bias=-4.0),
    "beam_small": dict(V=131, A=48, D=32, E=24, k=3, wseed=4, iseed=78, n_img=8, emb_scale=30.0,
                       end_gain=30.0, end_bias=-2.0),
}


def att_inputs(case):
    enc = synthetic.features(case["B"], seed=case["iseed"])
    caps, lens = synthetic.captions(case["B"], case["V"], max_len=case["max_len"], seed=case["iseed"],
                                    lengths=case["lengths"])
    return enc, caps, lens


def dropout_masks_like_reference(case, decode_lengths, D):
    """The keep-masks nn.Dropout draws inside the reference forward on CPU, in call order
    (one (batch_size_t, D) mask per step, models/attention.py:279), reproduced from the seed."""
    torch.manual_seed(case["mask_seed"])
    masks = []
    for t in range(max(decode_lengths)):
        bt = sum(l > t for l in decode_lengths)
        masks.append((torch.nn.functional.dropout(torch.ones(bt, D), case["dropout"], True) > 0).float())
    return masks


def stack_masks(masks, B, D):
    T = len(masks)
    out = torch.zeros(T, B, D, dtype=torch.uint8)
    for t, m in enumerate(masks):
        out[t, :m.shape[0]] = m.to(torch.uint8)
    return out


def build_attention_module(case, module_cls, params_cls, vocab, device="cpu"):
    p = params_cls()
    p.attention_dim, p.decoder_dim, p.embed_size = case["A"], case["D"], case["E"]
    p.dropout = case["dropout"]
    p.vocab = vocab
    torch.manual_seed(case["wseed"])
    dec = module_cls(torch.device(device), p)
    if case.get("glove"):
        dec.load_pretrained_embeddins(synthetic.glove_like_table(case["V"], case["E"]))
    dec.fine_tune_embeddings(case["fine_tune_embedding"])
    dec.train(case["train"])
    return dec


def build_baseline_module(case, module_cls, params_cls):
    p = params_cls()
    p.hidden_size, p.embed_size, p.vocab_size = case["H"], case["E"], case["V"]
    torch.manual_seed(case["wseed"])
    return module_cls(p)


def base_inputs(case):
    g = torch.Generator().manual_seed(case["iseed"])
    img = torch.randn(case["B"], case["E"], generator=g)
    caps, lens = synthetic.captions(case["B"], case["V"], max_len=case["L"], seed=case["iseed"],
                                    lengths=case["lengths"])
    return img, caps, lens


def apply_beam_recipe(dec, case):
    V = case["V"]
    with torch.no_grad():
        dec.embedding.weight *= case["emb_scale"]
        dec.fc.weight[V - 2] *= case["end_gain"]
        dec.fc.bias[V - 2] = case["end_bias"]


def beam_features(case):
    g = torch.Generator().manual_seed(case["iseed"])
    return torch.randn(case["n_img"], 14, 14, 2048, generator=g).abs()


# --------------------------------------------------------------------------------------------
def checksum(t):
    return hashlib.sha256(t.detach().cpu().contiguous().numpy().tobytes()).hexdigest()[:16]


def state_checksums(module):
    return {k: checksum(v) for k, v in module.state_dict().items()}


def sample_index(name, numel, n=48):
    seed = int(hashlib.sha256(name.encode()).hexdigest()[:8], 16)
    g = torch.Generator().manual_seed(seed)
    return torch.randint(0, numel, (min(n, numel),), generator=g)


def digest(name, t):
    t = t.detach().cpu().double().reshape(-1)
    idx = sample_index(name, t.numel())
    return dict(norm=float(t.norm()), sum=float(t.sum()), absmax=float(t.abs().max()),
                samples=t[idx].numpy().astype(np.float64))


def rel_err(a, b):
    a = a.detach().cpu().double()
    b = b.detach().cpu().double()
    den = float(b.norm())
    return float((a - b).norm()) / (den if den > 0 else 1.0)


def assert_close_norm(a, b, tol, what, atol=0.0):
    """Norm-wise relative comparison (SURVEY.md 8c caution (i)), with an absolute floor for tensors whose
    true value is ~0 (full_att.bias.grad)."""
    a = a.detach().cpu().double()
    b = b.detach().cpu().double()
    assert a.shape == b.shape, "%s: shape %s vs %s" % (what, tuple(a.shape), tuple(b.shape))
    err = float((a - b).norm())
    den = float(b.norm())
    assert err <= tol * den + atol, "%s: |a-b|=%.3e > %.1e*|b|(%.3e) + %.1e" % (what, err, tol, den, atol)


def assert_digest_close(name, t, dg, tol, atol=0.0):
    """Compare a tensor with a stored digest: norm, sum and the sampled entries."""
    t64 = t.detach().cpu().double().reshape(-1)
    idx = sample_index(name, t64.numel())
    norm = float(dg["norm"])
    assert abs(float(t64.norm()) - norm) <= tol * norm + atol, "%s: norm %.6e vs %.6e" % (name, float(t64.norm()), norm)
    s = torch.from_numpy(np.asarray(dg["samples"], dtype=np.float64))
    # sampled entries: error measured against the tensor's RMS scale, not the individual entry
    rms = norm / max(1.0, t64.numel() ** 0.5)
    err = float((t64[idx] - s).abs().max())
    assert err <= 20 * tol * rms + atol + tol * float(s.abs().max()), \
        "%s: sampled entries differ by %.3e (rms %.3e)" % (name, err, rms)
