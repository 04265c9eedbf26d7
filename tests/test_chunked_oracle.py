"""CPU tier: the row-chunked fp64 oracle (tests/helpers.oracle_fp64_chunked, the checker of the timed-configuration parity
test) equals the one-shot oracle — loss and every parameter gradient — for equal and for ragged (descending) lengths,
with a train-mode dropout keep-mask."""
import pytest
import torch

import helpers as H
from oracle import decoders as O


@pytest.mark.parametrize("lengths", [None, [9, 9, 8, 8, 6, 5, 3, 2, 2, 1]])
def test_chunked_oracle_equals_one_shot(lengths):
    import icd_b200.models.attention as my_att
    from icd_b200.vocabulary import synthetic_vocab
    case = dict(B=10, V=61, A=16, D=8, E=12, max_len=9, lengths=lengths or [9] * 10, wseed=3, iseed=5, dropout=0.5,
                train=True, fine_tune_embedding=False)
    dec = H.build_attention_module(case, my_att.AttentionDecoder, my_att.AttentionDecoderParams,
                                   synthetic_vocab(case["V"]))
    sd = dec.state_dict()
    enc, caps, lens = H.att_inputs(case)
    dl = [l - 1 for l in lens]
    T = max(dl)
    keep = H.seeded_keep_mask(T, case["B"], case["D"], 0.5, seed=9)
    w = {k: v.detach().double().clone().requires_grad_(k != "embedding.weight") for k, v in sd.items()}
    masks = [keep[t, :sum(1 for l in dl if l > t)].double() for t in range(T)]
    p, _, d, a = O.attention_decoder_forward(w, enc.double(), caps, lens, dropout_p=0.5, dropout_masks=masks)
    loss = O.attention_loss(p, caps, d, a)
    loss.backward()
    r = H.oracle_fp64_chunked(sd, enc, caps, lens, keep_mask=keep, p=0.5, frozen=("embedding.weight",), chunk=3,
                              cuda_preds=p.detach().float(), cuda_alphas=a.detach().float())
    assert abs(r["loss"] - loss.item()) < 1e-12 * abs(loss.item())
    assert set(r["grads"]) == {k for k, v in w.items() if v.grad is not None}
    for k, g in r["grads"].items():
        H.assert_close_norm(g, w[k].grad, 1e-11, k, atol=1e-15)      # full_att.bias: true value 0, only noise
    assert r["err"]["predictions"] < 1e-6 and r["err"]["alphas"] < 1e-6
