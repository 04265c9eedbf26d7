"""Loss glue of the reference train loops as fused kernels.

``attention_caption_loss`` == models/attention.py:401-414:

    targets = captions[:, 1:]
    scores  = pack_padded_sequence(scores,  decode_lengths, batch_first=True).data
    targets = pack_padded_sequence(targets, decode_lengths, batch_first=True).data
    loss = CrossEntropyLoss()(scores, targets) + ((alpha_c - alphas.sum(dim=1)) ** 2).mean()

The mean cross-entropy over the packed rows does not depend on their order, so the 466 MB packed copy and the
separate log-softmax / NLL / softmax-backward passes are replaced by two streaming kernels (icd_cross_entropy_fwd:
per-row loss + log-sum-exp; icd_cross_entropy_bwd: gradient, scaled by the upstream gradient read on the device, plus
a bf16 copy for the tensor-core tier); rows with t >= batch_size_t carry target -1 and are skipped
(they are the rows pack_padded_sequence drops).  ``baseline_caption_loss`` == models/baseline.py:194-195,224-225
(CrossEntropyLoss(ignore_index=<pad>) over all (b, t)).
The doubly-stochastic regulariser over the (B,T,196) alphas is one forward and one backward kernel (``alpha_regulariser``).
"""
import numpy as np
import torch

from . import ops


# bf16 side-channel between the fused loss and the bf16 tier of the decoder backward: the loss backward already streams
# every logit once, so it also emits the bf16 copy of d(loss)/d(logits) that the vocabulary-layer backward contractions
# consume — the decoder backward then skips its own fp32 -> bf16 pass over the (B*T, V) gradient.  Keyed by the device
# address of the fp32 gradient, consumed once, and only honoured if that tensor has not been modified in place since
# (autograd accumulating a second gradient into it bumps its version counter).
_BF16_SIDECAR = {}


def take_bf16_sidecar(grad):
    """-> bf16 tensor (R, ld16) matching the fp32 gradient ``grad`` if the fused loss produced one, else None.
    Raises if ``grad`` is a HOLLOW fp32 gradient (``bf16_grad_only=True``) whose bf16 copy can no longer be trusted."""
    ent = _BF16_SIDECAR.pop(grad.data_ptr(), None)
    orphan_hollow = any(e[3] and e[1].numel() == grad.numel() for e in _BF16_SIDECAR.values())
    _BF16_SIDECAR.clear()
    if ent is None:
        if orphan_hollow:            # autograd summed the hollow gradient with another one into a new tensor
            raise RuntimeError("attention_caption_loss(bf16_grad_only=True): the logit gradient reaching the decoder is "
                               "not the one the loss produced (predictions feed something else too); use "
                               "bf16_grad_only=False")
        return None
    d16, d32, version, hollow = ent
    if d32.numel() != grad.numel() or d32._version != version or grad._version != version:
        if hollow:
            raise RuntimeError("attention_caption_loss(bf16_grad_only=True): the logit gradient was modified before it "
                               "reached the decoder (predictions feed something else too); use bf16_grad_only=False")
        return None
    return d16


class _FusedCE(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits2d, targets, n_valid, want_bf16, bf16_only=False):
        row_loss, lse = ops.cross_entropy_fwd(logits2d, targets)
        ctx.save_for_backward(logits2d, targets, lse)
        ctx.n_valid, ctx.want_bf16, ctx.bf16_only = n_valid, want_bf16, bool(bf16_only and want_bf16)
        return row_loss.sum() / n_valid

    @staticmethod
    def backward(ctx, g):
        logits2d, targets, lse = ctx.saved_tensors
        g = g.reshape(1).float().contiguous()
        d_logits, d16 = ops.cross_entropy_bwd(logits2d, targets, lse, 1.0 / ctx.n_valid, upstream=g,
                                              want_bf16=ctx.want_bf16, want_fp32=not ctx.bf16_only)
        _BF16_SIDECAR.clear()
        if d16 is not None:
            _BF16_SIDECAR[d_logits.data_ptr()] = (d16, d_logits, d_logits._version, ctx.bf16_only)
        return d_logits, None, None, None, None


class _AlphaReg(torch.autograd.Function):
    """((alpha_c - alphas.sum(dim=1)) ** 2).mean() as one forward and one backward kernel (the torch expression is eight)."""

    @staticmethod
    def forward(ctx, alphas, alpha_c):
        reg, resid = ops.alpha_regulariser_fwd(alphas, alpha_c)
        ctx.save_for_backward(resid)
        ctx.T = alphas.shape[1]
        return reg

    @staticmethod
    def backward(ctx, g):
        (resid,) = ctx.saved_tensors
        return ops.alpha_regulariser_bwd(resid, ctx.T, upstream=g.reshape(1).float().contiguous()), None


def alpha_regulariser(alphas, alpha_c=1.0):
    """Doubly stochastic attention regulariser of models/attention.py:413-414."""
    return _AlphaReg.apply(alphas.float().contiguous(), float(alpha_c))        # CUDA only, like every op of this package


def packed_targets(encoded_captions, decode_lengths, T, row_valid=None):
    """targets[b, t] = captions[b, t+1] where row b is active at step t (b < batch_size_t), else -1.
    ``row_valid``: the (B*T) uint8 activity mask the decoder forward already built on the device (no host->device
    copy, no sync); rebuilt from ``decode_lengths`` when absent."""
    B = encoded_captions.shape[0]
    dev = encoded_captions.device
    if row_valid is not None:
        active = row_valid.view(B, T).bool()
    else:
        dl = np.asarray(decode_lengths, dtype=np.int64)
        bt = (dl[None, :] > np.arange(T)[:, None]).sum(axis=1)                       # batch_size_t
        active = torch.from_numpy(np.arange(B)[:, None] < bt[None, :]).to(dev, non_blocking=True)
    tgt = encoded_captions[:, 1:T + 1]
    return torch.where(active, tgt, torch.full_like(tgt, -1)), int(sum(decode_lengths))


def attention_caption_loss(predictions, encoded_captions, decode_lengths, alphas, alpha_c=1.0, bf16_grad_only=False):
    """bf16_grad_only (bf16 tier only): the loss backward writes ONLY the bf16 copy of d(loss)/d(logits) that the decoder
    backward consumes and leaves the 466 MB fp32 gradient tensor hollow.  Valid when ``predictions`` feeds nothing but this
    loss in the autograd graph — the reference train loop (models/attention.py:401-414; its top-5 accuracy is computed
    without a graph).  If the gradient is modified on its way to the decoder the backward raises instead of reading it."""
    B, T, V = predictions.shape
    tgt, n_valid = packed_targets(encoded_captions, decode_lengths, T, getattr(predictions, "_icd_row_valid", None))
    want_bf16 = bool(getattr(predictions, "_icd_bf16_tier", False))
    ce = _FusedCE.apply(predictions.reshape(B * T, V), tgt.reshape(-1).contiguous(), n_valid, want_bf16, bf16_grad_only)
    return ce + alpha_regulariser(alphas, alpha_c)


def baseline_caption_loss(outputs, captions, pad_id=0):
    B, L, V = outputs.shape
    tgt = torch.where(captions == pad_id, torch.full_like(captions, -1), captions).reshape(-1).contiguous()
    n_valid = int((tgt >= 0).sum().item())
    return _FusedCE.apply(outputs.reshape(B * L, V), tgt, max(n_valid, 1), False)
