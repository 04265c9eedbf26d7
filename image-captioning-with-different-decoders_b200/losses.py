"""Loss glue of the reference train loops as fused kernels.

``attention_caption_loss`` == models/attention.py:401-414:

    targets = captions[:, 1:]
    scores  = pack_padded_sequence(scores,  decode_lengths, batch_first=True).data
    targets = pack_padded_sequence(targets, decode_lengths, batch_first=True).data
    loss = CrossEntropyLoss()(scores, targets) + ((alpha_c - alphas.sum(dim=1)) ** 2).mean()

The mean cross-entropy over the packed rows does not depend on their order, so the 466 MB packed copy and the
separate log-softmax / NLL / softmax-backward passes are replaced by ONE kernel (icd_cross_entropy_fwd_bwd) that
reads each logits row once and writes loss and gradient; rows with t >= batch_size_t carry target -1 and are skipped
(they are the rows pack_padded_sequence drops).  ``baseline_caption_loss`` == models/baseline.py:194-195,224-225
(CrossEntropyLoss(ignore_index=<pad>) over all (b, t)).
The doubly-stochastic regulariser touches only the (B,T,196) alphas and stays a torch expression.
"""
import torch

from . import ops


class _FusedCE(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits2d, targets, n_valid):
        row_loss, d_logits = ops.cross_entropy_fwd_bwd(logits2d, targets, 1.0 / n_valid, want_grad=True)
        ctx.save_for_backward(d_logits)
        return row_loss.sum() / n_valid

    @staticmethod
    def backward(ctx, g):
        (d_logits,) = ctx.saved_tensors
        return d_logits * g, None, None


def packed_targets(encoded_captions, decode_lengths, T):
    """targets[b, t] = captions[b, t+1] where row b is active at step t (b < batch_size_t), else -1."""
    B = encoded_captions.shape[0]
    bt = torch.tensor([sum(l > t for l in decode_lengths) for t in range(T)], device=encoded_captions.device)
    rows = torch.arange(B, device=encoded_captions.device).unsqueeze(1)
    active = rows < bt.unsqueeze(0)                                   # (B,T): first batch_size_t rows
    tgt = encoded_captions[:, 1:T + 1]
    return torch.where(active, tgt, torch.full_like(tgt, -1)), int(sum(decode_lengths))


def attention_caption_loss(predictions, encoded_captions, decode_lengths, alphas, alpha_c=1.0):
    B, T, V = predictions.shape
    tgt, n_valid = packed_targets(encoded_captions, decode_lengths, T)
    ce = _FusedCE.apply(predictions.reshape(B * T, V), tgt.reshape(-1).contiguous(), n_valid)
    return ce + ((alpha_c - alphas.sum(dim=1)) ** 2).mean()


def baseline_caption_loss(outputs, captions, pad_id=0):
    B, L, V = outputs.shape
    tgt = torch.where(captions == pad_id, torch.full_like(captions, -1), captions).reshape(-1).contiguous()
    n_valid = int((tgt >= 0).sum().item())
    return _FusedCE.apply(outputs.reshape(B * L, V), tgt, max(n_valid, 1))
