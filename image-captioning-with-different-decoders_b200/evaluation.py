"""Teacher-forced validation pass of the attention decoder — the decoder-side half of ``evaluate()``
(models/attention.py:454-567) for ANY batch size.

The reference validates with ``batch_size=1`` (:489-494): one decoder call, one loss and one ``loss.item()`` per image.
Rows of a batch are independent, so the same numbers come out of one batched call: per caption

    loss_j = mean_{t < decode_length_j} CE(scores[j, t], targets[j, t]) + mean_p (1 - sum_t alpha[j, t, p])^2     (:530-531)
    hypothesis_j = argmax_v scores[j, t, v] for t < decode_length_j, minus <start>/<end>/<pad>                     (:544-553)
    reference_j  = targets[j] minus <start>/<end>/<pad>, repeated once per target position                         (:536-541)

``evaluate_batch`` returns exactly the lists the reference appends to; ``evaluate`` runs a loader of such batches and
returns the reference's dictionary (``losses`` + whatever ``score_fn`` — e.g. the reference's ``metric.get_eval_score`` —
computes from references / hypotheses; the COCO caption scorers themselves are out of scope).
Dataset / encoder / DataLoader construction (:467-494) stay with the caller.
"""
import torch

from . import ops
from .vocabulary import END_TOKEN, PAD_TOKEN, START_TOKEN


def evaluate_batch(decoder, img_features, captions, caption_lengths, vocab):
    """-> (per_caption_losses list[float], hypotheses list[list[int]], references list[list[list[int]]], n_tokens list[int])
    for one batch; captions sorted by decreasing length or all equal (like every caller of the decoder)."""
    with torch.no_grad():
        scores, captions_sorted, decode_lengths, alphas = decoder(img_features, captions, caption_lengths)   # :522
        targets = captions_sorted[:, 1:]                                                                      # :523
        B, T, V = scores.shape
        dl = torch.tensor(decode_lengths, device=scores.device)
        active = torch.arange(T, device=scores.device)[None, :] < dl[:, None]                 # the rows pack_padded_sequence keeps
        tgt = torch.where(active, targets[:, :T], torch.full_like(targets[:, :T], -1))
        row_loss, _ = ops.cross_entropy_fwd(scores.reshape(B * T, V), tgt.reshape(-1).contiguous())      # 0 on dropped rows
        ce = row_loss.view(B, T).sum(dim=1) / dl.clamp_min(1).float()                                      # :529-530 per image
        reg = ((1.0 - alphas.sum(dim=1)) ** 2).mean(dim=1)                                                 # :531 per image
        losses = (ce + reg).tolist()
        _, preds = torch.max(scores, dim=2)                                                                # :544
        preds = preds.tolist()
        tlist = targets.tolist()
    special = [vocab(START_TOKEN), vocab(END_TOKEN), vocab(PAD_TOKEN)]
    references, hypotheses = [], []
    for j in range(B):
        row = tlist[j][:caption_lengths[j] - 1]             # the reference's batch of ONE carries no padding (:481-483)
        cleaned = [w for w in row if w not in special]                                                     # :538-539
        references.append([cleaned for _ in row])                                                          # :540 (sic: one copy per position)
        hypotheses.append([w for w in preds[j][:decode_lengths[j]] if w not in special])                   # :548-551
    return losses, hypotheses, references, list(decode_lengths)


def evaluate(device, args, encoder, decoder, val_loader, vocab, score_fn=None, verbose=False):
    """The loop of models/attention.py:454-567 over ``val_loader`` batches ``(imgs, captions, caption_lengths)`` of any size.
    Returns ``score_fn(references, hypotheses)`` (default: an empty dict) plus ``losses`` (one entry per caption, like the
    reference's batch-size-1 loop) and ``avg_loss`` (token-weighted running average, the AccumulatingMetric of :509, :532)."""
    decoder.eval()
    if hasattr(encoder, "eval"):
        encoder.eval()
    references, hypotheses, losses = [], [], []
    wsum, wcnt = 0.0, 0
    with torch.no_grad():
        for batch_idx, (imgs, captions, caption_lengths) in enumerate(val_loader):
            feats = encoder(imgs.to(device))                                                               # :520
            l, h, r, n = evaluate_batch(decoder, feats, captions.to(device), caption_lengths, vocab)
            losses.extend(l); hypotheses.extend(h); references.extend(r)
            wsum += sum(a * b for a, b in zip(l, n)); wcnt += sum(n)                                       # :532
            assert len(hypotheses) == len(references)
            if verbose and batch_idx % max(getattr(args, "print_freq", 1), 1) == 0:
                print(f'Batch {batch_idx + 1}, Loss {wsum / max(wcnt, 1):.4f}')
    metrics = score_fn(references, hypotheses) if score_fn is not None else {}
    metrics['losses'] = losses                                                                             # :562
    metrics['avg_loss'] = wsum / max(wcnt, 1)
    return metrics
