"""Synthetic inputs of the shapes the reference's decoders consume.

COCO, GloVe and the pretrained ResNet-101 are unavailable offline, so the encoder output is
replaced by seeded post-ReLU-like features and captions by seeded token ids.
"""
import torch


def features(batch, seed=1234, size=14, channels=2048, dtype=torch.float32):
    """(B, 14, 14, 2048) channel-last, ~50 % zeros — what EncoderAttention returns
    (models/encoder.py:107-110) up to memory layout."""
    g = torch.Generator().manual_seed(seed)
    return torch.randn(batch, size, size, channels, generator=g, dtype=dtype).clamp_min_(0)


def captions(batch, vocab_size, max_len=25, seed=1234, lengths=None):
    """int64 (B, max_len): <start>, words in 1..V-4, <end> at position len-1, <pad>=0 after.

    lengths: None => all rows max_len (the reference's own training regime);
             'ragged' => U{8..max_len} sorted descending; or an explicit list."""
    g = torch.Generator().manual_seed(seed + 1)
    start, end = vocab_size - 3, vocab_size - 2
    if lengths is None:
        lens = [max_len] * batch
    elif isinstance(lengths, str) and lengths == 'ragged':
        lo = min(8, max_len)
        lens = sorted(torch.randint(lo, max_len + 1, (batch,), generator=g).tolist(), reverse=True)
        lens[0] = max_len
    else:
        lens = list(lengths)
        assert len(lens) == batch
    hi = max(vocab_size - 3, 2)
    caps = torch.randint(1, hi, (batch, max_len), generator=g, dtype=torch.int64)
    caps[:, 0] = start
    for b, l in enumerate(lens):
        caps[b, l - 1] = end
        caps[b, l:] = 0
    return caps, lens


def glove_like_table(vocab_size, embed_size=300, seed=7):
    """fp64 (V, 300) table like load_glove_vectors() builds (embed.py:56,66-67): N(0, 0.6)."""
    g = torch.Generator().manual_seed(seed)
    return torch.randn(vocab_size, embed_size, generator=g, dtype=torch.float64) * 0.6
