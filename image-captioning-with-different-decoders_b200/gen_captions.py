"""Caption generation with beam search — drop-in for gen_captions.py:16-131, plus a batched entry point.

``attention_caption_image_beam_search(device, args, img, encoder, decoder, vocab)`` keeps the reference signature
and return tuple ``(seq, alphas, Caption_End)``; ``beam_search_batched`` decodes many images at once with the
device-side engine of ``icd_beam_search`` (csrc/beam.cu): no per-step host synchronisation, features indexed per
image, ``enc_att`` computed once per image.
"""
import ctypes

import torch

from . import _lib, ops
from ._lib import check, fill, lib, stream_ptr
from .vocabulary import END_TOKEN, START_TOKEN


def beam_search_batched(decoder, features, beam_size, start_id, end_id, max_steps=50,
                        want_alphas=True, want_trace=False, chunk=None, precision="fp32x3"):
    """features (n_img, 14, 14, C) or (n_img, P, C) CUDA fp32 -> dict with
         len    (n_img,) int32  caption length incl. <start>/<end>; 0 = no beam completed (reference failure tuple)
         seq    (n_img, max_steps+2) int32
         score  (n_img,) float32 raw summed log-prob of the winner (gen_captions.py:127)
         alpha  (n_img, max_steps+2, P) float32, frame 0 all ones (gen_captions.py:54)      [want_alphas]
         trace  (max_steps+1, n_img, k) int32 next-word ids per step, -1 = empty slot        [want_trace]
    The loop body runs for step = 1 .. max_steps+1, like ``step > 50`` at gen_captions.py:119 for max_steps=50.
    Caption generation needs fp32-grade logits for identical captions: ``precision`` is "fp32x3" (default:
    3-term bf16 split on the tcgen05 tensor cores, ~6e-6 relative) or "fp32" (fp32 FMA kernel); never "bf16"."""
    if precision not in ("fp32", "fp32x3"):
        raise ValueError("beam search runs in fp32-grade arithmetic only: precision must be 'fp32' or 'fp32x3'")
    if not features.is_cuda:
        raise _lib.IcdError("beam_search_batched needs CUDA tensors; there is no CPU fallback")
    n_img = features.shape[0]
    C = features.shape[-1]
    enc = features.reshape(n_img, -1, C).contiguous().float()
    P = enc.shape[1]
    k = int(beam_size)
    dev = enc.device
    a = decoder.attention
    A, D, E, V = a.enc_att.weight.shape[0], decoder.decode_step.weight_hh.shape[1], \
        decoder.embedding.weight.shape[1], decoder.fc.weight.shape[0]
    if chunk is None:
        chunk = max(1, min(n_img, 65535 // k))
    S2 = max_steps + 2
    out = dict(len=torch.empty(n_img, device=dev, dtype=torch.int32),
               seq=torch.zeros(n_img, S2, device=dev, dtype=torch.int32),
               score=torch.empty(n_img, device=dev, dtype=torch.float32))
    if want_alphas:
        out["alpha"] = torch.zeros(n_img, S2, P, device=dev, dtype=torch.float32)
    traces = []
    emb_w = decoder.embedding.weight
    ws_cache = None
    for i0 in range(0, n_img, chunk):
        n = min(chunk, n_img - i0)
        d = _lib.BeamDesc()
        trace = torch.empty(max_steps + 1, n, k, device=dev, dtype=torch.int32) if want_trace else None
        fill(d, n_img=n, k=k, max_steps=max_steps, P=P, C=C, A=A, D=D, E=E, V=V,
             precision=ops.precision_id(precision), emb_is_f64=int(emb_w.dtype == torch.float64),
             start_id=start_id, end_id=end_id, enc=enc[i0:i0 + n],
             enc_att_w=a.enc_att.weight, enc_att_b=a.enc_att.bias, dec_att_w=a.dec_att.weight,
             dec_att_b=a.dec_att.bias, full_att_w=a.full_att.weight, full_att_b=a.full_att.bias,
             w_ih=decoder.decode_step.weight_ih, w_hh=decoder.decode_step.weight_hh,
             b_ih=decoder.decode_step.bias_ih, b_hh=decoder.decode_step.bias_hh,
             h_lin_w=decoder.h_lin.weight, h_lin_b=decoder.h_lin.bias, c_lin_w=decoder.c_lin.weight,
             c_lin_b=decoder.c_lin.bias, f_beta_w=decoder.f_beta.weight, f_beta_b=decoder.f_beta.bias,
             fc_w=decoder.fc.weight, fc_b=decoder.fc.bias, emb_w=emb_w,
             out_len=out["len"][i0:i0 + n], out_seq=out["seq"][i0:i0 + n], out_score=out["score"][i0:i0 + n],
             out_alpha=(out["alpha"][i0:i0 + n] if want_alphas else None), trace_words=trace)
        need = int(lib().icd_beam_search_ws_bytes(ctypes.byref(d)))
        if ws_cache is None or ws_cache.numel() < need:
            ws_cache = torch.empty(need, device=dev, dtype=torch.uint8)
        d.ws = ws_cache.data_ptr()
        d.ws_bytes = need
        check(lib().icd_beam_search(ctypes.byref(d), stream_ptr()), "icd_beam_search")
        if want_trace:
            traces.append(trace)
    if want_trace:
        out["trace"] = torch.cat(traces, dim=1)
    return out


def beam_search_sharded(decoder, features_local, beam_size, start_id, end_id, max_steps=50, group=None, dst=0,
                        precision="fp32x3"):
    """BASELINE.json configs[4] / SURVEY.md 8e row 3: images are independent units, so every rank decodes its own shard
    with ``beam_search_batched`` (no collective during the decode) and ONE final gather brings caption lengths, token ids
    and scores to rank ``dst``.  All ranks must hold the same number of images.  Works unchanged in a single process.
    -> on ``dst``: dict(len (world*n,), seq (world*n, max_steps+2), score (world*n,)) in rank order (CUDA tensors);
       elsewhere: None."""
    import torch.distributed as dist
    res = beam_search_batched(decoder, features_local, beam_size, start_id, end_id, max_steps=max_steps,
                              want_alphas=False, want_trace=False, precision=precision)
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return {k: res[k] for k in ("len", "seq", "score")}
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    n, S2 = res["seq"].shape
    # one packed int32 message per rank: [len | score bits | seq]
    packed = torch.cat([res["len"].view(n, 1), res["score"].view(torch.int32).view(n, 1), res["seq"]], dim=1).contiguous()
    parts = [torch.empty_like(packed) for _ in range(world)] if rank == dst else None
    dist.gather(packed, parts, dst=dst, group=group)
    if rank != dst:
        return None
    allp = torch.cat(parts, dim=0)
    return {"len": allp[:, 0].contiguous(), "score": allp[:, 1].contiguous().view(torch.float32), "seq": allp[:, 2:].contiguous()}


def attention_caption_image_beam_search(device, args, img, encoder, decoder, vocab):
    """Reads an image and captions it with beam search (gen_captions.py:16-131).

    Returns ``(seq, alphas, Caption_End)``: ``seq`` a python list of token ids including <start> and <end>,
    ``alphas`` a nested list (len(seq) x 14 x 14, first frame all ones), or ``([start, end], [], False)`` when no
    beam emitted <end> within the 51 steps the reference allows."""
    k = args.beam_size
    vocab_size = len(vocab)                                                         # :35
    encoder_out = encoder(img)                                                      # :37 (1, s, s, C)
    enc_image_size = encoder_out.size(1)
    start_id, end_id = vocab(START_TOKEN), vocab(END_TOKEN)
    with torch.no_grad():
        res = beam_search_batched(decoder, encoder_out, k, start_id, end_id, max_steps=50,
                                  want_alphas=True, want_trace=True)
    n = int(res["len"][0].item())
    trace = res["trace"][:, 0, :].tolist()
    for words in trace:                                                             # :91 per-step print
        live = [w for w in words if w >= 0]
        if not live:
            break
        print([vocab.i2w[w] for w in live])
    if n == 0:                                                                      # :123-125
        return [start_id, end_id], [], False
    seq = res["seq"][0, :n].tolist()
    alphas = res["alpha"][0, :n].view(n, enc_image_size, enc_image_size).tolist()
    assert vocab_size == decoder.fc.weight.shape[0]
    return seq, alphas, True
