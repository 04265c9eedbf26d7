"""Drop-in ``BaselineDecoderParams`` / ``BaselineDecoder`` backed by libicd_b200.so.

Same constructor, attributes, ``state_dict`` keys (``embedding.weight``, ``lstm.weight_ih_l0``, ``lstm.weight_hh_l0``,
``lstm.bias_ih_l0``, ``lstm.bias_hh_l0``, ``linear.weight``, ``linear.bias``), default torch initialisation and
forward contract as the reference (models/baseline.py:19-111).  ``self.lstm`` / ``self.linear`` / ``self.embedding``
are parameter containers only; ``forward`` runs the CUDA kernels through the C ABI.
"""
import ctypes

import torch
import torch.nn as nn

from .. import _lib, ops
from .._lib import check, fill, lib, stream_ptr


class BaselineDecoderParams:
    hidden_size = 512
    embed_size = 512  # Use 300 if glove.
    vocab_size = None  # Must override.


class BaselineDecoder(nn.Module):
    precision = "fp32"      # class-level default (instances unpickled from a reference checkpoint do not carry it)

    def __init__(self, params):
        super().__init__()

        assert isinstance(params, BaselineDecoderParams)
        assert params.vocab_size is not None

        self.embed_size = params.embed_size
        self.hidden_size = params.hidden_size
        self.embedding = nn.Embedding(params.vocab_size, params.embed_size)          # :43
        self.lstm = nn.LSTM(input_size=params.embed_size, hidden_size=params.hidden_size, num_layers=1,
                            bias=True, batch_first=True, dropout=0, bidirectional=False)   # :47-53
        self.linear = nn.Linear(params.hidden_size, params.vocab_size)                # :57
        self.precision = "fp32"

    def load_pretrained_embeddins(self, embeddings):
        """(:59-66)"""
        self.embedding.weight = nn.Parameter(embeddings)

    def fine_tune_embeddings(self, on=True):
        """(:68-79)"""
        for param in self.embedding.parameters():
            param.requires_grad = on

    def forward(self, img_features, captions):
        """(B,E), (B,L) int64 -> caption scores (B, L, vocab_size)   (:81-111)"""
        return _BaselineDecoderFn.apply(img_features, captions, self.embedding.weight,
                                        self.lstm.weight_ih_l0, self.lstm.weight_hh_l0,
                                        self.lstm.bias_ih_l0, self.lstm.bias_hh_l0,
                                        self.linear.weight, self.linear.bias, self.precision)


class _BaselineDecoderFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, img, captions, emb_w, w_ih, w_hh, b_ih, b_hh, lin_w, lin_b, precision):
        if not img.is_cuda:
            raise _lib.IcdError("BaselineDecoder.forward needs CUDA tensors; there is no CPU fallback")
        dev = img.device
        _lib.remember_versions(ctx, [("embedding.weight", emb_w), ("lstm.weight_ih_l0", w_ih), ("lstm.weight_hh_l0", w_hh),
                                     ("lstm.bias_ih_l0", b_ih), ("lstm.bias_hh_l0", b_hh), ("linear.weight", lin_w),
                                     ("linear.bias", lin_b), ("img_features", img)])
        img = img.contiguous().float()                                               # :101 .float()
        captions = captions.contiguous()
        assert captions.dtype == torch.int64
        B, L = captions.shape
        E, H, V = emb_w.shape[1], w_hh.shape[1], lin_w.shape[0]
        assert img.shape == (B, E)
        emb_is_f64 = emb_w.dtype == torch.float64
        ws = [w.contiguous() for w in (w_ih, w_hh, b_ih, b_hh, lin_w, lin_b)]
        emb_c = emb_w.contiguous()           # kept in ctx.keep (a non-contiguous table must not leave a dangling pointer)
        f32 = dict(device=dev, dtype=torch.float32)
        bufs = dict(outputs=torch.empty(B, L, V, **f32), x=torch.empty(L, B, E, **f32),
                    xg=torch.empty(L, B, 4 * H, **f32), gates_act=torch.empty(L, B, 4 * H, **f32),
                    h_all=torch.empty(L + 1, B, H, **f32), c_all=torch.empty(L + 1, B, H, **f32),
                    hout=torch.empty(B, L, H, **f32), gates_pre=torch.empty(B, 4 * H, **f32))
        d = _lib.BaseDesc()
        fill(d, B=B, L=L, E=E, H=H, V=V, precision=ops.precision_id(precision), emb_is_f64=int(emb_is_f64),
             img_features=img, captions=captions, emb_w=emb_c, w_ih=ws[0], w_hh=ws[1], b_ih=ws[2],
             b_hh=ws[3], lin_w=ws[4], lin_b=ws[5], **bufs)
        need = int(lib().icd_baseline_decoder_ws_bytes(ctypes.byref(d)))
        tc_ws = torch.empty(need, device=dev, dtype=torch.uint8) if need else None
        fill(d, tc_ws=tc_ws, tc_ws_bytes=need)
        bufs["tc_ws"] = tc_ws
        check(lib().icd_baseline_decoder_fwd(ctypes.byref(d), stream_ptr()), "icd_baseline_decoder_fwd")
        outputs = bufs.pop("outputs")          # not kept in ctx: it carries this node as grad_fn (reference cycle)
        ctx.desc = d
        ctx.keep = (img, captions, emb_c, ws, bufs)
        ctx.dims = (B, L, E, H, V)
        return outputs

    @staticmethod
    def backward(ctx, d_out):
        B, L, E, H, V = ctx.dims
        if ctx.keep is None:
            raise RuntimeError("icd_b200: BaselineDecoder backward called a second time; its saved activations were "
                               "released after the first backward")
        _lib.check_versions(ctx, "BaselineDecoder")
        img, captions, emb_w, ws, bufs = ctx.keep
        dev = img.device
        f32 = dict(device=dev, dtype=torch.float32)
        d_out = d_out.contiguous().float()
        want_img, want_emb = ctx.needs_input_grad[0], ctx.needs_input_grad[2]
        g = dict(d_w_ih=torch.empty(4 * H, E, **f32), d_w_hh=torch.empty(4 * H, H, **f32),
                 d_b=torch.empty(4 * H, **f32), d_lin_w=torch.empty(V, H, **f32), d_lin_b=torch.empty(V, **f32),
                 d_emb_w=(torch.empty(V, E, device=dev, dtype=emb_w.dtype) if want_emb else None),
                 d_img_features=(torch.empty(B, E, **f32) if want_img else None))
        scratch = dict(d_hout=torch.empty(B, L, H, **f32), dg=torch.empty(L, B, 4 * H, **f32),
                       dh=torch.empty(B, H, **f32), dc=torch.empty(B, H, **f32),
                       d_x=(torch.empty(L, B, E, **f32) if (want_img or want_emb) else None))
        d = ctx.desc
        fill(d, d_outputs=d_out, **g, **scratch)
        check(lib().icd_baseline_decoder_bwd(ctypes.byref(d), stream_ptr()), "icd_baseline_decoder_bwd")
        ctx.keep = None          # release the saved activations now, not when the loss tensor dies
        ctx.desc = None
        return (g["d_img_features"], None, g["d_emb_w"], g["d_w_ih"], g["d_w_hh"], g["d_b"], g["d_b"].clone(),
                g["d_lin_w"], g["d_lin_b"], None)
