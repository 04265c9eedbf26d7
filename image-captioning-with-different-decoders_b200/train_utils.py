"""Optimiser glue of the reference train loops, B200-native.

``clip_gradient`` keeps the reference signature/semantics (train_utils.py:2-12: element-wise clamp of every
gradient to +-grad_clip).  ``FusedClipAdam`` is the fused equivalent of

    clip_gradient(optimizer, grad_clip); optimizer.step()      # models/attention.py:423-428

for ``torch.optim.Adam(params, lr)`` with default betas/eps: one kernel launch per parameter over flat fp32 state
(icd_clip_adam_step), with an optional ``grad_scale`` (1/world_size after the data-parallel all-reduce).
"""
import torch

from . import ops


def clip_gradient(optimizer, grad_clip):
    """Reference semantics (train_utils.py:2-12)."""
    for group in optimizer.param_groups:
        for param in group['params']:
            if param.grad is not None:
                param.grad.data.clamp_(-grad_clip, grad_clip)


class FusedClipAdam:
    """clamp(+-grad_clip) + Adam(lr, betas=(0.9, 0.999), eps=1e-8) in one kernel per parameter tensor.

    fp32 parameters run on the CUDA kernel; float64 parameters (the GloVe embedding table)
    are few and rare and are stepped by ``torch.optim.Adam`` after the same clamp."""

    def __init__(self, params, lr=1e-4, betas=(0.9, 0.999), eps=1e-8, grad_clip=5.0):
        self.params = [p for p in params if p.requires_grad]
        self.lr, self.betas, self.eps, self.grad_clip = lr, betas, eps, grad_clip
        self.step_count = 0
        self.state = {}
        self._f64 = [p for p in self.params if p.dtype != torch.float32]
        self._f64_opt = torch.optim.Adam(self._f64, lr=lr, betas=betas, eps=eps) if self._f64 else None

    def zero_grad(self):
        for p in self.params:
            p.grad = None

    @torch.no_grad()
    def step(self, grad_scale=1.0):
        self.step_count += 1
        for p in self.params:
            if p.grad is None or p.dtype != torch.float32:
                continue
            st = self.state.get(p)
            if st is None:
                st = (torch.zeros_like(p, memory_format=torch.contiguous_format),
                      torch.zeros_like(p, memory_format=torch.contiguous_format))
                self.state[p] = st
            g = p.grad if p.grad.is_contiguous() else p.grad.contiguous()
            ops.clip_adam_step(p.data, g, st[0], st[1], self.step_count, lr=self.lr, betas=self.betas,
                               eps=self.eps, grad_clip=self.grad_clip, grad_scale=grad_scale)
        if self._f64_opt is not None:
            for p in self._f64:
                if p.grad is not None:
                    p.grad.mul_(grad_scale).clamp_(-self.grad_clip, self.grad_clip)
            self._f64_opt.step()
