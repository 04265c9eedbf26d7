"""Batch-sharded data parallelism for the decoder train step: one process per GPU, weights
replicated, ONE all-reduce(sum) per step over a flat fp32 gradient buffer (NCCL over NVLink on the GPU box, gloo in
the CPU tests), then clamp(+-grad_clip) + Adam fused in a single kernel over the flat parameter buffer
(train_utils.py:2-12 + models/attention.py:417-430).  The reference has no distributed code; with world_size == 1
this is exactly its single-process step.

The gradient is scaled by 1/world_size BEFORE the clamp so that N ranks x B captions reproduce the single-process
(N*B)-caption step when every rank holds the same number of packed tokens (equal-length batches).
"""
import torch
import torch.distributed as dist


class FlatParamBuffer:
    """Re-points every trainable fp32 parameter of ``module`` at a view of one flat buffer (and keeps a flat
    gradient buffer of the same layout).  Device-agnostic; float64 parameters (GloVe table) are listed separately."""

    def __init__(self, module):
        self.params = [p for p in module.parameters() if p.requires_grad and p.dtype == torch.float32]
        self.other = [p for p in module.parameters() if p.requires_grad and p.dtype != torch.float32]
        assert self.params, "no trainable fp32 parameters"
        dev = self.params[0].device
        n = sum(p.numel() for p in self.params)
        self.flat = torch.empty(n, device=dev, dtype=torch.float32)
        self.flat_grad = torch.zeros(n, device=dev, dtype=torch.float32)
        self.views, self.grad_views = [], []
        off = 0
        for p in self.params:
            k = p.numel()
            v = self.flat[off:off + k].view_as(p)
            v.copy_(p.data)
            p.data = v
            self.views.append(v)
            self.grad_views.append(self.flat_grad[off:off + k].view_as(p))
            off += k

    def gather_grads(self):
        """Copy the per-parameter .grad tensors into the flat gradient buffer (one multi-tensor copy)."""
        src, dst = [], []
        for p, gv in zip(self.params, self.grad_views):
            if p.grad is None:
                gv.zero_()
            else:
                src.append(p.grad)
                dst.append(gv)
        if src:
            torch._foreach_copy_(dst, src)
        return self.flat_grad


def all_reduce_gradients(buf, group=None):
    """Sum the flat gradient (and any float64 gradients) over the ranks.  No-op when not initialised / 1 rank, or when
    ``group is False`` (a deliberately local optimiser inside a distributed job, e.g. the single-process reference step of
    tests/test_gpu_dist.py)."""
    if group is False or not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return 1
    dist.all_reduce(buf.flat_grad, op=dist.ReduceOp.SUM, group=group)
    for p in buf.other:
        if p.grad is not None:
            dist.all_reduce(p.grad, op=dist.ReduceOp.SUM, group=group)
    return dist.get_world_size(group)


class DataParallelClipAdam:
    """zero_grad / step pair replacing ``clip_gradient(opt, c); opt.step()`` of the reference train loop, with the
    gradient all-reduce in between.  ``step()`` = gather grads -> all-reduce(sum) -> one fused
    scale(1/world) + clamp(+-c) + Adam kernel over the flat buffers."""

    def __init__(self, module, lr=1e-4, betas=(0.9, 0.999), eps=1e-8, grad_clip=5.0, group=None):
        self._module_params = list(module.parameters())
        self.buf = FlatParamBuffer(module)
        self.lr, self.betas, self.eps, self.grad_clip, self.group = lr, betas, eps, grad_clip, group
        self.exp_avg = torch.zeros_like(self.buf.flat)
        self.exp_avg_sq = torch.zeros_like(self.buf.flat)
        self.step_count = 0
        self._other_opt = torch.optim.Adam(self.buf.other, lr=lr, betas=betas, eps=eps) if self.buf.other else None

    def zero_grad(self):
        for p in self.buf.params + self.buf.other:
            p.grad = None

    # -- torch.optim.Adam-shaped state (checkpoints the reference can resume from, checkpoint.py:51-59) --------------
    def as_torch_adam(self, module=None):
        """A real ``torch.optim.Adam`` with this optimiser's hyper-parameters, step count and moments.  ``module``: a copy
        of the optimised module (same parameter order) whose parameters the new optimiser should own; default: the
        module's own parameters."""
        params = self.buf.params + self.buf.other
        if module is not None:
            mine = {id(p): i for i, p in enumerate(self._module_params)}
            theirs = [p for p in module.parameters()]
            params = [theirs[mine[id(p)]] for p in params]
        opt = torch.optim.Adam(params, lr=self.lr, betas=self.betas, eps=self.eps)
        if self.step_count > 0:
            off = 0
            for p, q in zip(self.buf.params, params):
                k = p.numel()
                opt.state[q] = {"step": torch.tensor(float(self.step_count)),
                                "exp_avg": self.exp_avg[off:off + k].view_as(p).clone(),
                                "exp_avg_sq": self.exp_avg_sq[off:off + k].view_as(p).clone()}
                off += k
            if self._other_opt is not None:
                for p, q in zip(self.buf.other, params[len(self.buf.params):]):
                    st = self._other_opt.state.get(p)
                    if st:
                        opt.state[q] = {k_: (v.clone() if torch.is_tensor(v) else v) for k_, v in st.items()}
        return opt

    def load_torch_adam(self, opt):
        """Adopt step count and moments of a ``torch.optim.Adam`` over the same parameters in the same order (e.g. the
        ``decoder_optimizer`` of a reference checkpoint)."""
        theirs = [p for g in opt.param_groups for p in g["params"]]
        assert len(theirs) == len(self.buf.params) + len(self.buf.other), "optimiser covers a different parameter set"
        off, steps = 0, []
        for p, q in zip(self.buf.params, theirs):
            k = p.numel()
            st = opt.state.get(q)
            if st:
                self.exp_avg[off:off + k].copy_(st["exp_avg"].reshape(-1))
                self.exp_avg_sq[off:off + k].copy_(st["exp_avg_sq"].reshape(-1))
                steps.append(int(st["step"]))
            off += k
        self.step_count = max(steps) if steps else 0
        if opt.param_groups:
            self.lr = opt.param_groups[0]["lr"]
            self.betas = tuple(opt.param_groups[0]["betas"])
            self.eps = opt.param_groups[0]["eps"]

    @torch.no_grad()
    def step(self):
        from . import ops
        self.step_count += 1
        self.buf.gather_grads()
        world = all_reduce_gradients(self.buf, self.group)
        ops.clip_adam_step(self.buf.flat, self.buf.flat_grad, self.exp_avg, self.exp_avg_sq, self.step_count,
                           lr=self.lr, betas=self.betas, eps=self.eps, grad_clip=self.grad_clip,
                           grad_scale=1.0 / world)
        if self._other_opt is not None:
            for p in self.buf.other:
                if p.grad is not None:
                    p.grad.mul_(1.0 / world).clamp_(-self.grad_clip, self.grad_clip)
            self._other_opt.step()
