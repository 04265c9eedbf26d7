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


# Flat-buffer layout for an AttentionDecoder, in the order its backward FINISHES the gradients ("bucket in
# reverse order"): bucket 0 is final before the time loop starts, bucket 1 after the hoisted recurrent weight-gradient
# contractions, bucket 2 after the attention-projection pass.  Inside bucket 1 the three row blocks of the stacked
# [W_dec; W_fbeta; W_hh] gradient (and of its bias) are adjacent, so the kernels write ONE (NZ, D) tile straight into place.
_ATT_BUCKETS = (
    ("fc.weight", "fc.bias"),
    ("h_lin.weight", "h_lin.bias", "c_lin.weight", "c_lin.bias",
     "attention.dec_att.weight", "f_beta.weight", "decode_step.weight_hh",
     "attention.dec_att.bias", "f_beta.bias", "decode_step.bias_hh", "decode_step.bias_ih",
     "decode_step.weight_ih", "embedding.weight"),
    ("attention.full_att.weight", "attention.full_att.bias", "attention.enc_att.bias", "attention.enc_att.weight"),
)


class FlatParamBuffer:
    """Re-points every trainable fp32 parameter of ``module`` at a view of one flat buffer (and keeps a flat
    gradient buffer of the same layout).  Device-agnostic; float64 parameters (GloVe table) are listed separately.

    ``bucketed=True`` (AttentionDecoder only): parameters are laid out in gradient-completion order (``_ATT_BUCKETS``), every
    parameter starts on a 16-byte boundary, and the buffer doubles as the decoder's GRADIENT SINK — the backward kernels write
    the weight gradients straight into ``flat_grad`` (no gather copy) and the buckets can be all-reduced while the rest of
    the backward still runs."""

    def __init__(self, module, bucketed=False):
        named = [(k, p) for k, p in module.named_parameters() if p.requires_grad]
        by_name = dict(named)
        self.other = [p for _, p in named if p.dtype != torch.float32]
        order = [(k, p) for k, p in named if p.dtype == torch.float32]
        self.bucket_ranges = None
        if bucketed:
            flat_names = [k for grp in _ATT_BUCKETS for k in grp]
            assert set(k for k, _ in order) <= set(flat_names), "bucketed layout only knows the AttentionDecoder parameters"
            order = [(k, by_name[k]) for k in flat_names if k in by_name and by_name[k].dtype == torch.float32]
        self.names = [k for k, _ in order]
        self.params = [p for _, p in order]
        assert self.params, "no trainable fp32 parameters"
        dev = self.params[0].device
        offs, off = [], 0
        for p in self.params:
            if bucketed:
                off = (off + 3) // 4 * 4                # 16-byte aligned views: the kernels store 128 bits at a time
            offs.append(off)
            off += p.numel()
        n = (off + 3) // 4 * 4 if bucketed else off
        self.flat = torch.zeros(n, device=dev, dtype=torch.float32)
        self.flat_grad = torch.zeros(n, device=dev, dtype=torch.float32)
        self.views, self.grad_views, self.offsets = [], [], {}
        for k, p, o in zip(self.names, self.params, offs):
            v = self.flat[o:o + p.numel()].view_as(p)
            v.copy_(p.data)
            p.data = v
            self.views.append(v)
            self.grad_views.append(self.flat_grad[o:o + p.numel()].view_as(p))
            self.offsets[k] = (o, p.numel())
        self.in_place = set()            # names whose gradient the last backward wrote straight into flat_grad
        if bucketed:
            self.bucket_ranges = []
            for grp in _ATT_BUCKETS:
                ks = [k for k in grp if k in self.offsets]
                lo = min(self.offsets[k][0] for k in ks)
                hi = max(self.offsets[k][0] + self.offsets[k][1] for k in ks)
                self.bucket_ranges.append((lo, (hi + 3) // 4 * 4))

    # -- gradient sink (bucketed layout) ------------------------------------------------------------------------
    def grad_view(self, name, shape=None):
        o, k = self.offsets[name]
        v = self.flat_grad[o:o + k]
        return v.view(shape) if shape is not None else v

    def group_view(self, names, shape):
        """One view over several ADJACENT parameters' gradients (the stacked [W_dec; W_fbeta; W_hh] tile); None if they are not
        contiguous in this layout."""
        o0 = self.offsets[names[0]][0]
        o = o0
        for k in names:
            if k not in self.offsets or self.offsets[k][0] != o:
                return None
            o += self.offsets[k][1]
        return self.flat_grad[o0:o].view(shape)

    def gather_grads(self):
        """Copy the per-parameter .grad tensors into the flat gradient buffer (one multi-tensor copy); gradients the backward
        already wrote in place (``in_place``) are skipped — they may even be mid-all-reduce."""
        src, dst = [], []
        for k, p, gv in zip(self.names, self.params, self.grad_views):
            if k in self.in_place:
                continue
            if p.grad is None:
                gv.zero_()
            elif p.grad.data_ptr() != gv.data_ptr():
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

    def __init__(self, module, lr=1e-4, betas=(0.9, 0.999), eps=1e-8, grad_clip=5.0, group=None, overlap=None):
        """overlap (default: on for an AttentionDecoder): the decoder backward writes its weight gradients straight into the
        flat buffer, in completion order, and each bucket's all-reduce is issued on NCCL's stream behind an event the backward
        records as soon as the bucket is final — the fc gradients travel during the whole time loop, the recurrent ones during
        the attention-projection pass; only the last 1.3 M elements are reduced after the backward.  Requires ONE backward per
        optimiser step (a second backward before ``step()`` falls back to the copy path by itself)."""
        self._module_params = list(module.parameters())
        is_att = hasattr(module, "decode_step") and hasattr(module, "attention") and hasattr(module, "f_beta")
        if overlap is None:
            overlap = is_att
        self.overlap = bool(overlap and is_att)
        self.buf = FlatParamBuffer(module, bucketed=self.overlap)
        self._works, self._events, self._comm_stream, self._reduced = [], None, None, set()
        if self.overlap:
            module._icd_grad_sink = self
        self.lr, self.betas, self.eps, self.grad_clip, self.group = lr, betas, eps, grad_clip, group
        self.exp_avg = torch.zeros_like(self.buf.flat)
        self.exp_avg_sq = torch.zeros_like(self.buf.flat)
        self.step_count = 0
        self._other_opt = torch.optim.Adam(self.buf.other, lr=lr, betas=betas, eps=eps) if self.buf.other else None

    def zero_grad(self):
        for p in self.buf.params + self.buf.other:
            p.grad = None
        self.buf.in_place = set()

    # -- hooks called by the decoder's autograd Function (models/attention.py) ---------------------------------
    def _world(self):
        if self.group is False or not (dist.is_available() and dist.is_initialized()):
            return 1
        return dist.get_world_size(self.group)

    def sink_ready(self):
        """The backward may write into the flat gradient buffer only if no gradient is pending on any parameter (a second
        backward before step() must ACCUMULATE, which in-place kernel stores would not do)."""
        return self.overlap and not self.buf.in_place and all(p.grad is None for p in self.buf.params)

    def bucket_events(self):
        """Two reusable CUDA events the backward records when bucket 0 / bucket 1 are final -> raw cudaEvent_t handles."""
        if self._events is None:
            self._events = [torch.cuda.Event(), torch.cuda.Event()]
            for e in self._events:
                e.record()                   # forces creation of the underlying cudaEvent_t
        return [int(e.cuda_event) for e in self._events]

    def backward_issued(self, names):
        """Called right after the backward C call returned (everything is only ENQUEUED): start the all-reduce of buckets 0
        and 1 behind their events, on NCCL's own stream, while the tail of the backward still runs on the compute stream."""
        self.buf.in_place = set(names)
        self._reduced = set()
        if self._world() == 1 or set(self.buf.names) - self.buf.in_place:
            return          # (a parameter without an in-place gradient, e.g. the unused table under use_bert: step() reduces all)
        if self._comm_stream is None:
            self._comm_stream = torch.cuda.Stream()
        for i, ev in enumerate(self._events):
            lo, hi = self.buf.bucket_ranges[i]
            self._comm_stream.wait_event(ev)
            with torch.cuda.stream(self._comm_stream):
                self._works.append(dist.all_reduce(self.buf.flat_grad[lo:hi], op=dist.ReduceOp.SUM, group=self.group,
                                                   async_op=True))
            self._reduced.add(i)

    # -- torch.optim.Adam-shaped state (checkpoints the reference can resume from, checkpoint.py:51-59) --------------
    def as_torch_adam(self, module=None):
        """A real ``torch.optim.Adam`` with this optimiser's hyper-parameters, step count and moments.  ``module``: a copy
        of the optimised module (same parameter order) whose parameters the new optimiser should own; default: the
        module's own parameters."""
        mine_req = [p for p in self._module_params if p.requires_grad]          # module order, like the reference's Adam
        to = {id(p): p for p in mine_req}
        if module is not None:
            idx = {id(p): i for i, p in enumerate(self._module_params)}
            theirs = list(module.parameters())
            to = {id(p): theirs[idx[id(p)]] for p in mine_req}
        opt = torch.optim.Adam([to[id(p)] for p in mine_req], lr=self.lr, betas=self.betas, eps=self.eps)
        if self.step_count > 0:
            for name, p in zip(self.buf.names, self.buf.params):
                off, k = self.buf.offsets[name]
                opt.state[to[id(p)]] = {"step": torch.tensor(float(self.step_count)),
                                        "exp_avg": self.exp_avg[off:off + k].view_as(p).clone(),
                                        "exp_avg_sq": self.exp_avg_sq[off:off + k].view_as(p).clone()}
            if self._other_opt is not None:
                for p in self.buf.other:
                    st = self._other_opt.state.get(p)
                    if st:
                        opt.state[to[id(p)]] = {k_: (v.clone() if torch.is_tensor(v) else v) for k_, v in st.items()}
        return opt

    def load_torch_adam(self, opt):
        """Adopt step count and moments of a ``torch.optim.Adam`` over the same parameters in the same order (e.g. the
        ``decoder_optimizer`` of a reference checkpoint)."""
        theirs = [p for g in opt.param_groups for p in g["params"]]
        assert len(theirs) == len(self.buf.params) + len(self.buf.other), "optimiser covers a different parameter set"
        # ``opt`` lists the parameters in module order; this buffer may hold them in bucket order
        mine = [p for p in self._module_params if p.requires_grad]
        their_of = {id(p): q for p, q in zip(mine, theirs)}
        steps = []
        for name, p in zip(self.buf.names, self.buf.params):
            off, k = self.buf.offsets[name]
            st = opt.state.get(their_of[id(p)])
            if st:
                self.exp_avg[off:off + k].copy_(st["exp_avg"].reshape(-1))
                self.exp_avg_sq[off:off + k].copy_(st["exp_avg_sq"].reshape(-1))
                steps.append(int(st["step"]))
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
        if self._reduced:                    # buckets 0 / 1 are already in flight: wait for them, reduce only the rest
            world = self._world()
            for w in self._works:
                w.wait()                     # (stream-side wait: the host does not block)
            self._works = []
            for i, (lo, hi) in enumerate(self.buf.bucket_ranges):
                if i not in self._reduced:
                    dist.all_reduce(self.buf.flat_grad[lo:hi], op=dist.ReduceOp.SUM, group=self.group)
            for p in self.buf.other:
                if p.grad is not None:
                    dist.all_reduce(p.grad, op=dist.ReduceOp.SUM, group=self.group)
            self._reduced = set()
        else:
            world = all_reduce_gradients(self.buf, self.group)
        ops.clip_adam_step(self.buf.flat, self.buf.flat_grad, self.exp_avg, self.exp_avg_sq, self.step_count,
                           lr=self.lr, betas=self.betas, eps=self.eps, grad_clip=self.grad_clip,
                           grad_scale=1.0 / world)
        if self._other_opt is not None:
            for p in self.buf.other:
                if p.grad is not None:
                    p.grad.mul_(1.0 / world).clamp_(-self.grad_clip, self.grad_clip)
            self._other_opt.step()
        self.buf.in_place = set()            # the next backward decides again whether it may write in place
