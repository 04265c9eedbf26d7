"""Optimiser glue of the reference train loops, B200-native.

``clip_gradient`` keeps the reference signature/semantics (train_utils.py:2-12: element-wise clamp of every
gradient to +-grad_clip).  The fused equivalent of

    clip_gradient(optimizer, grad_clip); optimizer.step()      # models/attention.py:423-428

is ``icd_b200.parallel.DataParallelClipAdam`` (one kernel over a flat parameter buffer, after the gradient all-reduce).
"""


def clip_gradient(optimizer, grad_clip):
    """Reference semantics (train_utils.py:2-12)."""
    for group in optimizer.param_groups:
        for param in group['params']:
            if param.grad is not None:
                param.grad.data.clamp_(-grad_clip, grad_clip)
