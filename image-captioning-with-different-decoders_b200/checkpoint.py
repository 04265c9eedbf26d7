"""Checkpoint helpers with the reference's interface (checkpoint.py:7-62), compatible with the reference in BOTH directions.

The reference pickles whole modules (``torch.save({'encoder': encoder, 'decoder': decoder, ...})``, checkpoint.py:51-59),
so its ``*.pth.tar`` files name the classes ``models.attention.AttentionDecoder`` / ``SoftAttention``,
``models.baseline.BaselineDecoder`` and ``vocabulary.Vocabulary`` by module path.  ``load_checkpoint`` maps those paths onto
the drop-in classes of this package while unpickling (everything else — e.g. the reference's encoder class — resolves
normally), so a decoder trained with the reference comes back as a CUDA-kernel-backed module with the same weights.
``save_checkpoint`` writes files the REFERENCE can load: while pickling, the drop-in classes are announced under the
reference's module paths (``models.attention.AttentionDecoder`` ...), parameters that were re-pointed into a flat buffer are
saved as ordinary tensors, and the fused optimiser is saved as a real ``torch.optim.Adam`` holding the same step count and
moments (the reference resumes with ``decoder_optimizer`` as it was pickled, models/attention.py:359-364).
torch >= 2.6 defaults ``torch.load`` to ``weights_only=True``; whole-module pickles (the reference's format) need
``weights_only=False``, which executes pickled code — only load checkpoints you trust.
"""
import contextlib
import copy
import sys
import types
import os
import pickle

import torch

CHECKPOINTS_DIR = 'checkpoints'

_ALIASES = {
    ("models.attention", "SoftAttention"): ("icd_b200.models.attention", "SoftAttention"),
    ("models.attention", "AttentionDecoderParams"): ("icd_b200.models.attention", "AttentionDecoderParams"),
    ("models.attention", "AttentionDecoder"): ("icd_b200.models.attention", "AttentionDecoder"),
    ("models.baseline", "BaselineDecoderParams"): ("icd_b200.models.baseline", "BaselineDecoderParams"),
    ("models.baseline", "BaselineDecoder"): ("icd_b200.models.baseline", "BaselineDecoder"),
    ("vocabulary", "Vocabulary"): ("icd_b200.vocabulary", "Vocabulary"),
}


class _AliasingUnpickler(pickle.Unpickler):
    def find_class(self, module, name):
        module, name = _ALIASES.get((module, name), (module, name))
        return super().find_class(module, name)


class _AliasingPickle:
    """Stand-in for the ``pickle`` module handed to ``torch.load(pickle_module=...)``."""
    __name__ = "icd_b200_aliasing_pickle"
    Unpickler = _AliasingUnpickler
    load = staticmethod(lambda f, **kw: _AliasingUnpickler(f, **kw).load())
    dump = staticmethod(pickle.dump)
    Pickler = pickle.Pickler


def load_checkpoint_file(path, device="cpu"):
    """Read a checkpoint written by this package OR by the reference's ``save_checkpoint``."""
    return torch.load(path, map_location=str(device), pickle_module=_AliasingPickle, weights_only=False)


def load_checkpoint(device, args, verbose=True):
    """Reference signature (checkpoint.py:7-18): loads ``checkpoints/{args.checkpoint}``."""
    path = os.path.join(CHECKPOINTS_DIR, f'{args.checkpoint}')
    if verbose:
        print(f'Loading checkpoint {path}')
    return load_checkpoint_file(path, device)


def unpack_checkpoint(chkpt):
    """(checkpoint.py:21-36)"""
    return (chkpt['epoch'], chkpt['encoder'], chkpt['decoder'], chkpt['encoder_optimizer'], chkpt['decoder_optimizer'],
            chkpt['metrics'])


@contextlib.contextmanager
def _announce_as_reference_classes():
    """While active, the drop-in classes carry the reference's module paths (``cls.__module__``) and those paths resolve
    to them in ``sys.modules`` — pickle records ``models.attention.AttentionDecoder`` etc., which is what the reference's
    own ``torch.load`` looks up.  Everything is restored afterwards (also modules of the real reference, if imported)."""
    import importlib
    todo = {}
    for (ref_mod, name), (my_mod, my_name) in _ALIASES.items():
        todo.setdefault(ref_mod, []).append((name, getattr(importlib.import_module(my_mod), my_name)))
    saved_modules = {k: sys.modules.get(k) for k in list(todo) + ["models"]}
    saved_cls = []
    try:
        pkg = types.ModuleType("models")
        pkg.__path__ = []
        sys.modules["models"] = pkg
        for ref_mod, items in todo.items():
            m = types.ModuleType(ref_mod)
            for name, cls in items:
                setattr(m, name, cls)
                saved_cls.append((cls, cls.__module__, cls.__qualname__))
                cls.__module__, cls.__qualname__ = ref_mod, name
            sys.modules[ref_mod] = m
            if ref_mod.startswith("models."):
                setattr(pkg, ref_mod.split(".", 1)[1], m)
        yield
    finally:
        for cls, mod, qn in saved_cls:
            cls.__module__, cls.__qualname__ = mod, qn
        for k, v in saved_modules.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v


def _portable_module(module):
    """A copy whose parameters own their storage (a FlatParamBuffer re-points them at views of one flat buffer) and which
    carries none of the B200-only attributes a reference class would not know."""
    if module is None or not isinstance(module, torch.nn.Module):
        return module
    m = copy.deepcopy(module)
    for p_ in m.parameters():
        p_.data = p_.data.clone()
    for sub in m.modules():
        sub.__dict__.pop("_dropout_mask_override", None)
    return m


def _portable_optimizer(opt, module):
    """DataParallelClipAdam -> torch.optim.Adam over ``module``'s parameters with the same step / exp_avg / exp_avg_sq."""
    if hasattr(opt, "as_torch_adam"):
        return opt.as_torch_adam(module)
    return opt


def save_checkpoint(args, epoch, encoder, decoder, encoder_optimizer, decoder_optimizer, metrics, verbose=True):
    """(checkpoint.py:39-62) — same dictionary, same file name; loadable by the reference's ``load_checkpoint``."""
    dec = _portable_module(decoder)
    state = {
        'epoch': epoch, 'metrics': metrics, 'encoder': encoder, 'decoder': dec,
        'encoder_optimizer': encoder_optimizer, 'decoder_optimizer': _portable_optimizer(decoder_optimizer, dec),
    }
    os.makedirs(CHECKPOINTS_DIR, exist_ok=True)
    path = os.path.join(CHECKPOINTS_DIR, f'{args.model_name}_{epoch}.pth.tar')
    with _announce_as_reference_classes():
        torch.save(state, path)
    if verbose:
        print(f'Saved checkpoint to {path}')
