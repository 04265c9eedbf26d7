"""Checkpoint helpers with the reference's interface (checkpoint.py:7-62) that also read the reference's OWN files.

The reference pickles whole modules (``torch.save({'encoder': encoder, 'decoder': decoder, ...})``, checkpoint.py:51-59),
so its ``*.pth.tar`` files name the classes ``models.attention.AttentionDecoder`` / ``SoftAttention``,
``models.baseline.BaselineDecoder`` and ``vocabulary.Vocabulary`` by module path.  ``load_checkpoint`` maps those paths onto
the drop-in classes of this package while unpickling (everything else — e.g. the reference's encoder class — resolves
normally), so a decoder trained with the reference comes back as a CUDA-kernel-backed module with the same weights.
torch >= 2.6 defaults ``torch.load`` to ``weights_only=True``; whole-module pickles need ``weights_only=False`` — only load
checkpoints you trust.
"""
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


def save_checkpoint(args, epoch, encoder, decoder, encoder_optimizer, decoder_optimizer, metrics, verbose=True):
    """(checkpoint.py:39-62)"""
    state = {
        'epoch': epoch, 'metrics': metrics, 'encoder': encoder, 'decoder': decoder,
        'encoder_optimizer': encoder_optimizer, 'decoder_optimizer': decoder_optimizer,
    }
    os.makedirs(CHECKPOINTS_DIR, exist_ok=True)
    path = os.path.join(CHECKPOINTS_DIR, f'{args.model_name}_{epoch}.pth.tar')
    torch.save(state, path)
    if verbose:
        print(f'Saved checkpoint to {path}')
