"""TEST INFRASTRUCTURE ONLY — never imported by the product path.

Imports the UNMODIFIED reference (``/root/reference``) in the build container so that
(a) the oracle restatement in this directory can be pinned against it, and
(b) golden vectors can be generated (``tests/golden/make_golden.py``).

``/root/reference`` does not exist on the GPU box, so nothing that runs there
(``-m gpu`` tests, ``smoke()``, ``bench.py``) may call :func:`load_reference`.

The reference imports five third-party modules that are absent offline and are not touched
by decoder arithmetic: ``pytorch_pretrained_bert`` (models/attention.py:7),
``bcolz`` (embed.py:4), ``nltk`` (dataset.py:10, vocabulary.py:4), ``pycocotools``
(dataset.py:4, vocabulary.py:2) and ``imageio`` (gen_captions.py:12).  They are registered
as empty stub modules; no reference source is copied.
"""
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("ICD_REFERENCE_ROOT", "/root/reference")


def reference_available(root=None):
    return os.path.isfile(os.path.join(root or REFERENCE_ROOT, "models", "attention.py"))


def _stub(name, **attrs):
    if name in sys.modules:
        return sys.modules[name]
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


class _Namespace(types.SimpleNamespace):
    pass


def load_reference(root=None):
    """Return a namespace with the reference's hot-path classes/functions.

    ``root``: where the unmodified reference lives — ``/root/reference`` in the build container (default), or the verbatim
    install under ``baseline/_ref`` (baseline/install_ref.py) that travels to the GPU box for ``bench.py --impl reference``.

    The reference uses top-level module names (``models``, ``vocabulary`` ...) that
    would shadow/be shadowed by other packages, so its modules are imported with
    ``/root/reference`` at the front of ``sys.path`` and then *renamed* in
    ``sys.modules`` under a ``_icd_ref.`` prefix to keep the import state clean.
    """
    root = root or REFERENCE_ROOT
    if not reference_available(root):
        raise RuntimeError("reference tree not present at %s" % root)
    if "_icd_ref" in sys.modules:
        return sys.modules["_icd_ref"].ns

    _stub("pytorch_pretrained_bert", BertTokenizer=object, BertModel=object)
    _stub("bcolz")
    _stub("nltk")
    _stub("imageio")
    pc = _stub("pycocotools")
    pcc = _stub("pycocotools.coco", COCO=object)
    pc.coco = pcc

    shadowed = ["models", "models.attention", "models.baseline", "models.encoder", "vocabulary",
                "embed", "dataset", "metric", "train_utils", "checkpoint", "gen_captions",
                "pathconf", "eval_func", "eval_func.bleu", "eval_func.bleu.bleu"]
    saved = {k: sys.modules.pop(k) for k in list(sys.modules) if k in shadowed or k.startswith("eval_func")}
    sys.path.insert(0, root)
    cwd = os.getcwd()
    try:
        import models.attention as ref_att      # noqa
        import models.baseline as ref_base      # noqa
        import gen_captions as ref_gen          # noqa
        import vocabulary as ref_vocab          # noqa
        import train_utils as ref_tu            # noqa
    finally:
        os.chdir(cwd)
        sys.path.remove(root)
    ns = _Namespace(
        attention=ref_att, baseline=ref_base, gen_captions=ref_gen, vocabulary=ref_vocab,
        train_utils=ref_tu,
        SoftAttention=ref_att.SoftAttention, AttentionDecoder=ref_att.AttentionDecoder,
        AttentionDecoderParams=ref_att.AttentionDecoderParams,
        BaselineDecoder=ref_base.BaselineDecoder, BaselineDecoderParams=ref_base.BaselineDecoderParams,
        beam_search=ref_gen.attention_caption_image_beam_search,
        Vocabulary=ref_vocab.Vocabulary, clip_gradient=ref_tu.clip_gradient,
    )
    # move the reference's generically-named modules out of the way
    for k in list(sys.modules):
        if k in shadowed or k.startswith("eval_func"):
            mod = sys.modules.pop(k)
            sys.modules["_icd_ref." + k] = mod
    sys.modules.update(saved)
    holder = types.ModuleType("_icd_ref")
    holder.ns = ns
    sys.modules["_icd_ref"] = holder
    return ns


def make_reference_vocab(ns, vocab_size):
    """Synthetic ``Vocabulary`` with the reference's id layout (vocabulary.py:52-58)."""
    v = ns.Vocabulary()
    v.add_word("<pad>")
    for i in range(vocab_size - 4):
        v.add_word("w%d" % i)
    v.add_word("<start>")
    v.add_word("<end>")
    v.add_word("<unk>")
    assert len(v) == vocab_size
    return v
