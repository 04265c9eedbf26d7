"""Vocabulary type the decoders are constructed with.

Mirrors the reference's support type (vocabulary.py:8-35): ``w2i``/``i2w`` dicts, ``add_word``,
``__call__`` with ``<unk>`` fallback, ``__len__``.  Id layout produced by the reference's
``build_vocab`` (vocabulary.py:52-58): ``<pad>`` = 0, words, then ``<start>``, ``<end>``,
``<unk>`` as the last three ids.  Dataset-driven vocabulary *building* (COCO + nltk) is out of
scope; :func:`synthetic_vocab` builds an instance of any size.
"""

PAD_TOKEN = '<pad>'
START_TOKEN = '<start>'
END_TOKEN = '<end>'
UNK_TOKEN = '<unk>'


class _VocabularyMeta(type):
    """``isinstance(v, Vocabulary)`` also accepts the reference's own ``vocabulary.Vocabulary`` (a distinct class
    object with the same surface), so a checkpointed / pickled reference vocabulary passes the constructor assert
    of the drop-in decoders (models/attention.py:84)."""

    def __instancecheck__(cls, obj):
        if type.__instancecheck__(cls, obj):
            return True
        return (type(obj).__name__ == "Vocabulary" and hasattr(obj, "w2i") and hasattr(obj, "i2w")
                and callable(obj) and hasattr(obj, "__len__"))


class Vocabulary(object, metaclass=_VocabularyMeta):
    def __init__(self):
        self.w2i = {}
        self.i2w = {}
        self.idx = 0

    def add_word(self, word):
        if word not in self.w2i:
            self.w2i[word] = self.idx
            self.i2w[self.idx] = word
            self.idx += 1

    def __call__(self, word):
        if word not in self.w2i:
            return self.w2i[UNK_TOKEN]
        return self.w2i[word]

    def __len__(self):
        return len(self.w2i)


def synthetic_vocab(vocab_size):
    """``<pad>``, ``vocab_size - 4`` dummy words, ``<start>``, ``<end>``, ``<unk>``."""
    assert vocab_size >= 5
    v = Vocabulary()
    v.add_word(PAD_TOKEN)
    for i in range(vocab_size - 4):
        v.add_word('w%d' % i)
    v.add_word(START_TOKEN)
    v.add_word(END_TOKEN)
    v.add_word(UNK_TOKEN)
    return v
