"""Import alias for the package directory ``image-captioning-with-different-decoders_b200/``.

The mandated directory name contains hyphens and so cannot be imported with a plain
``import`` statement; this shim makes ``import icd_b200`` (and ``icd_b200.models.attention``
etc.) resolve to that directory.  No code lives here.
"""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                      "image-captioning-with-different-decoders_b200")
__path__ = [_real]
__file__ = _os.path.join(_real, "__init__.py")
with open(__file__) as _f:
    exec(compile(_f.read(), __file__, "exec"))
del _os, _f
