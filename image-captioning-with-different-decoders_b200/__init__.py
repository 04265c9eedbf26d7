"""B200-native captioning-decoder hot path (drop-in for models/attention.py, models/baseline.py and
gen_captions.py of SarahAlkhateeb/Image-Captioning-with-Different-Decoders).

Import as ``icd_b200`` (the directory name contains hyphens; ``icd_b200/__init__.py`` aliases it).
The CUDA kernels live in ``csrc/`` and are reached only through the C ABI of ``libicd_b200.so``
(``include/icd_b200.h``); there is no CPU or eager fallback.
"""
__version__ = "0.1.0"
