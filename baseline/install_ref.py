"""Install the UNMODIFIED reference for the benchmark's reference arm (`bench.py --impl reference`).

The reference (SarahAlkhateeb/Image-Captioning-with-Different-Decoders) is a flat directory of Python scripts with no
setup.py / pyproject.toml of its own (its only build file belongs to the vendored cocoapi), so
`pip install --target baseline/_ref /root/reference` has nothing to install.  The equivalent for a script tree is a verbatim
copy of its Python import closure — top-level *.py, models/*.py, eval_func/**/*.py; no data, notebooks, logs or the
vendored cocoapi — into baseline/_ref/, which is git-ignored (never part of this repository's history) but NOT
gpurun-ignored, so it travels to the GPU box where /root/reference does not exist.  Nothing is edited: a SHA-256 manifest
of the copied files is written next to them and checked by tests/test_abi_and_host.py when the source tree is present.

    python baseline/install_ref.py            # run in the build container (also called by __graft_entry__.build())
"""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")
SRC = os.environ.get("ICD_REFERENCE_ROOT", "/root/reference")


def _closure(src):
    out = [f for f in sorted(os.listdir(src)) if f.endswith(".py")]
    for sub in ("models", "eval_func"):
        for d, _, files in os.walk(os.path.join(src, sub)):
            out += [os.path.relpath(os.path.join(d, f), src) for f in sorted(files) if f.endswith(".py")]
    return out


def install(src=SRC, dest=DEST, verbose=False):
    """-> manifest dict, or None when the reference tree is not present (GPU box: the prebuilt copy is used)."""
    if not os.path.isfile(os.path.join(src, "models", "attention.py")):
        return None
    manifest = {}
    for rel in _closure(src):
        d = os.path.join(dest, rel)
        os.makedirs(os.path.dirname(d), exist_ok=True)
        shutil.copyfile(os.path.join(src, rel), d)
        manifest[rel] = hashlib.sha256(open(d, "rb").read()).hexdigest()
    with open(os.path.join(dest, "MANIFEST.json"), "w") as f:
        json.dump({"source": src, "files": manifest}, f, indent=1, sort_keys=True)
    if verbose:
        print("installed %d reference files into %s" % (len(manifest), dest))
    return manifest


def available(dest=DEST):
    return os.path.isfile(os.path.join(dest, "models", "attention.py"))


if __name__ == "__main__":
    m = install(verbose=True)
    sys.exit(0 if m else 1)
