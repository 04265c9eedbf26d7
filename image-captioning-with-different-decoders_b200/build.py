"""In-tree build of libicd_b200.so (the C-ABI CUDA library, include/icd_b200.h) for sm_100a.

    python -m icd_b200.build            # or: __graft_entry__.build()

nvcc cross-compiles without a GPU.  The library is written next to this file so it travels with the
repo snapshot to the GPU box (it is git-ignored, not gpurun-ignored).
"""
import hashlib
import json
import os
import subprocess
import sys
import time
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libicd_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
FLAGS = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden"] + os.environ.get("NVCC_EXTRA", "").split()


def _sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def _digest(paths):
    h = hashlib.sha256()
    for p in paths:
        with open(p, "rb") as f:
            h.update(f.read())
    h.update(" ".join(ARCH + FLAGS).encode())
    return h.hexdigest()


def build(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    srcs = _sources()
    hdrs = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith((".cuh", ".h"))]
    hdrs.append(os.path.join(os.path.dirname(HERE), "include", "icd_b200.h"))
    stamp_path = os.path.join(OBJ, "stamp.txt")
    stamp = _digest([os.path.join(CSRC, s) for s in srcs] + hdrs)
    if not force and os.path.exists(LIB) and os.path.exists(stamp_path) and open(stamp_path).read() == stamp:
        _record(stamp, rebuilt=False, seconds=0.0)
        return LIB
    t0 = time.time()

    def compile_one(src):
        obj = os.path.join(OBJ, src[:-3] + ".o")
        cmd = [NVCC] + ARCH + FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", os.path.join(CSRC, src), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, r.stdout, r.stderr))
        if verbose:
            sys.stderr.write(r.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        objs = list(ex.map(compile_one, srcs))
    cmd = [NVCC] + ARCH + ["-shared", "-o", LIB] + objs + ["-lcudart_static", "-ldl", "-lrt", "-lpthread"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
    with open(stamp_path, "w") as f:
        f.write(stamp)
    _record(stamp, rebuilt=True, seconds=time.time() - t0)
    return LIB


def _record(stamp, rebuilt, seconds):
    """build/last_build.json: what the last build() call did — compiled every source (and with which nvcc) or found the
    library up to date with the digest of sources + headers + flags.  Evidence only; nothing reads it."""
    try:
        with open(LIB, "rb") as f:
            lib_sha = hashlib.sha256(f.read()).hexdigest()
        ver = subprocess.run([NVCC, "--version"], capture_output=True, text=True).stdout.strip().splitlines()[-1] if rebuilt else None
        with open(os.path.join(OBJ, "last_build.json"), "w") as f:
            json.dump({"rebuilt_from_source": rebuilt, "seconds": round(seconds, 1), "sources": _sources(), "arch": ARCH, "flags": FLAGS,
                       "nvcc": ver, "sources_digest": stamp, "lib_sha256": lib_sha, "unix_time": int(time.time())}, f, indent=1)
    except OSError:
        pass


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
