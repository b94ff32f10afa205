"""Builds recbole_b200/librecbole_b200.so in-tree with nvcc for sm_100a.

    python recbole_b200/build.py [--force] [--verbose]

One translation unit per .cu under csrc/, compiled in parallel, linked into a single shared
library with a C ABI (include/recbole_b200.h).  nvcc cross-compiles without a GPU.
"""
import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "_obj")
LIB = os.path.join(HERE, "librecbole_b200.so")

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = [
    "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
    "-Xcompiler", "-fPIC", "-Xptxas", "-v", "--expt-relaxed-constexpr",
]


def _sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def _stamp(src):
    h = hashlib.sha1()
    for f in [src] + sorted(os.path.join(CSRC, x) for x in os.listdir(CSRC) if x.endswith((".cuh", ".h"))) + [
            os.path.join(os.path.dirname(HERE), "include", "recbole_b200.h")]:
        with open(f, "rb") as fh:
            h.update(fh.read())
    h.update(" ".join(FLAGS).encode())
    return h.hexdigest()


def _compile(name, force, verbose):
    src = os.path.join(CSRC, name)
    obj = os.path.join(OBJ, name[:-3] + ".o")
    stamp_file = obj + ".stamp"
    stamp = _stamp(src)
    if not force and os.path.exists(obj) and os.path.exists(stamp_file) and open(stamp_file).read() == stamp:
        return name, False, ""
    cmd = [NVCC] + FLAGS + ["-c", "-o", obj, src]
    p = subprocess.run(cmd, capture_output=True, text=True)
    log = p.stdout + p.stderr
    with open(obj + ".log", "w") as fh:
        fh.write(log)
    if p.returncode != 0:
        raise RuntimeError("nvcc failed for %s:\n%s" % (name, log[-6000:]))
    with open(stamp_file, "w") as fh:
        fh.write(stamp)
    return name, True, log


def build(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    srcs = _sources()
    with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        results = list(ex.map(lambda n: _compile(n, force, verbose), srcs))
    rebuilt = [n for n, r, _ in results if r]
    if rebuilt or not os.path.exists(LIB):
        objs = [os.path.join(OBJ, n[:-3] + ".o") for n in srcs]
        cmd = [NVCC, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB] + objs + ["-ldl"]
        p = subprocess.run(cmd, capture_output=True, text=True)
        if p.returncode != 0:
            raise RuntimeError("link failed:\n" + p.stdout + p.stderr)
    if verbose:
        print("built %s (recompiled: %s)" % (LIB, ", ".join(rebuilt) or "nothing"))
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose=True)
