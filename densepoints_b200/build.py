"""Builds the CUDA library in-tree: densepoints_b200/_build/libdensepoints_cuda.so
(nvcc, sm_100a only, -lineinfo so ncu's source page maps to the kernels)."""
from __future__ import annotations

import glob
import os
import shutil
import subprocess

_PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_PKG, "csrc")
OUT_DIR = os.path.join(_PKG, "_build")
LIB = os.path.join(OUT_DIR, "libdensepoints_cuda.so")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-shared", "-Xcompiler", "-fPIC"]


def _nvcc():
    for c in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if c and os.path.exists(c):
            return c
    raise RuntimeError("nvcc not found")


def sources():
    root = os.path.dirname(_PKG)
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")) + glob.glob(os.path.join(CSRC, "*.cuh")) +
                  glob.glob(os.path.join(CSRC, "*.h")) +
                  [os.path.join(root, "include", "densepoints_cuda.h")])


def source_hash() -> str:
    """sha256 over the CUDA sources and the C ABI header: ties a committed ncu extract
    (profiles/*.json, written by tools/ncu_extract.py) to the code it profiled."""
    import hashlib
    h = hashlib.sha256()
    for f in sources():
        h.update(os.path.basename(f).encode())
        h.update(open(f, "rb").read())
    return h.hexdigest()


def build_cuda(force: bool = False, verbose: bool = False) -> str:
    srcs = sources()
    stale = (not os.path.exists(LIB)) or any(os.path.getmtime(s) > os.path.getmtime(LIB)
                                              for s in srcs)
    if not (force or stale):
        return LIB
    os.makedirs(OUT_DIR, exist_ok=True)
    cmd = [_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + \
        ["-o", LIB, os.path.join(CSRC, "densepoints_cuda.cu")]
    env = dict(os.environ)
    env.pop("CC", None)
    r = subprocess.run(cmd, cwd=CSRC, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                       text=True)
    if verbose:
        print(r.stdout)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + r.stdout)
    return LIB


if __name__ == "__main__":
    import sys
    print(build_cuda(force="--force" in sys.argv, verbose="-v" in sys.argv))
