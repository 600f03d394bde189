"""Build librlctr_sm100a.so in-tree with nvcc (sm_100a only; cross-compiles without a GPU).

    python -m rl_ctr_prediction_b200.build [--force] [--verbose]

Each ``csrc/*.cu`` is compiled to an object in ``csrc/build/`` (in parallel, skipped when the
object is newer than every source/header) and linked into
``rl_ctr_prediction_b200/librlctr_sm100a.so``.  The library travels to the GPU box with the
repo snapshot; nothing is JIT-compiled at import time.
"""
from __future__ import annotations

import concurrent.futures as cf
import glob
import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
OBJ = os.path.join(CSRC, "build")
LIB = os.path.join(PKG, "librlctr_sm100a.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "-Xptxas", "-v",
    "--expt-relaxed-constexpr",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: librlctr_sm100a.so cannot be built (no fallback path exists)")


def _newest(paths) -> float:
    return max(os.path.getmtime(p) for p in paths)


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def headers():
    return sorted(glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(ROOT, "include", "*.h")))


def _compile(nvcc, src, obj, verbose):
    cmd = [nvcc, *NVCC_FLAGS, "-I", os.path.join(ROOT, "include"), "-c", src, "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed on {src}:\n{r.stdout}\n{r.stderr}")
    with open(obj + ".ptxas.txt", "w") as fh:            # registers / spills / smem per kernel
        fh.write(r.stderr)
    if verbose:
        sys.stderr.write(r.stderr)
    return obj


def build(force: bool = False, verbose: bool = False) -> str:
    nvcc = _nvcc()
    os.makedirs(OBJ, exist_ok=True)
    srcs, hdrs = sources(), headers()
    dep_time = _newest(hdrs + [os.path.abspath(__file__)])
    jobs, objs = [], []
    for src in srcs:
        obj = os.path.join(OBJ, os.path.basename(src)[:-3] + ".o")
        objs.append(obj)
        if force or not os.path.exists(obj) or os.path.getmtime(obj) < max(dep_time, os.path.getmtime(src)):
            jobs.append((src, obj))
    if jobs:
        with cf.ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            list(ex.map(lambda j: _compile(nvcc, j[0], j[1], verbose), jobs))
    if jobs or not os.path.exists(LIB) or os.path.getmtime(LIB) < _newest(objs):
        cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB, *objs]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
