"""Build ``libmuvo_b200.so`` in-tree with nvcc for sm_100a.

``python -m muvo_b200.build`` (or ``__graft_entry__.build()``).  nvcc cross-compiles
without a GPU; the resulting ``.so`` lives next to this file so that it travels with
the source tree.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
INCLUDE = os.path.join(os.path.dirname(HERE), "include")
LIB_PATH = os.path.join(HERE, "libmuvo_b200.so")
OBJ_DIR = os.path.join(HERE, "csrc", "_obj")

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden"]
# points.cu reproduces numpy's float64 arithmetic: no FMA contraction allowed there.
PER_FILE = {"points.cu": ["-fmad=false"] + os.environ.get("MUVO_NVCC_EXTRA", "").split(),
            "points_mega.cu": ["-fmad=false"] + os.environ.get("MUVO_NVCC_EXTRA", "").split(),
            "merge.cu": ["-fmad=false"]}   # MUVO_NVCC_EXTRA: tuning builds only
SOURCES = ["api.cu", "points.cu", "points_mega.cu", "ssc.cu", "bev.cu", "bev_stream.cu", "merge.cu", "pyramid.cu", "scal.cu", "pillar.cu"]


def _nvcc() -> str:
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found; cannot build libmuvo_b200.so")
    return nvcc


def _stale(target: str, deps) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    nvcc = _nvcc()
    os.makedirs(OBJ_DIR, exist_ok=True)
    headers = [os.path.join(CSRC, "common.cuh"), os.path.join(CSRC, "points_dev.cuh"), os.path.join(CSRC, "tma.cuh"), os.path.join(CSRC, "bev_stream.cuh"), os.path.join(INCLUDE, "muvo_b200.h")]
    objs = []
    for src in SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(OBJ_DIR, src.replace(".cu", ".o"))
        objs.append(o)
        if force or _stale(o, [s] + headers):
            cmd = [nvcc, *ARCH, *COMMON, *PER_FILE.get(src, []), *os.environ.get("MUVO_NVCC_ALL", "").split(), "-I", INCLUDE, "-c", s, "-o", o]
            if verbose:
                cmd[1:1] = ["-Xptxas", "-v"]
                print(" ".join(cmd), flush=True)
            subprocess.run(cmd, check=True)
    if force or _stale(LIB_PATH, objs):
        cmd = [nvcc, *ARCH, "-shared", "-o", LIB_PATH, *objs, "-cudart", "static"]
        if verbose:
            print(" ".join(cmd), flush=True)
        subprocess.run(cmd, check=True)
    return LIB_PATH


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(path)
