"""Builds libdartb.so in-tree with nvcc for sm_100a (the .so travels to the GPU box)."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
SO = os.path.join(HERE, "libdartb.so")
SOURCES = ["dartb.cu"]
DEPS = ["dartb.cu", "planar_kernels.cuh", "planar_model.h", "lower.h", os.path.join("..", "..", "include", "dartb.h")]


def nvcc_path() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def is_stale() -> bool:
    if not os.path.exists(SO):
        return True
    t = os.path.getmtime(SO)
    return any(os.path.getmtime(os.path.join(CSRC, d)) > t for d in DEPS)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not is_stale():
        return SO
    cmd = [nvcc_path(), "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo",
           "-Xcompiler", "-fPIC", "-shared", "--expt-relaxed-constexpr", "-Xptxas", "-v" if verbose else "-O3",
           "-o", SO] + [os.path.join(CSRC, s) for s in SOURCES]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("nvcc failed")
    if verbose:
        sys.stderr.write(res.stdout + res.stderr)
    return SO


if __name__ == "__main__":
    print(build(force=True, verbose="-v" in sys.argv))
