"""Builds libdartb.so in-tree with nvcc for sm_100a (the .so travels to the GPU box).

Each (topology, precision) instantiation of the kernels is its own translation unit (inst.cu
compiled with -DINST_*), built in parallel, then linked with the host API (dartb.cu)."""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
# developer A/B builds: DARTB_SO_SUFFIX=_x DARTB_NVCC_FLAGS="-DFOO" give libdartb_x.so next to the product library
_SFX = os.environ.get("DARTB_SO_SUFFIX", "")
OBJ = os.path.join(HERE, "build" + _SFX)
SO = os.path.join(HERE, "libdartb%s.so" % _SFX)
DEPS = ["dartb.cu", "inst.cu", "kernels.cuh", "planar_kernels.cuh", "planar_loop.cuh", "planar_coop.cuh", "task_kinds.cuh", "warp_group.cuh", "planar_model.h", "lower.h",
        os.path.join("..", "..", "include", "dartb.h")]
INSTANCES = [("hopper", "TopoHopper"), ("walker", "TopoWalker"), ("cheetah", "TopoCheetah"), ("snake", "TopoSnake"),
             ("loop", None)]
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr"] + os.environ.get("DARTB_NVCC_FLAGS", "").split()


def nvcc_path() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


STAMP = SO + ".srchash"   # sha256 of the sources the library was built from (travels with the .so)


def source_hash() -> str:
    h = hashlib.sha256()
    for d in sorted(DEPS):
        with open(os.path.join(CSRC, d), "rb") as f:
            h.update(d.encode() + b"\0" + f.read() + b"\0")
    h.update(" ".join(ARCH + COMMON).encode())
    return h.hexdigest()


def is_stale() -> bool:
    """The library is stale when its sources changed.  Decided by CONTENT when the stamp written at build time is there:
    a copy of the tree (the snapshot a GPU box receives) need not preserve modification times, and a spurious rebuild there
    costs minutes inside the first test that loads the library.  Without a stamp: by modification time."""
    if not os.path.exists(SO):
        return True
    if os.path.exists(STAMP):
        with open(STAMP) as f:
            return f.read().strip() != source_hash()
    t = os.path.getmtime(SO)
    return any(os.path.getmtime(os.path.join(CSRC, d)) > t for d in DEPS)


def _run(cmd, verbose):
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(" ".join(cmd) + "\n" + res.stdout + res.stderr)
        raise RuntimeError("nvcc failed")
    return res.stdout + res.stderr


def build(force: bool = False, verbose: bool = False, jobs: int = 0) -> str:
    if not force and not is_stale():
        return SO
    nvcc = nvcc_path()
    os.makedirs(OBJ, exist_ok=True)
    ptxas = ["-Xptxas", "-v"] if verbose else []
    jobs_list = [(os.path.join(OBJ, "dartb.o"), [nvcc] + ARCH + COMMON + ptxas + ["-c", os.path.join(CSRC, "dartb.cu")])]
    for name, topo in INSTANCES:
        for real, sfx in (("float", "f"), ("double", "d")):
            defs = ["-DINST_REAL=%s" % real, "-DINST_SUFFIX=%s_%s" % (name, sfx)]
            defs += ["-DINST_LOOP=1"] if topo is None else ["-DINST_TOPO=%s" % topo]
            obj = os.path.join(OBJ, "inst_%s_%s.o" % (name, sfx))
            jobs_list.append((obj, [nvcc] + ARCH + COMMON + ptxas + defs + ["-c", os.path.join(CSRC, "inst.cu")]))
    only = os.environ.get("DARTB_BUILD_ONLY")   # developer shortcut: "hopper_f,dartb" recompiles just those objects
    if only:
        keep = set(only.split(","))
        jobs_all = jobs_list
        jobs_list = [j for j in jobs_list if any(os.path.basename(j[0]) in ("inst_%s.o" % k, "%s.o" % k) for k in keep)]
        missing = [j[0] for j in jobs_all if j not in jobs_list and not os.path.exists(j[0])]
        if missing:
            raise RuntimeError("DARTB_BUILD_ONLY needs the other objects to exist: %s" % missing)
    else:
        jobs_all = jobs_list
    nj = jobs or min(len(jobs_list), os.cpu_count() or 4)
    with ThreadPoolExecutor(nj) as ex:
        logs = list(ex.map(lambda j: _run(j[1] + ["-o", j[0]], verbose), jobs_list))
    log = _run([nvcc] + ARCH + ["-shared", "-Xcompiler", "-fPIC", "-o", SO] + [j[0] for j in jobs_all], verbose)
    if verbose:
        sys.stderr.write("".join(logs) + log)
    if not only:   # (a partial developer rebuild leaves the other objects as they were: keep the previous verdict)
        with open(STAMP, "w") as f:
            f.write(source_hash() + "\n")
    elif os.path.exists(STAMP):
        os.remove(STAMP)
    return SO


if __name__ == "__main__":
    print(build(force=True, verbose="-v" in sys.argv))
