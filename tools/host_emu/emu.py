"""ctypes wrapper for tools/host_emu/libhost_emu.so: the product's device code compiled for the
CPU (debug + no-GPU regression tests).  Not a product path."""
import ctypes as C
import os
import subprocess

import numpy as np

from dart_env_b200.cstructs import CModel, CTask, pack_model, pack_task

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def lib():
    global _LIB
    if _LIB is None:
        subprocess.check_call(["make", "-C", _HERE, "libhost_emu.so"], stdout=subprocess.DEVNULL)
        L = C.CDLL(os.path.join(_HERE, "libhost_emu.so"))
        L.emu_last_error.restype = C.c_char_p
        L.emu_reset_uniform.restype = C.c_float
        L.emu_reset_uniform.argtypes = [C.c_uint64, C.c_int64, C.c_uint32, C.c_int]
        _LIB = L
    return _LIB


def substep(model, task, q, dq, tau=None, fext=None, f64=False, lcp_mode=0, pgs_iters=30, maxc=8, variant=0, hints=None):
    L = lib()
    cm, ct = pack_model(model), pack_task(task)
    q = np.ascontiguousarray(q, dtype=np.float64)
    dq = np.ascontiguousarray(dq, dtype=np.float64)
    n, nd = q.shape
    dp = lambda a: None if a is None else a.ctypes.data_as(C.POINTER(C.c_double))
    tau = None if tau is None else np.ascontiguousarray(tau, dtype=np.float64)
    fext = None if fext is None else np.ascontiguousarray(fext, dtype=np.float64)
    q2, dq2 = np.zeros_like(q), np.zeros_like(dq)
    cnt = np.zeros(n, dtype=np.int32)
    body = -np.ones((n, maxc), dtype=np.int32)
    data = np.zeros((n, maxc, 10), dtype=np.float32)
    rc = L.emu_substep(C.byref(cm), C.byref(ct), int(f64), n, dp(q), dp(dq), dp(tau), dp(fext), lcp_mode, pgs_iters,
                       dp(q2), dp(dq2), cnt.ctypes.data_as(C.POINTER(C.c_int32)), body.ctypes.data_as(C.POINTER(C.c_int32)),
                       data.ctypes.data_as(C.POINTER(C.c_float)), maxc, int(variant),
                       None if hints is None else hints.ctypes.data_as(C.POINTER(C.c_uint64)))
    if rc:
        raise RuntimeError(L.emu_last_error().decode())
    return q2, dq2, cnt, body, data


def substep_body_params(model, task, q, dq, tau=None, mass=None, friction=None, f64=False, lcp_mode=0, pgs_iters=30, maxc=8):
    """One sub-step of the loop kernel with per-world bodynode masses / friction coefficients [n, n_bodies]."""
    L = lib()
    cm, ct = pack_model(model), pack_task(task)
    q = np.ascontiguousarray(q, dtype=np.float64)
    dq = np.ascontiguousarray(dq, dtype=np.float64)
    n, nd = q.shape
    dp = lambda a: None if a is None else a.ctypes.data_as(C.POINTER(C.c_double))
    tau, mass, friction = (None if a is None else np.ascontiguousarray(a, dtype=np.float64) for a in (tau, mass, friction))
    q2, dq2 = np.zeros_like(q), np.zeros_like(dq)
    cnt = np.zeros(n, dtype=np.int32)
    body = -np.ones((n, maxc), dtype=np.int32)
    data = np.zeros((n, maxc, 10), dtype=np.float32)
    rc = L.emu_substep_body_params(C.byref(cm), C.byref(ct), int(f64), n, dp(q), dp(dq), dp(tau), dp(mass), dp(friction),
                                   lcp_mode, pgs_iters, dp(q2), dp(dq2), cnt.ctypes.data_as(C.POINTER(C.c_int32)),
                                   body.ctypes.data_as(C.POINTER(C.c_int32)), data.ctypes.data_as(C.POINTER(C.c_float)), maxc)
    if rc:
        raise RuntimeError(L.emu_last_error().decode())
    return q2, dq2, cnt, body, data


def lcp(A, b, lo, hi, findex, mode=0, f64=True):
    """Kernel-source LCP solvers on the CPU (mode: 0 dispatch, 1 small<8>, 2 bpp_local, 3 dantzig,
    4 small<4>, 5 small<6>).  Returns (x, failed)."""
    L = lib()
    A = np.ascontiguousarray(A, dtype=np.float64)
    n = A.shape[0]
    b, lo, hi = (np.ascontiguousarray(v, dtype=np.float64) for v in (b, lo, hi))
    fi = np.ascontiguousarray(findex, dtype=np.int32)
    x = np.zeros(n)
    dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    rc = L.emu_lcp(int(f64), n, dp(A), dp(b), dp(lo), dp(hi), fi.ctypes.data_as(C.POINTER(C.c_int32)), int(mode), dp(x))
    return x, rc


def task_kind(model, task, q, dq, aux=None, a2=None, f64=True, do_reset=False, seed=0):
    """obs / reward / done of the contact-free task kinds (csrc/task_kinds.cuh) for given post-step states; with
    do_reset the states are first replaced by the kernel's reset_model() draws.  Returns (obs, reward, done, q, dq, aux)."""
    L = lib()
    cm, ct = pack_model(model), pack_task(task)
    q = np.ascontiguousarray(q, dtype=np.float64).copy()
    dq = np.ascontiguousarray(dq, dtype=np.float64).copy()
    n = q.shape[0]
    aux = np.zeros((n, 3)) if aux is None else np.ascontiguousarray(aux, dtype=np.float64).copy()
    a2 = np.zeros(n) if a2 is None else np.ascontiguousarray(a2, dtype=np.float64)
    obs = np.zeros((n, task.n_obs), dtype=np.float32)
    rew = np.zeros(n)
    done = np.zeros(n, dtype=np.int32)
    dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    rc = L.emu_task_kind(C.byref(cm), C.byref(ct), int(f64), n, dp(q), dp(dq), dp(aux), dp(a2), int(do_reset), C.c_uint64(seed),
                         obs.ctypes.data_as(C.POINTER(C.c_float)), dp(rew), done.ctypes.data_as(C.POINTER(C.c_int32)))
    if rc:
        raise RuntimeError(L.emu_last_error().decode())
    return obs, rew, done.astype(bool), q, dq, aux
