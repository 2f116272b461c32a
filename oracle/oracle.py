"""ctypes wrapper around oracle/libdart_oracle.so (fp64 CPU restatement of the DART step).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  PARITY UNPINNED for the physics (see dart_oracle.c).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from dart_env_b200.cstructs import CModel, CTask, Task, pack_model, pack_task
from dart_env_b200.skel import Model

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "libdart_oracle.so")
    src = os.path.join(_HERE, "dart_oracle.c")
    hdr = os.path.join(_HERE, "..", "include", "dartb.h")
    stale = (not os.path.exists(so)) or any(
        os.path.exists(p) and os.path.getmtime(p) > os.path.getmtime(so) for p in (src, hdr))
    if force or stale:
        subprocess.check_call(["make", "-C", _HERE, "-B", "libdart_oracle.so"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        P, D, I = C.c_void_p, C.c_double, C.c_int
        dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int)
        sig = {
            "orc_create": (P, [C.POINTER(CModel)]), "orc_destroy": (None, [P]),
            "orc_num_dofs": (I, [P]), "orc_num_bodies": (I, [P]),
            "orc_set_option": (None, [P, I, D]), "orc_set_mass": (None, [P, I, D]),
            "orc_set_friction": (None, [P, I, D]),
            "orc_set_state": (None, [P, dp, dp]), "orc_get_state": (None, [P, dp, dp]),
            "orc_set_forces": (None, [P, dp]), "orc_step": (None, [P]), "orc_reset": (None, [P]),
            "orc_body_transform": (None, [P, I, dp]), "orc_body_com": (None, [P, I, dp]),
            "orc_body_com_spatial_velocity": (None, [P, I, dp]), "orc_add_ext_force": (None, [P, I, dp]),
            "orc_num_contacts": (I, [P]), "orc_get_contact": (None, [P, I, ip, dp, dp, dp, dp]),
            "orc_limit_active": (I, [P, I]), "orc_shape_gap": (D, [P, I]), "orc_shape_tilt": (D, [P, I]), "orc_lcp_rows": (I, [P]),
            "orc_get_lcp": (None, [P, dp, dp, dp, dp, dp, ip]), "orc_lcp_failed": (I, [P]),
            "orc_mass_matrix": (None, [P, dp]), "orc_forward_dynamics": (None, [P, dp]),
            "orc_energy": (D, [P]),
            "orc_solve_lcp_dantzig": (I, [I, dp, dp, dp, dp, dp, dp, ip]),
            "orc_solve_lcp_pgs": (I, [I, dp, dp, dp, dp, dp, ip, I]),
            "orc_reset_uniform": (C.c_float, [C.c_uint64, C.c_int64, C.c_uint32, I]),
            "orc_env_create": (P, [C.POINTER(CModel), C.POINTER(CTask), C.c_uint64, C.c_int64]),
            "orc_env_destroy": (None, [P]), "orc_env_world": (P, [P]),
            "orc_env_obs": (None, [P, dp]), "orc_env_reset": (None, [P, dp]),
            "orc_env_step": (None, [P, dp, dp, dp, ip]),
            "orc_bench": (D, [C.POINTER(CModel), C.POINTER(CTask), I, I, I, C.c_uint64, dp]),
        }
        for name, (res, args) in sig.items():
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args
        _LIB = L
    return _LIB


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int))


class OracleWorld:
    """One fp64 world: the pydart2 World/Skeleton surface the reference envs use."""

    def __init__(self, model: Model, _handle=None, _owner=None):
        self.model = model
        self._cm = pack_model(model)
        self._L = lib()
        self._owner = _owner
        self._h = _handle if _handle is not None else self._L.orc_create(C.byref(self._cm))
        self.nd = self._L.orc_num_dofs(self._h)
        self.nb = self._L.orc_num_bodies(self._h)

    def __del__(self):
        if getattr(self, "_h", None) and self._owner is None:
            self._L.orc_destroy(self._h)
            self._h = None

    # state
    def set_state(self, q=None, dq=None):
        q = None if q is None else np.ascontiguousarray(q, dtype=np.float64)
        dq = None if dq is None else np.ascontiguousarray(dq, dtype=np.float64)
        self._L.orc_set_state(self._h, None if q is None else _dp(q), None if dq is None else _dp(dq))

    def get_state(self):
        q, dq = np.zeros(self.nd), np.zeros(self.nd)
        self._L.orc_get_state(self._h, _dp(q), _dp(dq))
        return q, dq

    def set_forces(self, tau):
        tau = np.ascontiguousarray(tau, dtype=np.float64)
        self._L.orc_set_forces(self._h, _dp(tau))

    def step(self):
        self._L.orc_step(self._h)

    def reset(self):
        self._L.orc_reset(self._h)

    def set_option(self, key, val):
        self._L.orc_set_option(self._h, int(key), float(val))

    def set_friction(self, body, mu):
        self._L.orc_set_friction(self._h, int(body), float(mu))

    def set_mass(self, body, mass):
        self._L.orc_set_mass(self._h, int(body), float(mass))

    # queries
    def body_transform(self, i):
        t = np.zeros(12)
        self._L.orc_body_transform(self._h, i, _dp(t))
        T = np.eye(4)
        T[:3, :4] = t.reshape(3, 4)
        return T

    def body_com(self, i):
        c = np.zeros(3)
        self._L.orc_body_com(self._h, i, _dp(c))
        return c

    def body_com_spatial_velocity(self, i):
        v = np.zeros(6)
        self._L.orc_body_com_spatial_velocity(self._h, i, _dp(v))
        return v

    def add_ext_force(self, i, f):
        f = np.ascontiguousarray(f, dtype=np.float64)
        self._L.orc_add_ext_force(self._h, i, _dp(f))

    def contacts(self):
        out = []
        for i in range(self._L.orc_num_contacts(self._h)):
            body, depth = C.c_int(), C.c_double()
            p, n, f = np.zeros(3), np.zeros(3), np.zeros(3)
            self._L.orc_get_contact(self._h, i, C.byref(body), _dp(p), _dp(n), C.byref(depth), _dp(f))
            out.append(dict(body=body.value, point=p, normal=n, depth=depth.value, force=f))
        return out

    def shape_gaps(self):
        """per robot shape: (distance - radius) to the nearest static, and capsule end-height difference"""
        n = len(self.model.shapes)
        return (np.array([self._L.orc_shape_gap(self._h, i) for i in range(n)]),
                np.array([self._L.orc_shape_tilt(self._h, i) for i in range(n)]))

    def limit_active(self):
        return np.array([self._L.orc_limit_active(self._h, d) for d in range(self.nd)])

    def lcp(self):
        n = self._L.orc_lcp_rows(self._h)
        A, x, b, lo, hi = np.zeros((n, n)), np.zeros(n), np.zeros(n), np.zeros(n), np.zeros(n)
        fi = np.zeros(n, dtype=np.int32)
        if n:
            self._L.orc_get_lcp(self._h, _dp(A), _dp(x), _dp(b), _dp(lo), _dp(hi), _ip(fi))
        return dict(A=A, x=x, b=b, lo=lo, hi=hi, findex=fi)

    def lcp_failed(self):
        return bool(self._L.orc_lcp_failed(self._h))

    def mass_matrix(self):
        M = np.zeros((self.nd, self.nd))
        self._L.orc_mass_matrix(self._h, _dp(M))
        return M

    def forward_dynamics(self):
        dd = np.zeros(self.nd)
        self._L.orc_forward_dynamics(self._h, _dp(dd))
        return dd

    def energy(self):
        return self._L.orc_energy(self._h)


class OracleEnv:
    """One env = world + task layer (restates hopper.py / walker2d.py / half_cheetah.py /
    snake_7link.py step(), _get_obs(), reset_model())."""

    def __init__(self, model: Model, task: Task, seed: int = 0, world_id: int = 0):
        self._L = lib()
        self.model, self.task = model, task
        self._cm, self._ct = pack_model(model), pack_task(task)
        self._h = self._L.orc_env_create(C.byref(self._cm), C.byref(self._ct), seed, world_id)
        self.world = OracleWorld(model, _handle=self._L.orc_env_world(self._h), _owner=self)

    def __del__(self):
        if getattr(self, "_h", None):
            self._L.orc_env_destroy(self._h)
            self._h = None

    def reset(self):
        obs = np.zeros(self.task.n_obs)
        self._L.orc_env_reset(self._h, _dp(obs))
        return obs

    def obs(self):
        obs = np.zeros(self.task.n_obs)
        self._L.orc_env_obs(self._h, _dp(obs))
        return obs

    def step(self, action):
        a = np.ascontiguousarray(action, dtype=np.float64)
        obs = np.zeros(self.task.n_obs)
        r, d = C.c_double(), C.c_int()
        self._L.orc_env_step(self._h, _dp(a), _dp(obs), C.byref(r), C.byref(d))
        return obs, r.value, bool(d.value)


def solve_lcp_dantzig(A, b, lo, hi, findex):
    A = np.ascontiguousarray(A, dtype=np.float64)
    n = A.shape[0]
    x, w = np.zeros(n), np.zeros(n)
    b = np.ascontiguousarray(b, dtype=np.float64)
    lo = np.array(lo, dtype=np.float64)
    hi = np.array(hi, dtype=np.float64)
    fi = np.ascontiguousarray(findex, dtype=np.int32)
    fail = lib().orc_solve_lcp_dantzig(n, _dp(A), _dp(x), _dp(b), _dp(w), _dp(lo), _dp(hi), _ip(fi))
    return x, w, lo, hi, fail


def solve_lcp_pgs(A, b, lo, hi, findex, iters):
    A = np.ascontiguousarray(A, dtype=np.float64)
    n = A.shape[0]
    x = np.zeros(n)
    b = np.ascontiguousarray(b, dtype=np.float64)
    lo = np.ascontiguousarray(lo, dtype=np.float64)
    hi = np.ascontiguousarray(hi, dtype=np.float64)
    fi = np.ascontiguousarray(findex, dtype=np.int32)
    lib().orc_solve_lcp_pgs(n, _dp(A), _dp(x), _dp(b), _dp(lo), _dp(hi), _ip(fi), int(iters))
    return x


def reset_uniform(seed, world, episode, i):
    return float(lib().orc_reset_uniform(seed, world, episode, i))


def cpu_bench(model: Model, task: Task, n_worlds: int, n_steps: int, n_threads: int, seed: int = 0):
    cm, ct = pack_model(model), pack_task(task)
    cs = C.c_double()
    sps = lib().orc_bench(C.byref(cm), C.byref(ct), n_worlds, n_steps, n_threads, seed, C.byref(cs))
    return sps, cs.value
