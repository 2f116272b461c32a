"""Thin Python handle over the C-ABI: torch tensors carry the batched arrays, every call is
enqueued on torch's current CUDA stream (no hidden syncs)."""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import capi
from .cstructs import (OPT_FRICTION_ALL, OPT_LCP_MODE, OPT_PGS_ITERS, Task, pack_model, pack_task)
from .skel import Model

OPT_MAX_EPISODE_STEPS = 4
OPT_KERNEL_VARIANT = 5   # -1 auto, 0 one world per thread, 1 loop / generic, 2 lane-cooperative
OPT_WORLDS_PER_WARP = 6
OPT_CONTACTS = 7         # 1: the fused step records collision_result.contacts of its last sub-step


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


class Engine:
    """N worlds of one skeleton on one GPU (one dartb handle)."""

    def __init__(self, model: Model, task: Task, num_worlds: int, device: int = 0, seed: int = 0,
                 world_offset: int = 0, f64: bool = False, kernel_variant: Optional[int] = None):
        if not torch.cuda.is_available():
            raise capi.DartbError("no CUDA device: the B200 engine has no CPU fallback")
        self.L = capi.load()
        self.model, self.task = model, task
        self.n, self.nd, self.nbodies = int(num_worlds), model.n_dofs, model.n_bodies
        self.n_act, self.n_obs = task.n_act, task.n_obs
        self.device = torch.device("cuda", device)
        self.f64 = f64
        self._cm, self._ct = pack_model(model), pack_task(task)
        h = C.c_void_p()
        create = self.L.dartb_create_f64 if f64 else self.L.dartb_create
        capi.check(create(C.byref(self._cm), C.byref(self._ct), self.n, device, seed, world_offset, C.byref(h)))
        self.h = h
        self.max_contacts = self.L.dartb_max_contacts(self.h)
        if kernel_variant is not None:
            self.set_option(OPT_KERNEL_VARIANT, kernel_variant)

    def close(self):
        if getattr(self, "h", None):
            self.L.dartb_destroy(self.h)
            self.h = None

    __del__ = close

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _chk(self, t: torch.Tensor, shape, dtype):
        if t.device != self.device or t.dtype != dtype or tuple(t.shape) != tuple(shape) or not t.is_contiguous():
            raise ValueError("expected contiguous %s tensor of shape %s on %s, got %s %s on %s"
                             % (dtype, tuple(shape), self.device, t.dtype, tuple(t.shape), t.device))

    # --- options
    def set_option(self, key: int, value: float):
        capi.check(self.L.dartb_set_option(self.h, key, float(value)))

    def set_lcp(self, mode: int, pgs_iters: Optional[int] = None):
        self.set_option(OPT_LCP_MODE, mode)
        if pgs_iters is not None:
            self.set_option(OPT_PGS_ITERS, pgs_iters)

    def set_friction_all(self, mu: float):
        self.set_option(OPT_FRICTION_ALL, mu)

    def set_contacts(self, on: bool):
        """walker2d.py:38-41 reads world.collision_result.contacts after a step; the fused step records them
        only when asked (356 B per world per step).  Engine.substep() always records."""
        self.set_option(OPT_CONTACTS, 1 if on else 0)

    def set_max_episode_steps(self, n: int):
        self.set_option(OPT_MAX_EPISODE_STEPS, n)

    @property
    def kernel_name(self) -> str:
        return self.L.dartb_kernel_name(self.h).decode()

    @property
    def launch_count(self) -> int:
        return int(self.L.dartb_launch_count(self.h))

    # --- state
    def set_state(self, q: Optional[torch.Tensor], dq: Optional[torch.Tensor]):
        for t in (q, dq):
            if t is not None:
                self._chk(t, (self.n, self.nd), t.dtype)
        dt = (q if q is not None else dq).dtype
        fn = self.L.dartb_set_state_f64 if dt == torch.float64 else self.L.dartb_set_state
        capi.check(fn(self.h, _ptr(q), _ptr(dq), self._stream()))

    def get_state(self, dtype=torch.float32):
        q = torch.empty((self.n, self.nd), dtype=dtype, device=self.device)
        dq = torch.empty_like(q)
        fn = self.L.dartb_get_state_f64 if dtype == torch.float64 else self.L.dartb_get_state
        capi.check(fn(self.h, _ptr(q), _ptr(dq), self._stream()))
        return q, dq

    def set_obs_peers(self, peer_ptrs, float_offset: int):
        """Fused observation all-gather: step() also stores its observation rows into the buffers at `peer_ptrs`
        (device pointers valid on this GPU, e.g. symmetric-memory buffer_ptrs) at `float_offset`.  [] switches it off."""
        arr = (C.c_void_p * max(1, len(peer_ptrs)))(*[C.c_void_p(int(p)) for p in peer_ptrs])
        capi.check(self.L.dartb_set_obs_peers(self.h, arr, len(peer_ptrs), int(float_offset)))

    def set_body_params(self, mass=None, friction=None):
        """Per-world bodynode masses / friction coefficients, numpy [n_worlds, n_bodynodes] each or None (= the model's):
        `bodynodes[i].set_mass`, `.set_friction_coeff` on single worlds of the batch (snake_7link.py:115-120).  Both None
        returns to the shared model.  Synchronising; the batch then runs on the loop kernels."""
        import numpy as np
        arrs = []
        for a in (mass, friction):
            if a is None:
                arrs.append(None)
                continue
            a = np.ascontiguousarray(a, dtype=np.float64)
            if a.ndim != 2 or a.shape[0] != self.n:
                raise ValueError("expected [n_worlds, n_bodynodes]")
            if arrs and arrs[0] is not None and arrs[0].shape != a.shape:
                raise ValueError("mass and friction shapes differ")
            arrs.append(a)
        nb = self.nbodies
        for a in arrs:
            if a is not None and a.shape[1] != nb:
                raise ValueError(f"expected [n_worlds, {nb}] (DART bodynode order)")
        ptr = [a.ctypes.data_as(C.c_void_p) if a is not None else None for a in arrs]
        capi.check(self.L.dartb_set_body_params(self.h, ptr[0], ptr[1]))

    def set_randomize(self, mass_range: float = 0.0, friction_range: float = 0.0):
        """snake_7link.py:115-120 inside the kernel: at every reset of a world (reset() and the auto-reset of step()) its
        bodynode masses become original + U(-mass_range, mass_range) and its friction coefficients original +
        U(-friction_range, friction_range), clipped at 0.  0 switches a redraw off."""
        self.set_option(8, float(mass_range))       # DARTB_OPT_RANDOMIZE_MASS
        self.set_option(9, float(friction_range))   # DARTB_OPT_RANDOMIZE_FRICTION

    def get_body_table(self):
        """the per-world table the kernels read: numpy [4 nb + ns, n_worlds] float64 — rows mass, cx, cy, izz per planar
        body, then the friction coefficient per capsule"""
        import numpy as np
        rows = C.c_int32(0)
        capi.check(self.L.dartb_get_body_table(self.h, None, C.byref(rows)))
        out = np.empty((rows.value, self.n), dtype=np.float64)
        capi.check(self.L.dartb_get_body_table(self.h, out.ctypes.data_as(C.c_void_p), None))
        return out

    def set_aux(self, aux: torch.Tensor):
        """per-world task state [n, 3] float64 (the reacher's `self.target`, reacher2d.py:7,57-63)"""
        self._chk(aux, (self.n, 3), torch.float64)
        capi.check(self.L.dartb_set_aux(self.h, _ptr(aux), self._stream()))

    def get_aux(self) -> torch.Tensor:
        out = torch.empty((self.n, 3), dtype=torch.float64, device=self.device)
        capi.check(self.L.dartb_get_aux(self.h, _ptr(out), self._stream()))
        return out

    # --- stepping
    def reset(self, mask: Optional[torch.Tensor] = None, obs: Optional[torch.Tensor] = None) -> torch.Tensor:
        if self.n_obs == 0:  # physics-only handle: world.reset() only
            if mask is not None:
                self._chk(mask, (self.n,), torch.uint8)
            capi.check(self.L.dartb_reset(self.h, _ptr(mask), None, self._stream()))
            return None
        if obs is None:
            obs = torch.empty((self.n, self.n_obs), dtype=torch.float32, device=self.device)
        self._chk(obs, (self.n, self.n_obs), torch.float32)
        if mask is not None:
            self._chk(mask, (self.n,), torch.uint8)
        capi.check(self.L.dartb_reset(self.h, _ptr(mask), _ptr(obs), self._stream()))
        return obs

    def step(self, action: torch.Tensor, obs: torch.Tensor, reward: torch.Tensor, done: torch.Tensor,
             auto_reset: bool = True):
        self._chk(action, (self.n, self.n_act), torch.float32)
        self._chk(obs, (self.n, self.n_obs), torch.float32)
        self._chk(reward, (self.n,), torch.float32)
        self._chk(done, (self.n,), torch.uint8)
        capi.check(self.L.dartb_step(self.h, _ptr(action), _ptr(obs), _ptr(reward), _ptr(done), int(auto_reset),
                                     self._stream()))

    def step_host(self, action, obs, reward, done, auto_reset: bool = True):
        """env.step() on host (numpy) buffers: float32 [n,n_act] in; float32 [n,n_obs], float32 [n],
        uint8 [n] out.  H2D, launch, D2H and the sync happen inside the library (pinned staging)."""
        import numpy as np
        for a, shape, dt in ((action, (self.n, self.n_act), np.float32), (obs, (self.n, self.n_obs), np.float32),
                             (reward, (self.n,), np.float32), (done, (self.n,), np.uint8)):
            if a.shape != shape or a.dtype != dt or not a.flags["C_CONTIGUOUS"]:
                raise ValueError("expected C-contiguous %s array of shape %s" % (dt.__name__, shape))
        capi.check(self.L.dartb_step_host(self.h, C.c_void_p(action.ctypes.data), C.c_void_p(obs.ctypes.data),
                                          C.c_void_p(reward.ctypes.data), C.c_void_p(done.ctypes.data), int(auto_reset),
                                          self._stream()))

    def substep(self, tau: Optional[torch.Tensor], fext: Optional[torch.Tensor] = None):
        dt = (tau if tau is not None else fext)
        dt = torch.float32 if dt is None else dt.dtype
        if tau is not None:
            self._chk(tau, (self.n, self.nd), dt)
        if fext is not None:
            self._chk(fext, (self.n, self.nbodies, 3), dt)
        fn = self.L.dartb_substep_f64 if dt == torch.float64 else self.L.dartb_substep
        capi.check(fn(self.h, _ptr(tau), _ptr(fext), self._stream()))

    def contacts(self):
        cnt = torch.empty((self.n,), dtype=torch.int32, device=self.device)
        body = torch.empty((self.n, self.max_contacts), dtype=torch.int32, device=self.device)
        data = torch.empty((self.n, self.max_contacts, 10), dtype=torch.float32, device=self.device)
        capi.check(self.L.dartb_get_contacts(self.h, _ptr(cnt), _ptr(body), _ptr(data), self._stream()))
        return cnt, body, data

    def truncated(self) -> torch.Tensor:
        out = torch.empty((self.n,), dtype=torch.uint8, device=self.device)
        capi.check(self.L.dartb_get_truncated(self.h, _ptr(out), self._stream()))
        return out
