"""Batched DartEnv: the host-side mirror of the reference's gym/envs/dart/dart_env.py:DartEnv.

Same constructor arguments (where meaningful), same reset / step / seed / set_state /
state_vector / dt / do_simulation surface and Box spaces, but `num_envs` independent worlds are
stepped by one CUDA launch through the C-ABI (libdartb.so) instead of one pydart2 World.step()
per Python call.  No GL, no pydart2.  There is no CPU fallback: constructing an env without a
CUDA device raises.

Return types
  num_envs == 1 and batched=False (the default for gym.make-style use): the reference's own
      types — obs float64 ndarray [nobs], reward float, done bool, info dict (test_envs.py:27-28).
  batched=True: VectorEnv conventions (gym/vector/sync_vector_env.py:44-47,73-84) —
      obs [N, nobs], rewards [N], dones [N] with auto-reset; `output="torch"` keeps everything on
      the GPU (zero copies), `output="numpy"` returns host arrays through pinned buffers.
"""
from __future__ import annotations

import ctypes as C
import os
import sys
from typing import Optional, Sequence

import numpy as np
import torch

from . import capi
from .cstructs import Task
from .engine import Engine
from .skel import load_model
from .spaces import Box, batch_space


class _PinnedOutPool:
    """Output arrays of the batched host path.  VectorEnv(copy=True) hands the caller arrays it may keep
    (gym/vector/sync_vector_env.py:44-47,83); here the step kernel writes the reference's return types (float32 obs,
    float64 rewards, bool dones) straight into page-locked numpy arrays over PCIe, so "a fresh copy" is a slot of this
    pool that nobody references any more (CPython refcounts: any array or view the caller still holds keeps its slot
    out of circulation).  No np.empty, no memcpy, no conversion pass per step."""

    def __init__(self, engine, n: int, n_obs: int, with_trunc: bool, max_slots: int = 64, alloc=None):
        self.engine = engine
        self._alloc = alloc or self._alloc_pinned
        self.n, self.n_obs, self.with_trunc, self.max_slots = n, n_obs, with_trunc, max_slots
        a16 = lambda v: (v + 15) // 16 * 16
        self.off_rew = a16(n * n_obs * 4)
        self.off_done = self.off_rew + a16(n * 8)
        self.off_trunc = self.off_done + a16(n)
        self.nbytes = self.off_trunc + (a16(n) if with_trunc else 0)
        self.slots = []
        self._next = 0

    def _alloc_pinned(self, nbytes):
        """page-locked block registered with the handle (zero-copy without per-step driver queries)"""
        t = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
        capi.check(self.engine.L.dartb_register_host(self.engine.h, C.c_void_p(t.data_ptr()), nbytes))
        return t, t.data_ptr()

    def _new_slot(self):
        t, p = self._alloc(self.nbytes)
        base = (C.c_char * self.nbytes).from_address(p)
        n, no = self.n, self.n_obs
        obs1 = np.frombuffer(base, dtype=np.float32, count=n * no, offset=0)
        obs = obs1.reshape(n, no)
        rew = np.frombuffer(base, dtype=np.float64, count=n, offset=self.off_rew)
        done = np.frombuffer(base, dtype=np.bool_, count=n, offset=self.off_done)
        trunc = np.frombuffer(base, dtype=np.bool_, count=n, offset=self.off_trunc) if self.with_trunc else None
        ptrs = (C.c_void_p(p), C.c_void_p(p + self.off_rew), C.c_void_p(p + self.off_done),
                C.c_void_p(p + self.off_trunc) if self.with_trunc else None)
        slot = {"keep": (t, base), "obs": obs, "rew": rew, "done": done, "trunc": trunc, "ptrs": ptrs,
                "watched": [a for a in (obs1, obs, rew, done, trunc) if a is not None]}
        del obs1, obs, rew, done, trunc   # the idle refcounts must not include this frame's locals
        slot["idle"] = [sys.getrefcount(a) for a in slot["watched"]]
        self.slots.append(slot)
        return slot

    def take(self):
        """a slot whose arrays nobody outside this pool references, or None when max_slots are all held"""
        ns = len(self.slots)
        for k in range(ns):
            slot = self.slots[(self._next + k) % ns]
            if [sys.getrefcount(a) for a in slot["watched"]] == slot["idle"]:
                self._next = (self._next + k + 1) % ns
                return slot
        if ns < self.max_slots:
            return self._new_slot()
        return None


class DartEnv:
    """Superclass for all (batched) Dart environments."""

    metadata = {"render.modes": []}

    def __init__(self, model_paths, frame_skip, observation_size, action_bounds, dt=0.002, obs_type="parameter",
                 action_type="continuous", visualize=True, disableViewer=False, screen_width=80, screen_height=45, *,
                 task: Optional[Task] = None, num_envs: int = 1, batched: Optional[bool] = None, output: str = "torch",
                 device: int = 0, seed: Optional[int] = None, world_offset: int = 0, auto_reset: Optional[bool] = None,
                 max_episode_steps: int = 0, friction_all: Optional[float] = None, f64: bool = False,
                 collidable: bool = True, copy: bool = True, kernel_variant: Optional[int] = None, contacts: bool = False):
        assert obs_type in ("parameter", "image")
        assert action_type in ("continuous", "discrete")
        if obs_type == "image":
            raise NotImplementedError("pixel observations need the GL viewer (out of scope)")
        if isinstance(model_paths, str):
            model_paths = [model_paths]
        if len(model_paths) < 1:
            raise ValueError("At least one model file is needed.")
        if not model_paths[0].endswith(".skel"):
            raise NotImplementedError("URDF/SDF loading (dart_env.py:56-59) is not implemented; no in-tree env uses it")
        # dart_env.py:54-67: build the world, robot = last skeleton, enforce every limited dof
        self.model = load_model(model_paths[0], dt)
        self.model.enforce_limits()
        self._skel_frictions = [b.friction_coeff for b in self.model.bodies]   # as loaded, before friction_all
        self._body_mass = self._body_mu = None   # per-world bodynode parameters (set_body_params), None = the model's
        if friction_all is not None:
            for b in self.model.bodies:
                b.friction_coeff = float(friction_all)
        if not collidable:  # reacher2d.py:11-14: every bodynode set_collidable(False)
            self.model.shapes, self.model.ground = [], []
        # task=None: physics-only handle; the subclass computes obs / reward / done on the host side
        # (torch ops over the batched state), calling do_simulation() like the reference classes do
        self.fused = task is not None
        if task is None:
            task = Task.physics_only(frame_skip)
        elif task.frame_skip != frame_skip or task.n_obs != observation_size:
            raise ValueError("task does not match frame_skip / observation_size")
        self.task = task
        self.frame_skip = frame_skip
        self.obs_dim = observation_size
        self.act_dim = len(action_bounds[0])
        self.num_envs = int(num_envs)
        self.batched = (self.num_envs > 1) if batched is None else bool(batched)
        if not self.batched and self.num_envs != 1:
            raise ValueError("batched=False needs num_envs == 1")
        if output not in ("torch", "numpy"):
            raise ValueError("output must be 'torch' or 'numpy'")
        self.output = output if self.batched else "numpy"
        self.auto_reset = self.batched if auto_reset is None else bool(auto_reset)
        self.world_offset = int(world_offset)
        self.copy = bool(copy)  # sync_vector_env.py:44-47: return a copy of the observation buffer
        self.disableViewer = True
        self.viewer = None
        # random perturbation (dart_env.py:74-78): off by default, as in the reference
        self.add_perturbation = False
        self.perturbation_parameters = [0.05, 5, 2]  # probability, magnitude, bodyid
        self.perturbation_duration = 40
        self.perturb_force = None
        self._device_index = device
        self._f64 = f64
        self._kernel_variant = kernel_variant
        self._contacts = bool(contacts)   # record collision_result.contacts in the fused step (walker2d.py:38-41)

        self.action_space = Box(np.asarray(action_bounds[1], dtype=np.float64), np.asarray(action_bounds[0], dtype=np.float64))
        high = np.inf * np.ones(self.obs_dim)
        self.observation_space = Box(-high, high)
        self.single_action_space, self.single_observation_space = self.action_space, self.observation_space
        if self.batched:
            self.action_space = batch_space(self.single_action_space, self.num_envs)
            self.observation_space = batch_space(self.single_observation_space, self.num_envs)

        self._seed_value = 0
        self.engine: Optional[Engine] = None
        self.seed(seed)
        self.max_episode_steps = int(max_episode_steps)
        self.metadata = {"render.modes": [], "video.frames_per_second": int(np.round(1.0 / self.dt))}

    # ------------------------------------------------------------------ engine / buffers
    def _build_engine(self):
        if self.engine is not None:
            self.engine.close()
        self.engine = Engine(self.model, self.task, self.num_envs, device=self._device_index, seed=self._seed_value,
                             world_offset=self.world_offset, f64=self._f64, kernel_variant=self._kernel_variant)
        if getattr(self, "max_episode_steps", 0):
            self.engine.set_max_episode_steps(self.max_episode_steps)
        if self._contacts:
            self.engine.set_contacts(True)
        dev, n = self.engine.device, self.num_envs
        self._obs = torch.empty((n, self.obs_dim), dtype=torch.float32, device=dev)
        self._rew = torch.empty((n,), dtype=torch.float32, device=dev)
        self._done = torch.empty((n,), dtype=torch.uint8, device=dev)
        self._act = torch.empty((n, self.act_dim), dtype=torch.float32, device=dev)
        # pinned host mirrors for the host-facing (numpy) path
        self._h_act = torch.empty((n, self.act_dim), dtype=torch.float32).pin_memory()
        self._h_obs = torch.empty((n, self.obs_dim), dtype=torch.float32).pin_memory()
        self._h_rew = torch.empty((n,), dtype=torch.float32).pin_memory()
        self._h_done = torch.empty((n,), dtype=torch.uint8).pin_memory()
        # page-locked output arrays of the host path: dartb_step_host DMAs straight into them
        self._n_obs = torch.empty((n, self.obs_dim), dtype=torch.float32).pin_memory().numpy()
        self._n_rew = torch.empty((n,), dtype=torch.float32).pin_memory().numpy()
        self._n_done = torch.empty((n,), dtype=torch.uint8).pin_memory().numpy()
        self._n_act = torch.empty((n, self.act_dim), dtype=torch.float32).pin_memory().numpy()
        self._c_obs, self._c_rew, self._c_done, self._c_act = (C.c_void_p(x.ctypes.data) for x in (self._n_obs, self._n_rew, self._n_done, self._n_act))
        for x in (self._n_obs, self._n_rew, self._n_done, self._n_act):   # zero-copy without per-step driver queries
            capi.check(self.engine.L.dartb_register_host(self.engine.h, C.c_void_p(x.ctypes.data), x.nbytes))
        self._pool = None   # pinned output slots of the batched host path (built on first use)
        # the host (numpy) path is synchronous: it keeps using the stream that was current when the engine was built
        # instead of asking torch for the current stream on every step
        self._host_stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)

    @property
    def max_episode_steps(self):
        return self._max_episode_steps

    @max_episode_steps.setter
    def max_episode_steps(self, n):
        self._max_episode_steps = int(n)
        if self.engine is not None:
            self.engine.set_max_episode_steps(self._max_episode_steps)

    def seed(self, seed=None):
        """dart_env.py:117-119.  Reset noise comes from a counter-based generator keyed by
        (seed, global world id, episode), so results do not depend on sharding."""
        if isinstance(seed, (list, tuple, np.ndarray)):
            # VectorEnv.seed([s_0, ...]) (gym/vector/sync_vector_env.py:50-57): world i draws what a single env seeded s_i draws
            seeds = np.ascontiguousarray([int(s) & 0xFFFFFFFFFFFFFFFF for s in seed], dtype=np.uint64)
            if seeds.shape != (self.num_envs,):
                raise ValueError("seed list must have one entry per env")
            self.seed(int(seeds[0]))
            capi.check(self.engine.L.dartb_seed_worlds(self.engine.h, seeds.ctypes.data))
            return [int(s) for s in seeds]
        if seed is None:
            seed = int.from_bytes(os.urandom(4), "little")
        self._seed_value = int(seed) & 0xFFFFFFFFFFFFFFFF
        self.np_random = np.random.RandomState(self._seed_value & 0xFFFFFFFF)
        if self.engine is None:
            self._build_engine()
        else:   # like the reference, seeding does not touch the physics state
            capi.check(self.engine.L.dartb_seed(self.engine.h, self._seed_value))
        return [self._seed_value]

    def close(self):
        if self.engine is not None:
            self.engine.close()
            self.engine = None

    # ------------------------------------------------------------------ reference surface
    @property
    def dart_world(self):
        """pydart2-shaped view of the batched world (dart_env.py:54-62): dt, step(), reset(), skeletons,
        collision_result.contacts."""
        if getattr(self, "_world_view", None) is None:
            from .pydart_view import WorldView
            self._world_view = WorldView(self)
        return self._world_view

    @property
    def robot_skeleton(self):
        """`self.robot_skeleton = self.dart_world.skeletons[-1]` (dart_env.py:62): q, dq, ndofs, set_positions,
        set_velocities, set_forces, q_lower / q_upper, bodynodes[i].com() / to_world() / com_spatial_velocity() ..."""
        return self.dart_world.robot

    @property
    def dt(self):
        return self.model.dt * self.frame_skip

    def reset(self):
        self.perturbation_duration = 0
        if self.perturb_force is not None:
            self._perturb_count.zero_()
        if not self.fused:
            return self._out_obs(self.reset_model())
        obs = self.engine.reset(None, self._obs)
        return self._out_obs(obs)

    def reset_model(self):
        """Reset the robot degrees of freedom (qpos and qvel).  Implemented by host-side task classes
        (dart_env.py:125-130 of the reference); fused envs reset inside the kernel."""
        raise NotImplementedError

    def set_state(self, qpos, qvel):
        q = torch.as_tensor(np.asarray(qpos, dtype=np.float64).reshape(self.num_envs, -1), device=self.engine.device)
        v = torch.as_tensor(np.asarray(qvel, dtype=np.float64).reshape(self.num_envs, -1), device=self.engine.device)
        assert q.shape == (self.num_envs, self.model.n_dofs) and v.shape == q.shape
        self.engine.set_state(q.contiguous(), v.contiguous())

    def set_body_params(self, mass=None, friction=None):
        """`bodynodes[i].set_mass(m)` / `.set_friction_coeff(mu)` per world (snake_7link.py:115-120): arrays
        [num_envs, n_bodynodes] or None (= the skeleton's value); both None returns to the shared model.  While set, the
        batch runs on the topology-generic kernels (`engine.kernel_name`)."""
        mass = None if mass is None else np.array(mass, dtype=np.float64).reshape(self.num_envs, -1)
        friction = None if friction is None else np.array(friction, dtype=np.float64).reshape(self.num_envs, -1)
        self.engine.set_body_params(mass, friction)
        self._body_mass, self._body_mu = mass, friction

    def _body_param_array(self, which):
        cur = self._body_mass if which == "mass" else self._body_mu
        if cur is not None:
            return cur.copy()
        vals = [b.mass if which == "mass" else b.friction_coeff for b in self.model.bodies]
        return np.tile(np.asarray(vals, dtype=np.float64), (self.num_envs, 1))

    def set_state_vector(self, state):
        state = np.asarray(state, dtype=np.float64).reshape(self.num_envs, -1)
        nd = self.model.n_dofs
        self.set_state(state[:, :nd], state[:, nd:])

    def state_vector(self):
        q, dq = self.engine.get_state(torch.float64)
        s = torch.cat([q, dq], dim=1).cpu().numpy()
        return s if self.batched else s[0]

    def do_simulation(self, tau, n_frames):
        """dart_env.py:158-175: n_frames x {set_forces(tau); world.step()} (perturbation off)."""
        if isinstance(tau, torch.Tensor):
            t = tau.reshape(self.num_envs, -1).to(self.engine.device).contiguous()
        else:
            t = torch.as_tensor(np.asarray(tau, dtype=np.float64).reshape(self.num_envs, -1), device=self.engine.device).contiguous()
        fext = None
        if self.add_perturbation:
            # dart_env.py:159-172, per world: when the countdown is 0 the force is cleared and, with
            # probability p, a new +-magnitude force along x or y is drawn; it is applied at the origin of
            # bodynodes[bodyid] (add_ext_force) on every sub-step of this call.
            n, dev = self.num_envs, self.engine.device
            if self.perturb_force is None:
                self.perturb_force = torch.zeros((n, 3), dtype=t.dtype, device=dev)
                self._perturb_count = torch.zeros((n,), dtype=torch.int64, device=dev)
            zero = self._perturb_count == 0
            self.perturb_force[zero] = 0
            p, mag, body = self.perturbation_parameters
            draw = zero & (torch.rand(n, device=dev) < p)
            axis = torch.randint(0, 2, (n,), device=dev)
            sign = torch.randint(0, 2, (n,), device=dev) * 2 - 1
            newf = torch.zeros((n, 3), dtype=t.dtype, device=dev)
            newf[torch.arange(n, device=dev), axis] = (sign * mag).to(t.dtype)
            self.perturb_force = torch.where(draw[:, None], newf, self.perturb_force)
            self._perturb_count = torch.where(zero, self._perturb_count, self._perturb_count - 1)
            fext = torch.zeros((n, self.model.n_bodies, 3), dtype=t.dtype, device=dev)
            fext[:, int(body)] = self.perturb_force
        for _ in range(n_frames):
            self.engine.substep(t, fext)

    def step(self, a):
        if not self.fused:
            raise NotImplementedError("host-side task classes implement step()")
        if self.add_perturbation:
            raise NotImplementedError("add_perturbation (dart_env.py:159-172) is applied by do_simulation(); the fused "
                                      "step() of this env does not draw perturbation forces")
        if isinstance(a, torch.Tensor) and a.is_cuda:
            act = a.reshape(self.num_envs, self.act_dim).to(torch.float32).contiguous()
            self.engine.step(act, self._obs, self._rew, self._done, self.auto_reset)
        else:
            # host path: ONE library call (launch + sync); the kernel reads the actions from, and writes the results to,
            # page-locked host memory that is registered with the handle
            np.copyto(self._n_act, np.asarray(a).reshape(self.num_envs, self.act_dim), casting="same_kind")
            act_ptr = self._c_act
            eng = self.engine
            if self.batched and self.copy:
                # reference return types in one library call, written by the kernel into an unreferenced pinned slot
                if self._pool is None or self._pool.with_trunc != bool(self._max_episode_steps):
                    self._pool = _PinnedOutPool(eng, self.num_envs, self.obs_dim, bool(self._max_episode_steps))
                slot = self._pool.take()
                if slot is not None:
                    po, pr, pd, pt = slot["ptrs"]
                    rc = eng.L.dartb_step_host_gym(eng.h, act_ptr, po, pr, pd, pt, int(self.auto_reset), self._host_stream)
                    if rc:
                        capi.check(rc)
                    return slot["obs"], slot["rew"], slot["done"], ({"TimeLimit.truncated": slot["trunc"]} if pt is not None else {})
                # (every slot is still held by the caller: plain pageable arrays, filled from the staging block)
                n = self.num_envs
                obs = np.empty((n, self.obs_dim), dtype=np.float32)
                rew = np.empty((n,), dtype=np.float64)
                done = np.empty((n,), dtype=np.bool_)
                trunc = np.empty((n,), dtype=np.bool_) if self._max_episode_steps else None
                rc = eng.L.dartb_step_host_gym(eng.h, act_ptr, obs.ctypes.data, rew.ctypes.data, done.ctypes.data,
                                               trunc.ctypes.data if trunc is not None else None, int(self.auto_reset),
                                               self._host_stream)
                if rc:
                    capi.check(rc)
                return obs, rew, done, ({"TimeLimit.truncated": trunc} if trunc is not None else {})
            # (the output arrays are this env's own page-locked buffers: their pointers are cached)
            rc = eng.L.dartb_step_host(eng.h, act_ptr, self._c_obs, self._c_rew, self._c_done, int(self.auto_reset),
                                       self._host_stream)
            if rc:
                capi.check(rc)
            d = self._n_done  # bit 0 done, bit 1 truncated by the time limit (no extra transfer)
            if self.batched:
                infos = {}
                if self._max_episode_steps:
                    infos["TimeLimit.truncated"] = (d & 2) != 0
                return self._n_obs, self._n_rew.astype(np.float64), d != 0, infos
            info = {}
            done = bool(d[0] != 0)
            if self._max_episode_steps and done:
                info["TimeLimit.truncated"] = bool(d[0] & 2)
            return self._n_obs[0].astype(np.float64), float(self._n_rew[0]), done, info
        if self.batched and self.output == "torch":
            return self._obs, self._rew, self._done.bool(), {}
        # a CUDA-tensor action with numpy output: copy the results out
        self._h_obs.copy_(self._obs, non_blocking=True)
        self._h_rew.copy_(self._rew, non_blocking=True)
        self._h_done.copy_(self._done, non_blocking=True)
        torch.cuda.current_stream(self.engine.device).synchronize()
        if self.batched:
            infos = {}
            if self._max_episode_steps:
                infos["TimeLimit.truncated"] = self.engine.truncated().cpu().numpy().astype(bool)
            return (self._h_obs.numpy().copy(), self._h_rew.numpy().astype(np.float64), self._h_done.numpy().astype(np.bool_),
                    infos)
        info = {}
        done = bool(self._h_done[0].item())
        if self._max_episode_steps and done:
            info["TimeLimit.truncated"] = bool(self.engine.truncated()[0].item())
        return self._h_obs.numpy()[0].astype(np.float64), float(self._h_rew[0].item()), done, info

    def _out_obs(self, obs):
        if self.batched and self.output == "torch":
            return obs
        o = obs.cpu().numpy()
        return o.copy() if self.batched else o[0].astype(np.float64)

    # viewer entry points of the reference that have no meaning without GL
    def render(self, mode="human", close=False):
        if close:
            return None
        raise NotImplementedError("rendering needs the GL viewer (out of scope, SURVEY.md §2 row 11)")

    def viewer_setup(self):
        pass

    def contacts(self):
        """world.collision_result.contacts of the last sub-step: (count[N], body[N,C], data[N,C,10]).
        After a fused step() this needs DartEnv(contacts=True); do_simulation() always records."""
        return self.engine.contacts()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
