"""the contact-free planar envs — DartCartPole-v1, DartCartPoleSwingUp-v1,
DartDoubleInvertedPendulumEnv-v1, DartReacher-v1 (reference gym/envs/dart/cart_pole.py, cartpole_swingup.py,
inverted_double_pendulum.py, reacher2d.py) — as SPECS of the fused CUDA task layer.

Each env.step() is one launch of the topology-generic kernel (csrc/kernels.cuh::k_env_step_loop): action ->
frame_skip DART steps -> the env's own obs / reward / done (`dartb_task_t.kind`) -> TimeLimit -> masked auto-reset
with counter-based reset draws, exactly like the four locomotion envs.  Nothing here computes on the host: the
classes only state the constants the reference constructors hard-code."""
from __future__ import annotations

import math

import numpy as np
import torch

from .cstructs import TASK_CARTPOLE, TASK_DOUBLE_PENDULUM, TASK_REACHER2D, TASK_SWINGUP, Task
from .dart_env import DartEnv
from .skel import load_model

_INF = math.inf


class _FusedEnv(DartEnv):
    SKEL, DT, FRAME_SKIP, N_OBS = "", 0.0, 2, 0
    MAX_EPISODE_STEPS = 0

    def _task(self, model) -> Task:
        raise NotImplementedError

    def __init__(self, **kw):
        kw.setdefault("max_episode_steps", 0)
        model = load_model(self.SKEL, self.DT)
        task = self._task(model)
        n_act = task.n_act
        self.control_bounds = np.array([[1.0] * n_act, [-1.0] * n_act])
        DartEnv.__init__(self, self.SKEL, self.FRAME_SKIP, self.N_OBS, self.control_bounds, dt=self.DT, task=task, **kw)


class DartCartPoleEnv(_FusedEnv):
    """cart_pole.py:5-41: tau[0] = 100 a[0] (not clamped), obs [q, dq], reward 1, done unless finite and |q1| <= 0.2;
    reset q0 + U(+-0.01), dq0 + U(+-0.01)."""
    SKEL, DT, FRAME_SKIP, N_OBS, MAX_EPISODE_STEPS = "cartpole.skel", 0.02, 2, 4, 1000
    action_scale = 100

    def _task(self, model):
        return Task(frame_skip=2, act_dof=[0], act_scale=[100.0], act_lo=[-_INF], act_hi=[_INF], n_obs=4, kind=TASK_CARTPOLE,
                    reset_noise=0.01)


class DartCartPoleSwingUpEnv(_FusedEnv):
    """cartpole_swingup.py:7-56: reward 6 - |q1| - 0.01 a^2 - 0.01 |q0|; done |q1| > 8 pi or |dq1| > 25 or |q0| > 5;
    reset q0 + U(+-0.1), dq0 + U(+-0.01), q1 += +-pi."""
    SKEL, DT, FRAME_SKIP, N_OBS, MAX_EPISODE_STEPS = "cartpole_swingup.skel", 0.01, 2, 4, 500
    action_scale = 40

    def _task(self, model):
        return Task(frame_skip=2, act_dof=[0], act_scale=[40.0], act_lo=[-_INF], act_hi=[_INF], n_obs=4, kind=TASK_SWINGUP,
                    reset_noise=0.1, reset_noise_dq=0.01)


class DartDoubleInvertedPendulumEnv(_FusedEnv):
    """inverted_double_pendulum.py:8-71: obs [q0, sin q1:, cos q1:, dq]; reward 10 - dist - vel penalties from the height of
    'weight' above 'cart'; done height <= 1; reset q0 + U(+-0.1), dq0 + 0.1 randn."""
    SKEL, DT, FRAME_SKIP, N_OBS, MAX_EPISODE_STEPS = "inverted_double_pendulum.skel", 0.01, 2, 8, 1000
    action_scale = 40

    def _task(self, model):
        names = [b.name for b in model.bodies]
        return Task(frame_skip=2, act_dof=[0], act_scale=[40.0], act_lo=[-_INF], act_hi=[_INF], n_obs=8,
                    kind=TASK_DOUBLE_PENDULUM, reset_noise=0.1, reset_noise_dq=0.1,
                    probe_body=(names.index("cart"), names.index("weight")))


class DartReacher2dEnv(_FusedEnv):
    """reacher2d.py:5-68 (registered as DartReacher-v1): tau = 200 clamp(a); obs [cos q, sin q, target x z, dq, tip - target];
    reward -|tip - target| - a^2; never done (TimeLimit 50); reset q0 + U(+-0.01), dq0 + U(+-0.005), target redrawn inside
    the 0.2 disc.  Joint Coulomb friction (reacher2d.skel 0.05) enters the LCP as +-friction*dt rows."""
    SKEL, DT, FRAME_SKIP, N_OBS, MAX_EPISODE_STEPS = "reacher2d.skel", 0.01, 2, 11, 50
    action_scale = np.array([200.0, 200.0])

    def _task(self, model):
        tip = model.n_bodies - 1
        return Task(frame_skip=2, act_dof=[0, 1], act_scale=[200.0, 200.0], n_obs=11, kind=TASK_REACHER2D, reset_noise=0.01,
                    reset_noise_dq=0.005, probe_body=(tip, -1), probe_local=(tuple(model.bodies[tip].com), (0.0, 0.0, 0.0)))

    def __init__(self, **kw):
        kw.setdefault("collidable", False)   # reacher2d.py:11-14: set_collidable(False) on every bodynode
        _FusedEnv.__init__(self, **kw)

    # `self.target` of the reference: per-world [N, 3] (world x, y, z) living next to the physics state on the device
    @property
    def target(self):
        t = self.engine.get_aux()
        return t if (self.batched and self.output == "torch") else (t.cpu().numpy() if self.batched else t[0].cpu().numpy())

    @target.setter
    def target(self, value):
        self.set_target(value)

    def set_target(self, value):
        v = torch.as_tensor(np.asarray(value, dtype=np.float64) if not isinstance(value, torch.Tensor) else value,
                            device=self.engine.device).to(torch.float64).reshape(-1, 3)
        self.engine.set_aux(v.expand(self.num_envs, 3).contiguous())


CONTACT_FREE = {"DartReacher-v1": (DartReacher2dEnv, 50),
                "DartCartPole-v1": (DartCartPoleEnv, 1000), "DartCartPoleSwingUp-v1": (DartCartPoleSwingUpEnv, 500),
                "DartDoubleInvertedPendulumEnv-v1": (DartDoubleInvertedPendulumEnv, 1000)}
