"""the contact-free planar envs behind the same engine — DartCartPole-v1,
DartCartPoleSwingUp-v1, DartDoubleInvertedPendulumEnv-v1 (reference gym/envs/dart/cart_pole.py,
cartpole_swingup.py, inverted_double_pendulum.py).

Their skeletons run on the topology-generic loop kernel through `do_simulation` (dartb_substep, the
literal `set_forces; world.step()` drop-in); obs / reward / done are a few batched torch ops on the
device, restating the reference step() line by line.  Reset noise uses torch's generator (the
reference uses np_random; parity tests set states explicitly)."""
from __future__ import annotations

import math

import numpy as np
import torch

from .dart_env import DartEnv
from .kinematics import body_point_world


class _HostTaskEnv(DartEnv):
    """Shared plumbing: batched state on the device, reference return types for num_envs == 1."""

    def _state(self):
        return self.engine.get_state(torch.float64)

    def _finish(self, ob, reward, done):
        if self.batched and self.auto_reset and bool(done.any()):
            self._reset_worlds(done)
            ob = torch.where(done[:, None], self._get_obs(), ob)
        if self.batched and self.output == "torch":
            return ob, reward, done, {}
        ob, reward, done = ob.cpu().numpy(), reward.cpu().numpy(), done.cpu().numpy()
        if self.batched:
            return ob, reward, done, {}
        return ob[0], float(reward[0]), bool(done[0]), {}

    def _action(self, a):
        a = torch.as_tensor(np.asarray(a, dtype=np.float64) if not isinstance(a, torch.Tensor) else a,
                            device=self.engine.device).to(torch.float64).reshape(self.num_envs, self.act_dim)
        return a

    def _reset_worlds(self, mask):
        q, dq = self._state()
        qn, dqn = self._sample_reset(int(self.num_envs))
        q = torch.where(mask[:, None], qn, q)
        dq = torch.where(mask[:, None], dqn, dq)
        self.engine.set_state(q.contiguous(), dq.contiguous())

    def reset_model(self):
        self.engine.reset()  # world.reset()
        q, dq = self._sample_reset(self.num_envs)
        self.engine.set_state(q.contiguous(), dq.contiguous())
        return self._get_obs()

    def _gen(self):
        if getattr(self, "_tgen", None) is None or self._tgen_seed != self._seed_value:
            self._tgen = torch.Generator(device=self.engine.device)
            self._tgen.manual_seed(self._seed_value & 0x7FFFFFFFFFFFFFFF)
            self._tgen_seed = self._seed_value
        return self._tgen

    def _uniform(self, lo, hi, shape):
        return torch.rand(shape, generator=self._gen(), device=self.engine.device, dtype=torch.float64) * (hi - lo) + lo

    def _q0(self):
        dev = self.engine.device
        return (torch.tensor(self.model.q_init(), device=dev).expand(self.num_envs, -1),
                torch.tensor(self.model.dq_init(), device=dev).expand(self.num_envs, -1))


class DartCartPoleEnv(_HostTaskEnv):
    """cart_pole.py:5-41."""

    def __init__(self, **kw):
        control_bounds = np.array([[1.0], [-1.0]])
        self.action_scale = 100
        DartEnv.__init__(self, "cartpole.skel", 2, 4, control_bounds, dt=0.02, task=None, **kw)

    def step(self, a):
        a = self._action(a)
        tau = torch.zeros((self.num_envs, self.model.n_dofs), dtype=torch.float64, device=self.engine.device)
        tau[:, 0] = a[:, 0] * self.action_scale
        self.do_simulation(tau, self.frame_skip)
        ob = self._get_obs()
        reward = torch.ones(self.num_envs, dtype=torch.float64, device=ob.device)
        notdone = torch.isfinite(ob).all(1) & (ob[:, 1].abs() <= .2)
        return self._finish(ob, reward, ~notdone)

    def _get_obs(self):
        q, dq = self._state()
        return torch.cat([q, dq], 1)

    def _sample_reset(self, n):
        q0, dq0 = self._q0()
        return q0 + self._uniform(-.01, .01, q0.shape), dq0 + self._uniform(-.01, .01, q0.shape)


class DartCartPoleSwingUpEnv(_HostTaskEnv):
    """cartpole_swingup.py:7-56."""

    def __init__(self, **kw):
        self.control_bounds = np.array([[1.0], [-1.0]])
        self.action_scale = 40
        DartEnv.__init__(self, "cartpole_swingup.skel", 2, 4, self.control_bounds, dt=0.01, task=None, **kw)

    def step(self, a):
        a = self._action(a)
        tau = torch.zeros((self.num_envs, self.model.n_dofs), dtype=torch.float64, device=self.engine.device)
        tau[:, 0] = a[:, 0] * self.action_scale
        self.do_simulation(tau, self.frame_skip)
        q, dq = self._state()
        ob = torch.cat([q, dq], 1)
        ang = q[:, 1]
        reward = 6.0 - 1.0 * ang.abs() - 0.01 * (a ** 2).sum(1) - 0.01 * q[:, 0].abs()
        done = (ang.abs() > 8 * math.pi) | (dq[:, 1].abs() > 25) | (q[:, 0].abs() > 5)
        return self._finish(ob, reward, done)

    def _get_obs(self):
        q, dq = self._state()
        return torch.cat([q, dq], 1)

    def _sample_reset(self, n):
        q0, dq0 = self._q0()
        q = q0 + self._uniform(-.1, .1, q0.shape)
        dq = dq0 + self._uniform(-.01, .01, q0.shape)
        flip = self._uniform(0, 1, (q0.shape[0],)) > 0.5
        q[:, 1] += torch.where(flip, math.pi, -math.pi)
        return q, dq


class DartDoubleInvertedPendulumEnv(_HostTaskEnv):
    """inverted_double_pendulum.py:8-71."""

    def __init__(self, **kw):
        control_bounds = np.array([[1.0], [-1.0]])
        self.action_scale = 40
        DartEnv.__init__(self, "inverted_double_pendulum.skel", 2, 8, control_bounds, dt=0.01, task=None, **kw)
        names = [b.name for b in self.model.bodies]
        self._cart, self._weight = names.index("cart"), names.index("weight")

    def step(self, a):
        a = self._action(a)
        tau = torch.zeros((self.num_envs, self.model.n_dofs), dtype=torch.float64, device=self.engine.device)
        tau[:, 0] = a[:, 0] * self.action_scale
        self.do_simulation(tau, self.frame_skip)
        q, dq = self._state()
        ob = self._obs_from(q, dq)
        base = body_point_world(self.model, q, self._cart)[:, 1]
        raw_height = body_point_world(self.model, q, self._weight)[:, 1]
        height = 2.0 * (raw_height - base - 0.02) / 0.6
        v1, v2 = dq[:, 1], dq[:, 2]
        dist_penalty = 0.01 * ob[:, 0] ** 2 + (height - 2.) ** 2
        vel_penalty = 1e-3 * v1 ** 2 + 5e-3 * v2 ** 2
        reward = 10. - dist_penalty - vel_penalty
        return self._finish(ob, reward, height <= 1)

    def _obs_from(self, q, dq):
        return torch.cat([q[:, :1], torch.sin(q[:, 1:]), torch.cos(q[:, 1:]), dq], 1)

    def _get_obs(self):
        return self._obs_from(*self._state())

    def _sample_reset(self, n):
        q0, dq0 = self._q0()
        q = q0 + self._uniform(-.1, .1, q0.shape)
        dq = dq0 + torch.randn(q0.shape, generator=self._gen(), device=self.engine.device, dtype=torch.float64) * 0.1
        return q, dq


class DartReacher2dEnv(_HostTaskEnv):
    """reacher2d.py:5-68 (registered as DartReacher-v1).  Joint Coulomb friction (reacher2d.skel
    friction 0.05) enters the LCP as +-friction*dt rows (DART JointCoulombFrictionConstraint)."""

    def __init__(self, **kw):
        self.action_scale = np.array([200.0, 200.0])
        self.control_bounds = np.array([[1.0, 1.0], [-1.0, -1.0]])
        DartEnv.__init__(self, "reacher2d.skel", 2, 11, self.control_bounds, dt=0.01, task=None, collidable=False, **kw)
        dev = self.engine.device
        self.target = torch.tensor([0.1, 0.01, -0.1], dtype=torch.float64, device=dev).repeat(self.num_envs, 1)
        self._tip = self.model.n_bodies - 1
        self._tip_com = tuple(self.model.bodies[self._tip].com)

    def _tip_vec(self, q):
        return body_point_world(self.model, q, self._tip, self._tip_com) - self.target

    def step(self, a):
        a = self._action(a)
        lo = torch.tensor(self.control_bounds[1], device=a.device)
        hi = torch.tensor(self.control_bounds[0], device=a.device)
        tau = torch.minimum(torch.maximum(a, lo), hi) * torch.tensor(self.action_scale, device=a.device)
        self.do_simulation(tau, self.frame_skip)
        q, dq = self._state()
        ob = self._obs_from(q, dq)
        reward = -self._tip_vec(q).norm(dim=1) - (a ** 2).sum(1)
        return self._finish(ob, reward, torch.zeros(self.num_envs, dtype=torch.bool, device=ob.device))

    def _obs_from(self, q, dq):
        return torch.cat([torch.cos(q), torch.sin(q), self.target[:, [0, 2]], dq, self._tip_vec(q)], 1)

    def _get_obs(self):
        return self._obs_from(*self._state())

    def _sample_reset(self, n):
        q0, dq0 = self._q0()
        q = q0 + self._uniform(-.01, .01, q0.shape)
        dq = dq0 + self._uniform(-.005, .005, q0.shape)
        # rejection-sample the target inside the 0.2 disc of the x-z plane (reacher2d.py:56-60)
        t = self._uniform(-.2, .2, (n, 3))
        t[:, 1] = 0.0
        for _ in range(64):
            bad = t.norm(dim=1) >= .2
            if not bool(bad.any()):
                break
            t2 = self._uniform(-.2, .2, (n, 3))
            t2[:, 1] = 0.0
            t = torch.where(bad[:, None], t2, t)
        t[:, 1] = 0.01
        self.target = t
        return q, dq

    def _reset_worlds(self, mask):
        old = self.target.clone()
        super()._reset_worlds(mask)
        self.target = torch.where(mask[:, None], self.target, old)


CONTACT_FREE = {"DartReacher-v1": (DartReacher2dEnv, 50),
                "DartCartPole-v1": (DartCartPoleEnv, 1000), "DartCartPoleSwingUp-v1": (DartCartPoleSwingUpEnv, 500),
                "DartDoubleInvertedPendulumEnv-v1": (DartDoubleInvertedPendulumEnv, 1000)}
