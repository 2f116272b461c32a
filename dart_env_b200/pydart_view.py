"""pydart2-shaped read/write views over the batched engine: `env.dart_world`, `env.robot_skeleton`,
`skel.bodynodes[i]` — the object surface the reference's env classes and downstream code touch
(`world.dt / step / reset / skeletons / collision_result.contacts`, `skel.ndofs / q / dq /
set_positions / set_velocities / set_forces / q_lower / q_upper / bodynodes / joints`, `bodynode.com /
to_world / local_com / com_spatial_velocity / mass / name`).

Everything is BATCHED: arrays carry a leading [N] axis when the env is batched and have the reference's
shapes when N = 1 and `batched=False`.  State lives on the GPU (the engine's SoA arrays); the views read
it through `dartb_get_state` and do their kinematics in torch (`kinematics.py`), so they are host-side
conveniences, not the hot path (which is `env.step`)."""
from __future__ import annotations

from typing import List

import numpy as np
import torch

from . import kinematics


class Contact:
    """one entry of `world.collision_result.contacts` (walker2d.py:38-41 reads `.force`)"""

    def __init__(self, world: int, bodynode_id: int, point, normal, depth: float, force):
        self.world, self.bodynode_id2 = world, bodynode_id
        self.point, self.normal, self.penetration_depth, self.force = point, normal, depth, force

    def __repr__(self):
        return "Contact(world=%d, body=%d, depth=%.3g, force=%s)" % (self.world, self.bodynode_id2, self.penetration_depth, self.force)


class CollisionResult:
    def __init__(self, env):
        self._env = env

    @property
    def contacts(self) -> List[Contact]:
        cnt, body, data = (t.cpu().numpy() for t in self._env.engine.contacts())
        out = []
        for w in range(len(cnt)):
            for c in range(int(cnt[w])):
                d = data[w, c].astype(np.float64)
                out.append(Contact(w, int(body[w, c]), d[0:3], d[3:6], float(d[6]), d[7:10]))
        return out


class BodyNodeView:
    def __init__(self, skel: "SkeletonView", index: int):
        self.skel, self.id = skel, index
        self._b = skel.model.bodies[index]
        self.name = self._b.name

    def _out(self, t: torch.Tensor):
        a = t.cpu().numpy()
        return a if self.skel.env.batched else a[0]

    def mass(self):
        """bn.mass() (snake_7link.py:24): a float, or the per-world array once set_mass gave the worlds different masses"""
        pw = self.skel.env._body_mass
        if pw is None:
            return float(self._b.mass)
        col = pw[:, self.id]
        return float(col[0]) if (col == col[0]).all() else col.copy()

    m = property(mass)

    def set_mass(self, mass):
        """bn.set_mass(m) (snake_7link.py:117): a scalar for every world of the batch, or one value per world"""
        env = self.skel.env
        M = env._body_param_array("mass")
        M[:, self.id] = mass
        env.set_body_params(M, env._body_mu)

    def set_friction_coeff(self, mu):
        """bn.set_friction_coeff(mu) (snake_7link.py:30,119): a scalar for every world, or one value per world"""
        env = self.skel.env
        F = env._body_param_array("friction")
        F[:, self.id] = mu
        env.set_body_params(env._body_mass, F)

    def local_com(self) -> np.ndarray:
        return np.array(self._b.com, dtype=np.float64)

    def com(self):
        return self._out(kinematics.body_com_world(self.skel.model, self.skel._q())[:, self.id])

    C = property(com)

    def to_world(self, p=(0.0, 0.0, 0.0)):
        return self._out(kinematics.body_point_world(self.skel.model, self.skel._q(), self.id, p))

    def transform(self):
        R, p = kinematics.body_transforms(self.skel.model, self.skel._q())
        n = R.shape[0]
        T = torch.eye(4, dtype=R.dtype, device=R.device).repeat(n, 1, 1)
        T[:, :3, :3], T[:, :3, 3] = R[:, self.id], p[:, self.id]
        return self._out(T)

    T = property(transform)

    def com_spatial_velocity(self):
        q, dq = self.skel._state()
        return self._out(kinematics.body_com_spatial_velocities(self.skel.model, q, dq)[:, self.id])

    def com_linear_velocity(self):
        q, dq = self.skel._state()
        return self._out(kinematics.body_com_spatial_velocities(self.skel.model, q, dq)[:, self.id, 3:])

    dC = property(com_linear_velocity)

    def friction_coeff(self):
        pw = self.skel.env._body_mu
        if pw is None:
            return float(self._b.friction_coeff)
        col = pw[:, self.id]
        return float(col[0]) if (col == col[0]).all() else col.copy()


class DofView:
    def __init__(self, body):
        self.name = body.joint_name
        self.position_lower_limit, self.position_upper_limit = body.q_lo, body.q_hi


class JointView:
    def __init__(self, body):
        self._b = body
        self.name = body.joint_name
        self.dofs = [DofView(body)] if body.dof >= 0 else []

    def has_position_limit(self, _index: int = 0) -> bool:
        return bool(self._b.has_limit)

    def is_position_limit_enforced(self) -> bool:
        return bool(self._b.limit_enforced)


class SkeletonView:
    """`env.robot_skeleton` (= `world.skeletons[-1]`, dart_env.py:62)."""

    def __init__(self, env):
        self.env, self.model = env, env.model
        self.name = env.model.name
        self.bodynodes = [BodyNodeView(self, i) for i in range(env.model.n_bodies)]
        self.joints = [JointView(b) for b in env.model.bodies]
        self.name_to_body = {b.name: b for b in self.bodynodes}

    # --- state
    def _state(self):
        return self.env.engine.get_state(torch.float64)

    def _q(self):
        return self._state()[0]

    def _out(self, t: torch.Tensor):
        a = t.cpu().numpy()
        return a if self.env.batched else a[0]

    @property
    def ndofs(self) -> int:
        return self.model.n_dofs

    num_dofs = ndofs

    @property
    def q(self):
        return self._out(self._state()[0])

    @property
    def dq(self):
        return self._out(self._state()[1])

    def positions(self):
        return self.q

    def velocities(self):
        return self.dq

    def _in(self, v):
        return torch.as_tensor(np.asarray(v, dtype=np.float64).reshape(self.env.num_envs, -1), device=self.env.engine.device).contiguous()

    def set_positions(self, q):
        self.env.engine.set_state(self._in(q), None)

    def set_velocities(self, dq):
        self.env.engine.set_state(None, self._in(dq))

    def set_forces(self, tau):
        """pydart2 semantics: the generalized forces applied by the NEXT world.step() (DART clears them after it)."""
        self.env._pending_tau = np.asarray(tau, dtype=np.float64).reshape(self.env.num_envs, -1)

    @property
    def q_lower(self) -> np.ndarray:
        return self.model.q_lower()

    @property
    def q_upper(self) -> np.ndarray:
        return self.model.q_upper()

    def bodynode(self, name: str) -> BodyNodeView:
        return self.name_to_body[name]

    def com(self):
        """skeleton COM (mass-weighted)"""
        c = kinematics.body_com_world(self.model, self._q())
        m = torch.tensor([b.mass for b in self.model.bodies], dtype=c.dtype, device=c.device)
        return self._out((c * m[None, :, None]).sum(1) / m.sum())

    C = property(com)


class WorldView:
    """`env.dart_world` (dart_world.py:5-22): dt, step(), reset(), skeletons, collision_result."""

    def __init__(self, env):
        self.env = env
        self.robot = SkeletonView(env)
        self.collision_result = CollisionResult(env)

    @property
    def dt(self) -> float:
        return self.env.model.dt

    @property
    def skeletons(self):
        # the reference worlds hold the ground skeleton(s) first and the robot last (dart_env.py:62)
        return [None] * self.env.model.n_static_skeletons + [self.robot]

    def step(self):
        """one DART time step with the forces of the last `set_forces` (then cleared, like DART does)."""
        tau = getattr(self.env, "_pending_tau", None)
        if tau is None:
            tau = np.zeros((self.env.num_envs, self.env.model.n_dofs))
        self.env.do_simulation(tau, 1)
        self.env._pending_tau = None

    def reset(self):
        q0 = np.tile(self.env.model.q_init(), (self.env.num_envs, 1))
        v0 = np.tile(self.env.model.dq_init(), (self.env.num_envs, 1))
        self.env.set_state(q0, v0)
        self.env._pending_tau = None
