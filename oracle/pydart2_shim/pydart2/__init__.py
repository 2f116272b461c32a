"""pydart2 shim — TEST INFRASTRUCTURE.

Presents the slice of the pydart2 Python API that the reference's gym/envs/dart/*.py call
(SURVEY.md §8b "exact call surface") on top of the fp64 CPU oracle (oracle/dart_oracle.c), so
that the UNMODIFIED reference env classes can be imported from /root/reference and run in this
container to mint golden vectors for the task layer (tests/golden/make_golden.py).
Never imported by the product.
"""
import numpy as np

from dart_env_b200.skel import parse_skel
from oracle.oracle import OracleWorld


def init(verbose=True):
    pass


class SkelVector(np.ndarray):
    """ndarray whose tuple keys select several entries: q[0, 2] -> (q[0], q[2])."""

    def __new__(cls, data):
        return np.asarray(data, dtype=np.float64).view(cls)

    def __getitem__(self, key):
        if isinstance(key, tuple):
            return np.array([np.ndarray.__getitem__(self, k) for k in key])
        return np.ndarray.__getitem__(self, key)


class _Contact(object):
    def __init__(self, c, world):
        self.force = np.array(c["force"])
        self.point = np.array(c["point"])
        self.normal = np.array(c["normal"])
        self.penetration_depth = c["depth"]
        self.skel_id1, self.bodynode_id1 = len(world.skeletons) - 1, c["body"]
        self.skel_id2, self.bodynode_id2 = 0, 0


class _CollisionResult(object):
    def __init__(self, world):
        self._world = world

    @property
    def contacts(self):
        return [_Contact(c, self._world) for c in self._world._ow.contacts()]


class _Dof(object):
    def __init__(self, name):
        self.name = name


class _Joint(object):
    def __init__(self, skel, body_index):
        self._skel, self._bi = skel, body_index
        b = skel._model.bodies[body_index]
        self.name = b.joint_name
        self.dofs = [_Dof(b.joint_name)] if b.dof >= 0 else []

    def has_position_limit(self, d):
        return bool(self._skel._model.bodies[self._bi].has_limit)

    def set_position_limit_enforced(self, flag=True):
        self._skel._model.bodies[self._bi].limit_enforced = bool(flag)
        self._skel._world._rebuild()


class _BodyNode(object):
    def __init__(self, skel, i):
        self._skel, self.id = skel, i
        self.name = skel._model.bodies[i].name

    @property
    def _ow(self):
        return self._skel._world._ow

    def com(self):
        return self._ow.body_com(self.id)

    C = property(com)

    def local_com(self):
        return np.array(self._skel._model.bodies[self.id].com)

    def mass(self):
        return self._skel._model.bodies[self.id].mass

    def friction_coeff(self):
        return self._skel._model.bodies[self.id].friction_coeff

    def set_friction_coeff(self, mu):
        self._skel._model.bodies[self.id].friction_coeff = float(mu)
        self._ow.set_friction(self.id, mu)

    def transform(self):
        return self._ow.body_transform(self.id)

    T = property(transform)

    def to_world(self, p=(0, 0, 0)):
        T = self._ow.body_transform(self.id)
        return T[:3, :3] @ np.asarray(p, dtype=np.float64) + T[:3, 3]

    def com_spatial_velocity(self):
        return self._ow.body_com_spatial_velocity(self.id)

    def set_collidable(self, flag):
        if not flag:
            m = self._skel._model
            m.shapes = [s for s in m.shapes if s.body != self.id]
            self._skel._world._rebuild()

    def add_ext_force(self, _force, _offset=None, _isForceLocal=False, _isOffsetLocal=True):
        assert _offset is None and not _isForceLocal, "shim supports the call form the reference uses"
        self._ow.add_ext_force(self.id, np.asarray(_force, dtype=np.float64))


class _StaticBody(object):
    """a body of an immobile skeleton: only set_collidable(False) is meaningful"""

    def __init__(self, world):
        self._world = world

    def set_collidable(self, flag):
        if not flag:
            self._world._model.ground = []
            self._world._rebuild()


class _Skeleton(object):
    def __init__(self, world, model, mobile=True):
        self._world, self._model, self.is_mobile = world, model, mobile
        self.name = model.name if model is not None else "ground skeleton"
        if model is None:
            self.joints, self.ndofs = [], 0
            self.bodynodes = [_StaticBody(world)]
            return
        self.bodynodes = [_BodyNode(self, i) for i in range(model.n_bodies)]
        self.name_to_body = {b.name: b for b in self.bodynodes}
        self.joints = [_Joint(self, i) for i in range(model.n_bodies)]
        self.ndofs = model.n_dofs

    def set_self_collision_check(self, flag):
        pass  # self-collision is never on for the in-scope skeletons

    @property
    def q(self):
        if self._model is None:
            return SkelVector(np.zeros(6))
        return SkelVector(self._world._ow.get_state()[0])

    @q.setter
    def q(self, v):
        if self._model is None:
            return  # moving an immobile, non-collidable marker (reacher2d target) has no dynamic effect
        self.set_positions(v)

    @property
    def dq(self):
        return SkelVector(self._world._ow.get_state()[1])

    @dq.setter
    def dq(self, v):
        self.set_velocities(v)

    def set_positions(self, q):
        self._world._ow.set_state(q=np.asarray(q, dtype=np.float64))

    def set_velocities(self, dq):
        self._world._ow.set_state(dq=np.asarray(dq, dtype=np.float64))

    def set_forces(self, tau):
        self._world._ow.set_forces(np.asarray(tau, dtype=np.float64))

    @property
    def q_lower(self):
        return SkelVector(self._model.q_lower())

    @property
    def q_upper(self):
        return SkelVector(self._model.q_upper())

    def com(self):
        m = np.array([b.mass for b in self._model.bodies])
        c = np.array([self._world._ow.body_com(i) for i in range(self._model.n_bodies)])
        return (m[:, None] * c).sum(0) / m.sum()


class World(object):
    def __init__(self, step, skel_path=None):
        if skel_path is None:
            raise NotImplementedError("shim only loads .skel worlds")
        self._model = parse_skel(skel_path, step)
        self._ow = OracleWorld(self._model)
        self.skeletons = [_Skeleton(self, None, mobile=False) for _ in range(self._model.n_static_skeletons)]
        self.skeletons.append(_Skeleton(self, self._model))
        self.collision_result = _CollisionResult(self)
        self.frame = 0

    def _rebuild(self):
        q, dq = self._ow.get_state()
        self._ow = OracleWorld(self._model)
        self._ow.set_state(q, dq)

    @property
    def dt(self):
        return self._model.dt

    def set_collision_detector(self, detector_id):
        if detector_id != 3:
            raise ValueError("shim implements the ODE detector (id 3) only")

    def step(self):
        self._ow.step()
        self.frame += 1

    def reset(self):
        self._ow.reset()
        self.frame = 0
