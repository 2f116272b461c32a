"""ctypes mirrors of the structs in include/dartb.h and the Model/Task -> struct packing."""
from __future__ import annotations

import ctypes as C
import math
from dataclasses import dataclass, field
from typing import List, Optional, Sequence

import numpy as np

from .skel import Model

MAX_BODIES, MAX_SHAPES, MAX_GROUND, MAX_ACT = 24, 24, 4, 16
PM_MAXB = 12   # csrc/planar_model.h: planar bodies after the weld merge (index base of the per-reset dynamics draws)

OBS_Q1_DQ, OBS_HEIGHT_Q2_DQ = 0, 1
TASK_LOCOMOTION, TASK_CARTPOLE, TASK_SWINGUP, TASK_DOUBLE_PENDULUM, TASK_REACHER2D = 0, 1, 2, 3, 4
OPT_LCP_MODE, OPT_PGS_ITERS, OPT_FRICTION_ALL = 1, 2, 3
LCP_EXACT, LCP_PGS = 0, 1


class CBody(C.Structure):
    _fields_ = [
        ("parent", C.c_int32), ("joint_type", C.c_int32), ("dof", C.c_int32), ("limit_enforced", C.c_int32),
        ("T_parent_joint", C.c_double * 12), ("T_child_joint", C.c_double * 12), ("axis", C.c_double * 3),
        ("q_lo", C.c_double), ("q_hi", C.c_double),
        ("damping", C.c_double), ("coulomb", C.c_double), ("spring_k", C.c_double), ("spring_rest", C.c_double),
        ("q_init", C.c_double), ("dq_init", C.c_double),
        ("mass", C.c_double), ("com", C.c_double * 3), ("inertia", C.c_double * 9),
        ("friction_coeff", C.c_double),
    ]


class CShape(C.Structure):
    _fields_ = [("body", C.c_int32), ("type", C.c_int32), ("size", C.c_double * 3), ("T", C.c_double * 12)]


class CModel(C.Structure):
    _fields_ = [
        ("dt", C.c_double), ("gravity", C.c_double * 3),
        ("n_bodies", C.c_int32), ("n_dofs", C.c_int32), ("n_shapes", C.c_int32), ("n_ground", C.c_int32),
        ("bodies", CBody * MAX_BODIES), ("shapes", CShape * MAX_SHAPES), ("ground", CShape * MAX_GROUND),
    ]


class CTask(C.Structure):
    _fields_ = [
        ("frame_skip", C.c_int32), ("n_act", C.c_int32), ("n_obs", C.c_int32),
        ("act_dof", C.c_int32 * MAX_ACT), ("act_scale", C.c_double * MAX_ACT),
        ("act_lo", C.c_double * MAX_ACT), ("act_hi", C.c_double * MAX_ACT),
        ("obs_mode", C.c_int32), ("dq_clip", C.c_double),
        ("height_body", C.c_int32), ("height_lo", C.c_double), ("height_hi", C.c_double),
        ("ang_max", C.c_double), ("alive_bonus", C.c_double), ("ctrl_cost", C.c_double), ("vel_weight", C.c_double),
        ("limit_pen_dof", C.c_int32), ("limit_pen_margin", C.c_double), ("limit_pen_weight", C.c_double),
        ("dev_cost", C.c_double), ("zero_reward_on_blowup", C.c_int32),
        ("fluid_force", C.c_int32), ("fluid_offset", C.c_double), ("fluid_coef", C.c_double),
        ("reset_noise", C.c_double), ("state_bound", C.c_double),
        ("kind", C.c_int32), ("reset_noise_dq", C.c_double), ("probe_body", C.c_int32 * 2), ("probe_local", (C.c_double * 3) * 2),
    ]


def _t12(T: np.ndarray):
    return (C.c_double * 12)(*np.asarray(T, dtype=np.float64)[:3, :4].reshape(-1))


def pack_model(m: Model) -> CModel:
    if m.n_bodies > MAX_BODIES or len(m.shapes) > MAX_SHAPES or len(m.ground) > MAX_GROUND:
        raise ValueError("model too large for dartb_model_t (bodies %d, shapes %d, ground %d)"
                         % (m.n_bodies, len(m.shapes), len(m.ground)))
    cm = CModel()
    cm.dt = m.dt
    cm.gravity = (C.c_double * 3)(*m.gravity)
    cm.n_bodies, cm.n_dofs, cm.n_shapes, cm.n_ground = m.n_bodies, m.n_dofs, len(m.shapes), len(m.ground)
    for i, b in enumerate(m.bodies):
        cb = cm.bodies[i]
        cb.parent, cb.joint_type, cb.dof, cb.limit_enforced = b.parent, b.joint_type, b.dof, int(b.limit_enforced)
        cb.T_parent_joint = _t12(b.T_parent_joint)
        cb.T_child_joint = _t12(b.T_child_joint)
        cb.axis = (C.c_double * 3)(*b.axis)
        cb.q_lo, cb.q_hi = b.q_lo, b.q_hi
        cb.damping, cb.coulomb, cb.spring_k, cb.spring_rest = b.damping, b.coulomb, b.spring_k, b.spring_rest
        cb.q_init, cb.dq_init = b.q_init, b.dq_init
        cb.mass = b.mass
        cb.com = (C.c_double * 3)(*b.com)
        cb.inertia = (C.c_double * 9)(*np.asarray(b.inertia, dtype=np.float64).reshape(-1))
        cb.friction_coeff = b.friction_coeff
    for arr, src in ((cm.shapes, m.shapes), (cm.ground, m.ground)):
        for i, s in enumerate(src):
            arr[i].body, arr[i].type = s.body, s.type
            arr[i].size = (C.c_double * 3)(*s.size)
            arr[i].T = _t12(s.T)
    return cm


@dataclass
class Task:
    """Parameterised task layer (include/dartb.h: dartb_task_t)."""
    frame_skip: int
    act_dof: Sequence[int]
    act_scale: Sequence[float]
    n_obs: int
    obs_mode: int = OBS_Q1_DQ
    act_lo: Optional[Sequence[float]] = None
    act_hi: Optional[Sequence[float]] = None
    dq_clip: float = 0.0
    height_body: int = -1
    height_lo: float = -math.inf
    height_hi: float = math.inf
    ang_max: float = math.inf
    alive_bonus: float = 0.0
    ctrl_cost: float = 0.0
    vel_weight: float = 1.0
    limit_pen_dof: int = -1
    limit_pen_margin: float = 0.05
    limit_pen_weight: float = 0.0
    dev_cost: float = 0.0
    zero_reward_on_blowup: bool = False
    fluid_force: bool = False
    fluid_offset: float = 0.05
    fluid_coef: float = 50.0
    reset_noise: float = 0.005
    state_bound: float = 100.0
    kind: int = TASK_LOCOMOTION
    reset_noise_dq: float = -1.0
    probe_body: Sequence[int] = (-1, -1)
    probe_local: Sequence[Sequence[float]] = ((0.0, 0.0, 0.0), (0.0, 0.0, 0.0))

    @property
    def n_act(self) -> int:
        return len(self.act_dof)

    @staticmethod
    def physics_only(frame_skip: int = 1) -> "Task":
        """No fused task layer: the handle serves dartb_substep / state access only (the env's
        obs / reward / done are then computed by the host class, like the reference does)."""
        return Task(frame_skip=frame_skip, act_dof=[], act_scale=[], n_obs=0, reset_noise=0.0)


def pack_task(t: Task) -> CTask:
    if t.n_act > MAX_ACT:
        raise ValueError("too many actuators")
    ct = CTask()
    ct.frame_skip, ct.n_act, ct.n_obs = t.frame_skip, t.n_act, t.n_obs
    lo = t.act_lo if t.act_lo is not None else [-1.0] * t.n_act
    hi = t.act_hi if t.act_hi is not None else [1.0] * t.n_act
    for i in range(t.n_act):
        ct.act_dof[i] = int(t.act_dof[i])
        ct.act_scale[i] = float(t.act_scale[i])
        ct.act_lo[i], ct.act_hi[i] = float(lo[i]), float(hi[i])
    ct.obs_mode, ct.dq_clip = t.obs_mode, t.dq_clip
    ct.height_body, ct.height_lo, ct.height_hi = t.height_body, t.height_lo, t.height_hi
    ct.ang_max, ct.alive_bonus, ct.ctrl_cost, ct.vel_weight = t.ang_max, t.alive_bonus, t.ctrl_cost, t.vel_weight
    ct.limit_pen_dof, ct.limit_pen_margin, ct.limit_pen_weight = t.limit_pen_dof, t.limit_pen_margin, t.limit_pen_weight
    ct.dev_cost, ct.zero_reward_on_blowup = t.dev_cost, int(t.zero_reward_on_blowup)
    ct.fluid_force, ct.fluid_offset, ct.fluid_coef = int(t.fluid_force), t.fluid_offset, t.fluid_coef
    ct.reset_noise, ct.state_bound = t.reset_noise, t.state_bound
    ct.kind, ct.reset_noise_dq = int(t.kind), float(t.reset_noise_dq)
    for k in range(2):
        ct.probe_body[k] = int(t.probe_body[k])
        for j in range(3):
            ct.probe_local[k][j] = float(t.probe_local[k][j])
    return ct
