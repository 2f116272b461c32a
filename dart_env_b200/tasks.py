"""Task constants of the in-scope Dart envs (reference gym/envs/dart/*.py constructors).

Every number below is hard-coded in the reference constructors / step functions
(SURVEY.md Appendix A); the reference takes no kwargs for them.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, Optional

from .cstructs import OBS_HEIGHT_Q2_DQ, OBS_Q1_DQ, Task


@dataclass
class EnvSpec:
    id: str
    skel: str
    dt: float                      # pydart World time step (dart_env.py:29 default 0.002)
    task: Task
    max_episode_steps: int
    friction_all: Optional[float] = None   # snake_7link.py:29-31 sets mu = 0 on every link
    reward_threshold: Optional[float] = None


SPECS: Dict[str, EnvSpec] = {}


def _reg(spec: EnvSpec) -> None:
    SPECS[spec.id] = spec


# hopper.py:8-12,24-74 ; gym/envs/__init__.py:206-211
_reg(EnvSpec("DartHopper-v1", "hopper_capsule.skel", 0.002, Task(
    frame_skip=4, act_dof=[3, 4, 5], act_scale=[200.0] * 3, n_obs=11, obs_mode=OBS_HEIGHT_Q2_DQ,
    dq_clip=10.0, height_body=2, height_lo=0.7, height_hi=1.8, ang_max=0.2, alive_bonus=1.0,
    ctrl_cost=1e-3, limit_pen_dof=4, limit_pen_margin=0.05, limit_pen_weight=0.5),
    max_episode_steps=1000, reward_threshold=3800.0))

# walker2d.py:8-12,22-74 ; gym/envs/__init__.py:266-270
_reg(EnvSpec("DartWalker2d-v1", "walker2d.skel", 0.002, Task(
    frame_skip=4, act_dof=[3, 4, 5, 6, 7, 8], act_scale=[100.0, 100.0, 20.0, 100.0, 100.0, 20.0], n_obs=17,
    obs_mode=OBS_HEIGHT_Q2_DQ, dq_clip=10.0, height_body=2, height_lo=0.8, height_hi=2.0, ang_max=1.0,
    alive_bonus=1.0, ctrl_cost=1e-3), max_episode_steps=1000))

# half_cheetah.py:7-17,27-85 ; gym/envs/__init__.py:213-218
_reg(EnvSpec("DartHalfCheetah-v1", "half_cheetah.skel", 0.01, Task(
    frame_skip=5, act_dof=[3, 4, 5, 6, 7, 8], act_scale=[120.0, 90.0, 60.0, 120.0, 60.0, 30.0], n_obs=17,
    obs_mode=OBS_Q1_DQ, ang_max=1.3, alive_bonus=1.0, ctrl_cost=1e-1, zero_reward_on_blowup=True),
    max_episode_steps=1000, reward_threshold=4800.0))

# snake_7link.py:8-31,35-99 ; gym/envs/__init__.py:290-294
_reg(EnvSpec("DartSnake7Link-v1", "snake_7link.skel", 0.002, Task(
    frame_skip=4, act_dof=[3, 4, 5, 6, 7, 8], act_scale=[200.0] * 6, n_obs=17, obs_mode=OBS_Q1_DQ,
    ang_max=1.5, alive_bonus=0.1, ctrl_cost=1e-3, dev_cost=0.1, fluid_force=True),
    max_episode_steps=1000, friction_all=0.0))
