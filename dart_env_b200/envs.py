"""The four in-scope task envs (reference gym/envs/dart/{hopper,walker2d,half_cheetah,
snake_7link}.py) and a gym.make-style registry (gym/envs/__init__.py:202-294)."""
from __future__ import annotations

import numpy as np

from .dart_env import DartEnv
from .tasks import SPECS


def _make_cls(env_id: str, clsname: str):
    spec = SPECS[env_id]
    n_act = spec.task.n_act

    class _Env(DartEnv):
        __doc__ = "%s (%s), batched on the GPU." % (env_id, spec.skel)
        spec_id = env_id

        def __init__(self, **kw):
            self.control_bounds = np.array([[1.0] * n_act, [-1.0] * n_act])
            self.action_scale = np.array(spec.task.act_scale)
            kw.setdefault("friction_all", spec.friction_all)
            DartEnv.__init__(self, spec.skel, spec.task.frame_skip, spec.task.n_obs, self.control_bounds,
                             dt=spec.dt, disableViewer=True, task=spec.task, **kw)

    _Env.__name__ = _Env.__qualname__ = clsname
    return _Env


DartHopperEnv = _make_cls("DartHopper-v1", "DartHopperEnv")
DartWalker2dEnv = _make_cls("DartWalker2d-v1", "DartWalker2dEnv")
DartHalfCheetahEnv = _make_cls("DartHalfCheetah-v1", "DartHalfCheetahEnv")


class DartSnake7LinkEnv(_make_cls("DartSnake7Link-v1", "_DartSnake7LinkBase")):
    """snake_7link.py.  `randomize_dynamics` (snake_7link.py:11, hard-coded False there) switches on the per-reset
    dynamics randomisation of snake_7link.py:20-25,115-120: at every `reset_model` of a world — `reset()` and the
    auto-reset inside `step()` — each bodynode's mass becomes original + U(-1.5, 1.5) and its friction coefficient original
    + U(-0.5, 0.5).  The draws happen inside the kernel from the engine's seeded reset generator (the reference draws from
    the unseeded global `np.random`); values are clipped at 0 where the reference would hand DART a negative mass (its two
    massless root bodynodes).  The friction base is the coefficient the env runs with (0: snake_7link.py:29-31)."""

    randomize_dynamics = False

    def __init__(self, randomize_dynamics: bool = False, **kw):
        super().__init__(**kw)
        self.randomize_dynamics = bool(randomize_dynamics)
        # snake_7link.py:20-25: captured BEFORE the constructor's set_friction_coeff(0) loop (snake_7link.py:29-31)
        self.bodynode_original_masses = [b.mass for b in self.model.bodies]
        self.bodynode_original_frictions = list(self._skel_frictions)
        if self.randomize_dynamics:
            self.engine.set_randomize(1.5, 0.5)


DartSnake7LinkEnv.__name__ = DartSnake7LinkEnv.__qualname__ = "DartSnake7LinkEnv"

REGISTRY = {"DartHopper-v1": DartHopperEnv, "DartWalker2d-v1": DartWalker2dEnv,
            "DartHalfCheetah-v1": DartHalfCheetahEnv, "DartSnake7Link-v1": DartSnake7LinkEnv}


def make(env_id: str, **kw) -> DartEnv:
    """gym.make(id): the env wrapped in TimeLimit(max_episode_steps) (registration.py:94-96);
    here the time limit runs inside the kernel (per-world elapsed counter)."""
    from .envs_contact_free import CONTACT_FREE
    if env_id in CONTACT_FREE:  # (gym/envs/__init__.py:219-264: the registered time limits)
        kw.setdefault("max_episode_steps", CONTACT_FREE[env_id][1])
        return CONTACT_FREE[env_id][0](**kw)
    if env_id not in REGISTRY:
        raise KeyError("No registered env with id: %s (in scope: %s)" % (env_id, sorted(REGISTRY) + sorted(CONTACT_FREE)))
    kw.setdefault("max_episode_steps", SPECS[env_id].max_episode_steps)
    return REGISTRY[env_id](**kw)
