"""gym-native surface (only when `gym` — the reference is a fork of gym 0.17 — is importable).

    import dart_env_b200.gym_adapter as ga
    ga.register()                              # the Dart ids now resolve to the B200 engine
    env = gym.make("DartHopper-v1")            # gym.Env subclass inside gym's own TimeLimit wrapper
    venv = gym.vector.make("DartHopper-v1", 4096)   # ONE batched engine behind the gym.vector.VectorEnv API

What a maintainer of the reference would change instead: the `entry_point` strings of
gym/envs/__init__.py:206-218,219-264,266-270,290-294.  `register()` does the same at run time by
replacing the specs of those ids in gym's registry, keeping their max_episode_steps / reward_threshold.

The single env returns the reference's types (float64 obs, Python float reward, bool done, dict info:
gym/envs/tests/test_envs.py:10-37); the vector env returns (obs [N, ...] float32, rewards float64 [N], dones bool [N],
infos LIST of N dicts: gym/vector/sync_vector_env.py:73-84) with auto-reset, and accepts seed(int | list)
(sync_vector_env.py:50-57)."""
from __future__ import annotations

import numpy as np

import gym
from gym.envs.registration import EnvSpec, registry
from gym.vector.vector_env import VectorEnv

from . import envs as _envs
from .envs_contact_free import CONTACT_FREE

IDS = list(_envs.REGISTRY) + list(CONTACT_FREE)


def _gym_box(space):
    return gym.spaces.Box(low=space.low.astype(np.float32), high=space.high.astype(np.float32), dtype=np.float32)


def _base_cls(env_id):
    return _envs.REGISTRY[env_id] if env_id in _envs.REGISTRY else CONTACT_FREE[env_id][0]


_SINGLE = {}


def single_env_class(env_id):
    """gym.Env subclass of one world of `env_id` (what gym.make instantiates; gym's TimeLimit wraps it)."""
    if env_id in _SINGLE:
        return _SINGLE[env_id]
    base = _base_cls(env_id)

    class _GymEnv(base, gym.Env):
        metadata = {"render.modes": []}

        def __init__(self, **kw):
            kw.update(num_envs=1, batched=False, max_episode_steps=0)   # the TimeLimit wrapper of gym.make counts steps
            base.__init__(self, **kw)
            self.action_space = _gym_box(self.single_action_space)
            self.observation_space = _gym_box(self.single_observation_space)
            self.reward_range = (-float("inf"), float("inf"))

        def render(self, mode="human", close=False):
            return base.render(self, mode=mode, close=close)

    _GymEnv.__name__ = _GymEnv.__qualname__ = base.__name__
    _SINGLE[env_id] = _GymEnv
    return _GymEnv


def _entry_point(env_id):
    cls = single_env_class(env_id)
    return lambda **kw: cls(**kw)


class DartVectorEnv(VectorEnv):
    """gym.vector.VectorEnv over ONE batched engine: N worlds, one CUDA launch per step_wait()."""

    def __init__(self, env_id, num_envs, **kw):
        kw.setdefault("output", "numpy")
        self.env = _envs.make(env_id, num_envs=num_envs, batched=True, **kw)
        VectorEnv.__init__(self, num_envs, _gym_box(self.env.single_observation_space), _gym_box(self.env.single_action_space))
        self._actions = None

    def seed(self, seeds=None):
        if seeds is None:
            seeds = [None] * self.num_envs
        if isinstance(seeds, int):
            seeds = [seeds + i for i in range(self.num_envs)]    # sync_vector_env.py:53-54
        assert len(seeds) == self.num_envs
        if any(s is None for s in seeds):
            return self.env.seed(None)
        return self.env.seed(list(seeds))

    def reset_async(self):
        pass

    def reset_wait(self, **kw):
        return self.env.reset()

    def step_async(self, actions):
        self._actions = actions

    def step_wait(self, **kw):
        obs, rew, done, info = self.env.step(np.asarray(self._actions, dtype=np.float32))
        infos = [{} for _ in range(self.num_envs)]               # sync_vector_env.py:73-84: one dict per env
        trunc = info.get("TimeLimit.truncated")
        if trunc is not None:
            for i in np.nonzero(done)[0]:
                infos[i]["TimeLimit.truncated"] = bool(trunc[i])  # gym/wrappers/time_limit.py:18-20
        return obs, rew, done, infos

    def close_extras(self, **kw):
        self.env.close()


# importable names for string entry points (`entry_point='dart_env_b200.gym_adapter:DartHopperEnv'`)
for _id in IDS:
    _cls = single_env_class(_id)
    globals()[_cls.__name__] = _cls
del _id, _cls

_ORIG_VECTOR_MAKE = None


def register(override_vector_make: bool = True):
    """Point the Dart ids of gym's registry at the B200 engine (idempotent)."""
    global _ORIG_VECTOR_MAKE
    for env_id in IDS:
        old = registry.env_specs.get(env_id)
        limit = getattr(old, "max_episode_steps", None) if old is not None else None
        if limit is None:
            limit = _envs.SPECS[env_id].max_episode_steps if env_id in _envs.SPECS else CONTACT_FREE[env_id][1]
        registry.env_specs[env_id] = EnvSpec(env_id, entry_point=_entry_point(env_id), max_episode_steps=limit,
                                             reward_threshold=getattr(old, "reward_threshold", None))
    if override_vector_make and _ORIG_VECTOR_MAKE is None:
        import gym.vector as gv
        _ORIG_VECTOR_MAKE = gv.make

        def make(id, num_envs=1, asynchronous=True, wrappers=None, **kwargs):
            if id in IDS and wrappers is None:
                return DartVectorEnv(id, num_envs, **kwargs)
            return _ORIG_VECTOR_MAKE(id, num_envs=num_envs, asynchronous=asynchronous, wrappers=wrappers, **kwargs)

        gv.make = make
    return IDS
