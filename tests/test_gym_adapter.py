"""gym-native surface: gym.make / gym.vector.make on the Dart ids reach the B200 engine (dart_env_b200/gym_adapter.py).

The bodies of the reference's own gym/envs/tests/test_envs.py:10-37 and test_determinism.py:6-54 are run on every
in-scope id.  gym here is the REFERENCE's gym 0.17 fork: importable from baseline/_ref (an offline `pip install
--no-deps --target baseline/_ref` of /root/reference minus its assets; git-ignored, travels to the GPU box) or from
/root/reference where that exists."""
import os
import sys
import warnings

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gym():
    try:
        import gym
        return gym
    except ImportError:
        pass
    for cand in (os.path.join(ROOT, "baseline", "_ref"), "/root/reference"):
        if os.path.isdir(os.path.join(cand, "gym")):
            sys.path.append(cand)
            import gym
            return gym
    pytest.skip("gym (the reference's fork) is not importable on this machine")


@pytest.fixture(scope="module")
def ga():
    gym = _gym()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        from dart_env_b200 import gym_adapter
        gym_adapter.register()
    return gym, gym_adapter


def test_registered_ids_make_gym_envs(ga):
    """gym/envs/tests/test_envs.py:10-37 on the Dart ids"""
    gym, adapter = ga
    for env_id in adapter.IDS:
        env = gym.make(env_id)
        assert isinstance(env.unwrapped, gym.Env) and type(env).__name__ == "TimeLimit"
        assert env.spec.max_episode_steps in (50, 500, 1000)
        ob_space, act_space = env.observation_space, env.action_space
        assert isinstance(ob_space, gym.spaces.Box) and isinstance(act_space, gym.spaces.Box)
        ob = env.reset()
        assert ob_space.contains(ob), "Reset observation: {!r} not in space".format(ob)
        a = act_space.sample()
        observation, reward, done, _info = env.step(a)
        assert ob_space.contains(observation), "Step observation: {!r} not in space".format(observation)
        assert np.isscalar(reward), "{} is not a scalar for {}".format(reward, env)
        assert isinstance(done, bool), "Expected {} to be a boolean".format(done)
        env.close()


def test_registered_ids_are_deterministic(ga):
    """gym/envs/tests/test_determinism.py:6-54 on the Dart ids"""
    gym, adapter = ga
    for env_id in adapter.IDS:
        runs = []
        for _ in range(2):
            env = gym.make(env_id)
            env.seed(0)
            ob0 = env.reset()
            env.action_space.seed(0)
            acts = [env.action_space.sample() for _ in range(4)]
            runs.append((ob0, acts, [env.step(a) for a in acts]))
            env.close()
        (o1, a1, s1), (o2, a2, s2) = runs
        assert all(np.array_equal(x, y) for x, y in zip(a1, a2))
        assert type(o1) == type(o2) and np.array_equal(o1, o2)
        for (ob1, r1, d1, i1), (ob2, r2, d2, i2) in zip(s1, s2):
            assert np.array_equal(ob1, ob2) and r1 == r2 and d1 == d2 and i1 == i2


def test_time_limit_wrapper_truncates_the_reacher(ga):
    gym, _ = ga
    env = gym.make("DartReacher-v1")          # done is always False: only gym's TimeLimit (50) ends the episode
    env.seed(1)
    env.reset()
    for t in range(50):
        ob, r, done, info = env.step(np.zeros(2))
        assert done == (t == 49)
    assert info.get("TimeLimit.truncated") is True
    env.close()


def test_vector_make_is_one_batched_engine_with_reference_types(ga):
    gym, adapter = ga
    n = 64
    venv = gym.vector.make("DartHopper-v1", n)
    assert isinstance(venv, gym.vector.VectorEnv) and isinstance(venv, adapter.DartVectorEnv)
    assert venv.observation_space.shape == (n, 11) and venv.single_action_space.shape == (3,)
    seeds = [100 + 7 * i for i in range(n)]
    venv.seed(seeds)
    obs = venv.reset()
    assert obs.shape == (n, 11) and obs.dtype == np.float32
    # sync_vector_env.py:50-57: env i is seeded seeds[i], i.e. draws what a single env seeded seeds[i] draws
    for i in (0, 3, n - 1):
        env = gym.make("DartHopper-v1")
        env.seed(seeds[i])
        assert np.allclose(env.reset(), obs[i], atol=1e-6)
        env.close()
    l0 = venv.env.engine.launch_count
    rng = np.random.RandomState(0)
    for t in range(30):
        obs, rew, done, infos = venv.step(rng.uniform(-1, 1, (n, 3)))
    assert venv.env.engine.launch_count - l0 == 30              # one launch per batched step
    assert obs.dtype == np.float32 and rew.dtype == np.float64 and rew.shape == (n,) and done.dtype == np.bool_
    assert isinstance(infos, list) and len(infos) == n and all(isinstance(d, dict) for d in infos)
    venv.seed(5)                                                 # int: seeds 5, 6, ... (sync_vector_env.py:53-54)
    venv.close()
