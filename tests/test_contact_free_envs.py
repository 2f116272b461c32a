"""SURVEY.md §8f.1 — DartCartPole / DartCartPoleSwingUp / DartDoubleInvertedPendulum.

Goldens: the reference's own classes (cart_pole.py, cartpole_swingup.py,
inverted_double_pendulum.py) run on the oracle through the pydart2 shim (make_golden.py).
CPU: the topology-generic loop kernel source (host emulation) reproduces the frame_skip DART steps.
GPU: the batched env classes reproduce obs / reward / done."""
import os

import numpy as np
import pytest

from dart_env_b200.cstructs import Task
from dart_env_b200.skel import load_model
from tools.host_emu import emu

GOLD = os.path.join(os.path.dirname(__file__), "golden")
CASES = {"DartCartPole-v1": ("cartpole.npz", "cartpole.skel", 0.02, 100.0),
         "DartCartPoleSwingUp-v1": ("cartpole_swingup.npz", "cartpole_swingup.skel", 0.01, 40.0),
         "DartDoubleInvertedPendulumEnv-v1": ("double_pendulum.npz", "inverted_double_pendulum.skel", 0.01, 40.0),
         "DartReacher-v1": ("reacher2d.npz", "reacher2d.skel", 0.01, 200.0)}


def _model(skel, dt):
    m = load_model(skel, dt)
    m.enforce_limits()
    if skel == "reacher2d.skel":  # reacher2d.py:11-14 set_collidable(False) on every body
        m.shapes, m.ground = [], []
    return m


def _tau(env_id, g, scale):
    a = g["step_action"]
    tau = np.zeros_like(g["step_q"])
    if env_id == "DartReacher-v1":
        tau[:] = np.clip(a, -1, 1) * scale
    else:
        tau[:, 0] = a[:, 0] * scale
    return tau


@pytest.mark.parametrize("env_id", list(CASES))
def test_loop_kernel_source_steps_contact_free_models(env_id):
    from dart_env_b200 import capi
    f, skel, dt, scale = CASES[env_id]
    g = np.load(os.path.join(GOLD, f))
    m = _model(skel, dt)
    assert "loop:generic" in capi.describe(m, Task.physics_only(2))
    q, dq = g["step_q"].copy(), g["step_dq"].copy()
    tau = _tau(env_id, g, scale)
    for _ in range(2):  # frame_skip
        q, dq, *_ = emu.substep(m, Task.physics_only(2), q, dq, tau, f64=True, variant=1)
    assert np.allclose(q, g["step_q2"], rtol=1e-9, atol=1e-10)
    assert np.allclose(dq, g["step_dq2"], rtol=1e-8, atol=1e-9)
    q, dq = g["step_q"].copy(), g["step_dq"].copy()
    for _ in range(2):
        q, dq, *_ = emu.substep(m, Task.physics_only(2), q, dq, tau, f64=False, variant=1)
    assert np.allclose(dq, g["step_dq2"], rtol=2e-4, atol=2e-4)


@pytest.mark.gpu
@pytest.mark.parametrize("env_id", list(CASES))
def test_batched_env_matches_reference_classes(env_id):
    import torch

    from dart_env_b200.envs import make
    f, skel, dt, scale = CASES[env_id]
    g = np.load(os.path.join(GOLD, f))
    n = len(g["step_q"])
    env = make(env_id, num_envs=n, output="numpy", seed=0, auto_reset=False, f64=True)
    assert "loop:generic" in env.engine.kernel_name
    env.set_state(g["step_q"], g["step_dq"])
    if env_id == "DartReacher-v1":
        env.target = torch.tensor(g["step_target"], device="cuda")
    ob, rew, done, _ = env.step(g["step_action"])
    s = env.state_vector()
    nd = g["step_q"].shape[1]
    assert np.allclose(s[:, :nd], g["step_q2"], rtol=1e-9, atol=1e-10)
    assert np.allclose(ob, g["step_obs"], rtol=1e-8, atol=1e-8)
    assert np.allclose(rew, g["step_reward"], rtol=1e-8, atol=1e-8)
    assert np.array_equal(done, g["step_done"].astype(bool))
    env.close()
    # fp32 product path + gym surface for one env
    env = make(env_id, seed=1)
    o = env.reset()
    assert env.observation_space.contains(o)
    o, r, d, info = env.step(env.action_space.sample())
    assert env.observation_space.contains(o) and np.isscalar(r) and isinstance(d, bool)
    env.close()
    # auto-reset in batched mode
    env = make(env_id, num_envs=64, output="torch", seed=2)
    env.reset()
    for _ in range(60):
        o, r, d, _ = env.step(torch.rand((64, env.act_dim), device="cuda") * 2 - 1)
    assert torch.isfinite(o).all()
    env.close()
