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


def _env_cls(env_id):
    from dart_env_b200.envs_contact_free import CONTACT_FREE
    return CONTACT_FREE[env_id][0]


@pytest.mark.parametrize("env_id", list(CASES))
def test_fused_task_kinds_match_reference_classes(env_id):
    """The CUDA task layer of the contact-free envs (csrc/task_kinds.cuh: obs / reward / done per dartb_task_t.kind),
    compiled for the CPU by tools/host_emu, on the goldens' post-step states: the numbers the reference's own classes
    returned.  fp64 instantiation: algorithmic equality; fp32: the stated tolerance."""
    f, skel, dt, scale = CASES[env_id]
    g = np.load(os.path.join(GOLD, f))
    m = _model(skel, dt)
    task = _env_cls(env_id)._task(None, m)
    a2 = (g["step_action"] ** 2).sum(1)
    for f64, tol in ((True, 1e-9), (False, 2e-5)):
        obs, rew, done, *_ = emu.task_kind(m, task, g["step_q2"], g["step_dq2"], aux=g["step_target"], a2=a2, f64=f64)
        assert np.allclose(obs, g["step_obs"], rtol=max(tol, 1e-6), atol=max(tol, 1e-6))     # obs are float32 at the boundary
        assert np.allclose(rew, g["step_reward"], rtol=max(tol, 1e-9), atol=max(tol, 1e-9) * 10)
        thr = {"DartCartPole-v1": np.abs(np.abs(g["step_q2"][:, 1]) - 0.2)}.get(env_id)
        safe = np.ones(len(done), dtype=bool) if thr is None else thr > 1e-5
        assert np.array_equal(done[safe], g["step_done"][safe].astype(bool))


@pytest.mark.parametrize("env_id", list(CASES))
def test_fused_reset_draws_follow_the_reference_distributions(env_id):
    """reset_model() inside the kernel: counter-based draws with the reference's ranges (cart_pole.py:31-36,
    cartpole_swingup.py:40-52, inverted_double_pendulum.py:56-63, reacher2d.py:50-64)."""
    f, skel, dt, scale = CASES[env_id]
    m = _model(skel, dt)
    task = _env_cls(env_id)._task(None, m)
    n, nd = 4000, m.n_dofs
    obs, rew, done, q, dq, aux = emu.task_kind(m, task, np.zeros((n, nd)), np.zeros((n, nd)), f64=True, do_reset=True, seed=7)
    q0, dq0 = np.array(m.q_init()), np.array(m.dq_init())
    dqn = dq - dq0
    if env_id == "DartCartPole-v1":
        assert np.abs(q - q0).max() <= 0.01 and np.abs(dqn).max() <= 0.01 and (q - q0).std() > 0.005
    elif env_id == "DartCartPoleSwingUp-v1":
        flip = q[:, 1] - q0[1]
        assert np.abs(np.abs(flip) - np.pi).max() <= 0.1 and 0.4 < (flip > 0).mean() < 0.6
        assert np.abs(q[:, 0] - q0[0]).max() <= 0.1 and np.abs(dqn).max() <= 0.01
    elif env_id == "DartDoubleInvertedPendulumEnv-v1":
        assert np.abs(q - q0).max() <= 0.1
        assert abs(dqn.std() - 0.1) < 0.005 and abs(dqn.mean()) < 0.005 and np.abs(dqn).max() > 0.3   # 0.1 * randn
    else:
        assert np.abs(q - q0).max() <= 0.01 and np.abs(dqn).max() <= 0.005
        assert (aux[:, 1] == 0.01).all() and np.hypot(aux[:, 0], aux[:, 2]).max() < 0.2
        assert np.abs(aux[:, [0, 2]]).max() > 0.15 and abs(aux[:, 0].mean()) < 0.01
    # determinism: the same (seed, world, episode) key gives the same draw
    _, _, _, q_b, dq_b, aux_b = emu.task_kind(m, task, np.zeros((n, nd)), np.zeros((n, nd)), f64=True, do_reset=True, seed=7)
    assert np.array_equal(q, q_b) and np.array_equal(dq, dq_b) and np.array_equal(aux, aux_b)


@pytest.mark.gpu
@pytest.mark.parametrize("env_id", list(CASES))
def test_batched_env_matches_reference_classes(env_id):
    import torch

    from dart_env_b200.envs import make
    f, skel, dt, scale = CASES[env_id]
    g = np.load(os.path.join(GOLD, f))
    n = len(g["step_q"])
    env = make(env_id, num_envs=n, output="numpy", seed=0, auto_reset=False, f64=True)
    assert "loop:generic" in env.engine.kernel_name and env.fused     # ONE launch per step: the fused task layer
    env.set_state(g["step_q"], g["step_dq"])
    if env_id == "DartReacher-v1":
        env.target = torch.tensor(g["step_target"], device="cuda")
    l0 = env.engine.launch_count
    ob, rew, done, _ = env.step(g["step_action"])
    assert env.engine.launch_count - l0 == 1
    s = env.state_vector()
    nd = g["step_q"].shape[1]
    # (actions cross the boundary as float32: tau carries their 6e-8 relative rounding)
    assert np.allclose(s[:, :nd], g["step_q2"], rtol=1e-6, atol=1e-7)
    assert np.allclose(ob, g["step_obs"], rtol=1e-6, atol=1e-6)   # observations cross the boundary as float32
    assert np.allclose(rew, g["step_reward"], rtol=1e-6, atol=1e-6)   # (the control cost sees the float32 action)
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
