"""The C oracle against the committed golden vectors (tests/golden/*.npz).

The goldens' task layer was executed by the reference's own env classes (make_golden.py), so
this pins oracle/dart_oracle.c's orc_env_step / orc_env_obs restatement to the reference code
(hopper.py:24-74, walker2d.py:22-74, half_cheetah.py:27-85, snake_7link.py:35-99)."""
import os

import numpy as np
import pytest

from dart_env_b200.tasks import SPECS
from oracle import oracle as orc

GOLD = os.path.join(os.path.dirname(__file__), "golden")
FILES = {"DartHopper-v1": "hopper.npz", "DartWalker2d-v1": "walker2d.npz",
         "DartHalfCheetah-v1": "halfcheetah.npz", "DartSnake7Link-v1": "snake7link.npz"}


def load(env_id):
    return np.load(os.path.join(GOLD, FILES[env_id]))


@pytest.mark.parametrize("env_id", list(SPECS))
def test_task_layer_matches_reference_classes(models, env_id):
    g = load(env_id)
    spec = SPECS[env_id]
    env = orc.OracleEnv(models[env_id], spec.task)
    n = len(g["step_q"])
    assert n >= 100
    for i in range(n):
        env.world.set_state(g["step_q"][i], g["step_dq"][i])
        obs, rew, done = env.step(g["step_action"][i])
        q2, dq2 = env.world.get_state()
        assert np.allclose(q2, g["step_q2"][i], rtol=0, atol=1e-12)
        assert np.allclose(dq2, g["step_dq2"][i], rtol=0, atol=1e-10)
        assert np.allclose(obs, g["step_obs"][i], rtol=0, atol=1e-10)
        assert abs(rew - g["step_reward"][i]) < 1e-9
        assert done == bool(g["step_done"][i])
    # reset obs layout (q0+noise state -> obs)
    for rec in g["step_reset_obs"]:
        nd = models[env_id].n_dofs
        env.world.set_state(rec[:nd], rec[nd:2 * nd])
        assert np.allclose(env.obs(), rec[2 * nd:], atol=1e-12)


@pytest.mark.parametrize("env_id", list(SPECS))
def test_substep_golden_replay(models, env_id):
    g = load(env_id)
    m = models[env_id]
    w = orc.OracleWorld(m)
    for i in range(len(g["sub_q"])):
        w.set_state(g["sub_q"][i], g["sub_dq"][i])
        for b in range(m.n_bodies):
            if np.any(g["sub_fext"][i][b] != 0):
                w.add_ext_force(b, g["sub_fext"][i][b])
        w.set_forces(g["sub_tau"][i])
        w.step()
        q2, dq2 = w.get_state()
        assert np.array_equal(q2, g["sub_q2"][i]) and np.array_equal(dq2, g["sub_dq2"][i])
        cs = w.contacts()
        assert len(cs) == g["sub_ncontact"][i]
        assert [c["body"] for c in cs] == list(g["sub_contact_body"][i][:len(cs)])
        assert np.array_equal(w.limit_active(), g["sub_limit_active"][i])


def test_golden_covers_edge_cases():
    g = load("DartWalker2d-v1")
    assert g["sub_ncontact"].max() >= 3          # multi-contact
    assert (g["sub_limit_active"] != 0).any()    # joint-limit rows
    assert (g["step_done"]).any()                # termination
    assert np.abs(g["step_action"]).max() > 1.0  # clamp exercised
    h = load("DartHopper-v1")
    assert (h["sub_tie_margin"] == 0).any()      # exactly-flat foot (ODE tie rule)
    s = load("DartSnake7Link-v1")
    assert s["sub_ncontact"].max() == 0          # never touches the ground


def test_reset_noise_is_counter_based_and_bounded():
    u = np.array([orc.reset_uniform(7, w, e, i) for w in range(3) for e in range(3) for i in range(12)])
    assert u.min() >= -1.0 and u.max() < 1.0 and len(np.unique(u)) == len(u)
    assert orc.reset_uniform(7, 1, 2, 3) == orc.reset_uniform(7, 1, 2, 3)
    spec = SPECS["DartHopper-v1"]


def test_pydart2_replay_script_reports_unavailable_cleanly():
    """oracle/replay_with_pydart2.py is the pin against real pydart2 for whoever has it; here (no pydart2, no
    DART) it must say so and exit 0 — and never mistake the oracle's own pydart2 shim for the real thing."""
    import subprocess
    import sys as _sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, PYTHONPATH=os.path.join(root, "oracle", "pydart2_shim"))
    r = subprocess.run([_sys.executable, os.path.join(root, "oracle", "replay_with_pydart2.py")], capture_output=True, text=True, env=env)
    assert r.returncode == 0 and "unavailable" in r.stdout, r.stdout + r.stderr
