"""GPU tests added in round 2 (run on the B200 box, through the C-ABI): product paths that had no test —
the fp32 contact-free envs against the reference-class goldens, the perturbation branch of do_simulation
(dart_env.py:159-172) against the oracle's add_ext_force, every host-buffer route of dartb_step_host /
dartb_step_host_gym (page-locked zero-copy, pageable staging, DARTB_ZEROCOPY=0 copies), the full-row-count
constraint set (every capsule touching and every limit active: no row may be dropped), the opt-in contact
read-back, dartb_seed, and the pinned output pool of the batched host API."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from dart_env_b200.tasks import SPECS

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
CF = {"DartCartPole-v1": "cartpole.npz", "DartCartPoleSwingUp-v1": "cartpole_swingup.npz",
      "DartDoubleInvertedPendulumEnv-v1": "double_pendulum.npz", "DartReacher-v1": "reacher2d.npz"}
TOL_CF = 2e-4   # fp32 contact-free env.step, max over all samples, relative to (1 + |x|)


def _make(env_id, **kw):
    from dart_env_b200.envs import make
    return make(env_id, **kw)


@pytest.mark.parametrize("env_id", list(CF))
def test_contact_free_envs_fp32_match_reference_classes(env_id):
    """the fp32 PRODUCT path of the four contact-free envs against goldens minted from the reference's own classes"""
    g = np.load(os.path.join(GOLD, CF[env_id]))
    n = len(g["step_q"])
    env = _make(env_id, num_envs=n, output="numpy", seed=0, auto_reset=False)
    assert "/f32" in env.engine.kernel_name
    env.set_state(g["step_q"], g["step_dq"])
    if env_id == "DartReacher-v1":
        env.set_target(g["step_target"])
    ob, rew, done, _ = env.step(g["step_action"])
    s = env.state_vector()
    nd = g["step_q"].shape[1]
    rel = lambda a, b: (np.abs(a - b) / (1 + np.abs(b))).max()
    assert rel(s[:, :nd], g["step_q2"]) < TOL_CF and rel(s[:, nd:], g["step_dq2"]) < TOL_CF
    assert rel(ob, g["step_obs"]) < TOL_CF
    assert rel(rew, g["step_reward"]) < 5 * TOL_CF
    # done flags bit-exact away from the thresholds (|angle| = 0.2 for the cart-pole, height = 1 for the pendulum)
    if env_id == "DartCartPole-v1":
        safe = np.abs(np.abs(g["step_obs"][:, 1]) - 0.2) > 1e-4
    else:
        safe = np.ones(n, dtype=bool)
    assert np.array_equal(done[safe], g["step_done"][safe].astype(bool))
    env.close()


@pytest.mark.parametrize("f64", [False, True])
def test_add_perturbation_matches_oracle_ext_force(f64):
    """dart_env.py:159-172: the drawn force is applied with add_ext_force at the origin of bodynodes[bodyid] on every
    sub-step of the call.  Compared with the oracle's add_ext_force on the same force."""
    from oracle import oracle as orc
    env_id = "DartHopper-v1"
    spec = SPECS[env_id]
    n = 32
    env = _make(env_id, num_envs=n, output="numpy", seed=4, auto_reset=False, f64=f64, kernel_variant=0)
    env.reset()
    env.add_perturbation = True
    env.perturbation_parameters = [1.0, 7.5, 3]   # always draw; magnitude; body id
    q0, dq0 = (x.cpu().numpy() for x in env.engine.get_state(torch.float64))
    rng = np.random.RandomState(0)
    tau = np.zeros((n, env.model.n_dofs))
    tau[:, 3:] = rng.uniform(-50, 50, (n, 3))
    env.do_simulation(tau, env.frame_skip)
    f = env.perturb_force.cpu().numpy().astype(np.float64)
    assert (np.abs(f).sum(1) == 7.5).all() and (f[:, 2] == 0).all()     # +-magnitude along x or y
    q1, dq1 = (x.cpu().numpy() for x in env.engine.get_state(torch.float64))
    worst = 0.0
    for w in range(n):
        ow = orc.OracleWorld(env.model)
        ow.set_state(q0[w], dq0[w])
        for _ in range(env.frame_skip):
            ow.add_ext_force(3, f[w])
            ow.set_forces(tau[w])
            ow.step()
        oq, odq = ow.get_state()
        worst = max(worst, float((np.abs(dq1[w] - odq) / (1 + np.abs(odq))).max()), float(np.abs(q1[w] - oq).max()))
    assert worst < (1e-8 if f64 else 2e-4), worst
    # the force really acted: without it the result differs
    env2 = _make(env_id, num_envs=n, output="numpy", seed=4, auto_reset=False, f64=f64, kernel_variant=0)
    env2.reset()
    env2.do_simulation(tau, env2.frame_skip)
    assert np.abs(env2.state_vector()[:, env.model.n_dofs:] - dq1).max() > 1e-3
    # the fused step() does not draw perturbations: it must refuse rather than ignore the flag
    with pytest.raises(NotImplementedError):
        env.step(np.zeros((n, 3), dtype=np.float32))
    env.close(); env2.close()


def _host_step_outputs(env_id, n, seed, route):
    """one env.step() through a host-buffer route; returns (obs, reward, done) as float64/bool numpy"""
    from dart_env_b200.engine import Engine
    from dart_env_b200.skel import load_model
    spec = SPECS[env_id]
    m = load_model(spec.skel, spec.dt)
    m.enforce_limits()
    eng = Engine(m, spec.task, n, seed=seed)
    eng.reset()
    rng = np.random.RandomState(5)
    act = rng.uniform(-1, 1, (n, spec.task.n_act)).astype(np.float32)
    for _ in range(3):   # a few steps so contacts exist
        if route == "device":
            obs = torch.empty((n, spec.task.n_obs), dtype=torch.float32, device="cuda")
            rew = torch.empty((n,), dtype=torch.float32, device="cuda")
            done = torch.empty((n,), dtype=torch.uint8, device="cuda")
            eng.step(torch.tensor(act, device="cuda"), obs, rew, done, True)
            torch.cuda.synchronize()
            out = obs.cpu().numpy(), rew.cpu().numpy().astype(np.float64), (done.cpu().numpy() & 1).astype(bool)
        elif route in ("pageable", "pinned"):
            mk = (lambda *sh, dt: torch.empty(sh, dtype=dt).pin_memory().numpy()) if route == "pinned" else \
                 (lambda *sh, dt: torch.empty(sh, dtype=dt).numpy())
            a = mk(n, spec.task.n_act, dt=torch.float32); a[:] = act
            obs, rew, done = mk(n, spec.task.n_obs, dt=torch.float32), mk(n, dt=torch.float32), mk(n, dt=torch.uint8)
            eng.step_host(a, obs, rew, done, True)
            out = obs.copy(), rew.astype(np.float64), (done & 1).astype(bool)
        else:   # gym types: float64 rewards, bool dones; pageable or pinned outputs
            pin = route == "gym_pinned"
            mk = (lambda *sh, dt: torch.empty(sh, dtype=dt).pin_memory().numpy()) if pin else (lambda *sh, dt: torch.empty(sh, dtype=dt).numpy())
            obs, rew = mk(n, spec.task.n_obs, dt=torch.float32), mk(n, dt=torch.float64)
            done = mk(n, dt=torch.uint8)
            import ctypes as C
            from dart_env_b200 import capi
            capi.check(eng.L.dartb_step_host_gym(eng.h, C.c_void_p(act.ctypes.data), C.c_void_p(obs.ctypes.data),
                                                 C.c_void_p(rew.ctypes.data), C.c_void_p(done.ctypes.data), None, 1, eng._stream()))
            assert set(np.unique(done)) <= {0, 1}
            out = obs.copy(), rew.copy(), done.astype(bool)
    eng.close()
    return out


def test_every_host_buffer_route_equals_the_device_path():
    """dartb_step_host with page-locked (zero-copy) and pageable (staging) buffers, dartb_step_host_gym with both: the
    same bits as dartb_step on device buffers."""
    ref = _host_step_outputs("DartHopper-v1", 300, 6, "device")
    for route in ("pinned", "pageable", "gym_pinned", "gym_pageable"):
        o, r, d = _host_step_outputs("DartHopper-v1", 300, 6, route)
        assert np.array_equal(o, ref[0]) and np.array_equal(r, ref[1]) and np.array_equal(d, ref[2]), route


def test_zerocopy_off_copy_path_equals_the_device_path():
    """DARTB_ZEROCOPY=0 (explicit H2D / D2H copies) is read once per process: run it in a child."""
    code = ("import sys; sys.path[:0] = [%r, %r]; import numpy as np; from test_gpu_round2 import _host_step_outputs as f\n"
            "ref = f('DartHopper-v1', 300, 6, 'device')\n"
            "for route in ('pinned', 'pageable', 'gym_pageable'):\n"
            "    o, r, d = f('DartHopper-v1', 300, 6, route)\n"
            "    assert np.array_equal(o, ref[0]) and np.array_equal(r, ref[1]) and np.array_equal(d, ref[2]), route\n"
            "print('copy-path ok')\n") % (ROOT, os.path.join(ROOT, "tests"))
    env = dict(os.environ, DARTB_ZEROCOPY="0")
    res = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0 and "copy-path ok" in res.stdout, res.stdout + res.stderr


@pytest.mark.parametrize("env_id", ["DartHopper-v1", "DartWalker2d-v1", "DartHalfCheetah-v1"])
@pytest.mark.parametrize("variant", [0, 2])
def test_full_constraint_set_no_row_dropped(env_id, variant):
    """Adversarial state: the skeleton lies in the ground with EVERY capsule touching and EVERY enforced limit violated,
    i.e. the maximum row count 2*NS + NL of the topology.  The fp64 kernels must reproduce the oracle's step (a silently
    dropped row would change dq by O(1)), and the contact count must be the capsule count."""
    from dart_env_b200.engine import Engine
    from dart_env_b200.skel import load_model
    from oracle import oracle as orc
    spec = SPECS[env_id]
    m = load_model(spec.skel, spec.dt)
    m.enforce_limits()
    nd = m.n_dofs
    rng = np.random.RandomState(3)
    NR = 2 * len(m.shapes) + sum(b.limit_enforced for b in m.bodies)
    cand = []
    ow = orc.OracleWorld(m)
    for trial in range(4000):
        q = np.array(m.q_init(), dtype=float)
        q[2] = rng.choice([-1.57, 1.57]) + rng.uniform(-0.15, 0.15)   # lying down
        q[1] = rng.uniform(-1.5, -0.6)
        for d, bi in enumerate(m.dof_bodies()):
            b = m.bodies[bi]
            if b.limit_enforced:
                if rng.rand() < 0.8:   # violate the limit that keeps the limb straighter
                    lim = b.q_lo if abs(b.q_lo) < abs(b.q_hi) else b.q_hi
                    q[d] = lim + np.sign(lim if lim != 0 else (1 if lim is b.q_hi else -1)) * rng.uniform(0.0, 0.02)
                    if lim == 0:
                        q[d] = rng.uniform(0.0, 0.02) * (1 if b.q_hi == 0 else -1)
                else:
                    q[d] = rng.uniform(b.q_lo, b.q_hi)
        dq = rng.uniform(-1, 1, nd)
        ow.set_state(q, dq); ow.set_forces(np.zeros(nd)); ow.step()
        if ow.lcp_failed():
            continue
        # rows in the planar kernels' terms: (normal, in-plane tangent) per contact + active limits (the oracle also
        # carries the inert out-of-plane tangent row of every contact)
        cand.append((2 * len(ow.contacts()) + int((ow.limit_active() != 0).sum()), trial, q, dq, ow.get_state()[1].copy(), len(ow.contacts())))
    cand.sort(key=lambda c: -c[0])
    S = [(c[2], c[3], c[4], c[0], c[5]) for c in cand[:8]]
    print("%s: largest row counts found %s of NR = %d" % (env_id, [c[0] for c in cand[:8]], NR))
    assert S[0][3] >= NR - (8 if env_id == "DartHalfCheetah-v1" else 0)   # hopper, walker: the FULL set, 2*NS + NL rows
    q, dq, ref = (np.array([s[k] for s in S]) for k in range(3))
    nrows = max(s[3] for s in S)
    eng = Engine(m, spec.task, len(S), f64=True, kernel_variant=variant)
    eng.set_state(torch.tensor(q, device="cuda"), torch.tensor(dq, device="cuda"))
    eng.substep(torch.zeros((len(S), nd), dtype=torch.float64, device="cuda"))
    _, dq2 = eng.get_state(torch.float64)
    cnt, body, data = eng.contacts()
    torch.cuda.synchronize()
    assert np.array_equal(cnt.cpu().numpy(), np.array([s[4] for s in S]))
    err = np.abs(dq2.cpu().numpy() - ref).max()
    assert err < 1e-6, (err, nrows)
    eng.close()


def test_contact_readback_is_opt_in():
    from dart_env_b200 import capi
    env = _make("DartWalker2d-v1", num_envs=64, output="torch", seed=1)
    env.reset()
    env.step(torch.zeros((64, 6), device="cuda"))
    with pytest.raises(capi.DartbError):
        env.contacts()
    env.do_simulation(np.zeros((64, 9)), 1)     # the literal World.step() always refreshes collision_result
    assert env.contacts()[0].shape == (64,)
    env.engine.set_contacts(True)
    env.step(torch.zeros((64, 6), device="cuda"))
    assert env.contacts()[0].shape == (64,)
    env.close()


def test_seed_rekeys_reset_noise_without_touching_state():
    from oracle import oracle as orc
    env_id = "DartHopper-v1"
    env = _make(env_id, num_envs=16, output="torch", seed=10)
    env.reset()
    q0, _ = env.engine.get_state(torch.float64)
    env.seed(77)
    q1, _ = env.engine.get_state(torch.float64)
    assert torch.equal(q0, q1)                   # dart_env.py:117-119: seeding only reseeds
    env.reset()                                  # second episode of every world, drawn under the new key
    q2, _ = env.engine.get_state(torch.float64)
    for w in (0, 5, 15):
        expect = [np.float32(env.model.q_init()[i]) + np.float32(orc.reset_uniform(77, w, 1, i)) * np.float32(0.005)
                  for i in range(env.model.n_dofs)]
        assert np.array_equal(q2[w].cpu().numpy().astype(np.float32), np.array(expect, dtype=np.float32))
    env.close()


def test_batched_host_step_returns_fresh_arrays_of_reference_types():
    """VectorEnv(copy=True) semantics on the pinned output pool: arrays the caller keeps are never overwritten."""
    env = _make("DartHopper-v1", num_envs=128, output="numpy", seed=3)
    env.reset()
    rng = np.random.RandomState(0)
    kept = []
    for i in range(6):
        o, r, d, info = env.step(rng.uniform(-1, 1, (128, 3)).astype(np.float32))
        assert o.dtype == np.float32 and r.dtype == np.float64 and d.dtype == np.bool_ and "TimeLimit.truncated" in info
        kept.append((o, o.copy(), r, r.copy(), d, d.copy()))
    for o, oc, r, rc, d, dc in kept:
        assert np.array_equal(o, oc) and np.array_equal(r, rc) and np.array_equal(d, dc)
    assert len({id(k[0]) for k in kept}) == 6
    view = kept[0][0][3]            # a VIEW keeps its slot out of circulation too
    del kept
    vc = view.copy()
    for i in range(80):
        env.step(rng.uniform(-1, 1, (128, 3)).astype(np.float32))
    assert np.array_equal(view, vc)
    assert len(env._pool.slots) <= 7     # the six arrays kept above + one in flight: the pool did not grow while stepping
    env.close()


def test_tma_staged_prologue_is_bit_identical():
    """k_env_step_coop with its CTA tile staged by TMA bulk copies (cp.async.bulk + mbarrier) and the observation tile
    written by one bulk store (DARTB_COOP_TMA=1; 2 = the actions too) returns the same bits as the plain-load form."""
    digests = {}
    for mode in ("0", "1", "2"):
        env = dict(os.environ, DARTB_COOP_TMA=mode)
        res = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "coop_tma_check.py")], env=env, capture_output=True,
                             text=True, timeout=600)
        assert res.returncode == 0, res.stdout + res.stderr
        digests[mode] = [ln for ln in res.stdout.splitlines() if ln.startswith("digest")][-1]
    assert digests["0"] == digests["1"] == digests["2"], digests


@pytest.mark.parametrize("variant,n", [(0, 300), (2, 300), (3, 300)])
def test_fused_obs_gather_peer_stores(variant, n):
    """dartb_set_obs_peers: the step kernel also stores every observation row into the peer buffers at the rank's offset
    (the obs all-gather fused into the kernel; on one GPU the "peers" are two more local buffers).  Every kernel form."""
    from dart_env_b200.engine import Engine
    from dart_env_b200.skel import load_model
    spec = SPECS["DartWalker2d-v1"]
    m = load_model(spec.skel, spec.dt)
    m.enforce_limits()
    eng = Engine(m, spec.task, n, seed=2, kernel_variant=variant)
    eng.reset()
    no = spec.task.n_obs
    peers = [torch.full((3 * n, no), -7.0, dtype=torch.float32, device="cuda") for _ in range(2)]
    eng.set_obs_peers([p.data_ptr() for p in peers], 1 * n * no)        # "rank 1 of 3"
    obs = torch.empty((n, no), dtype=torch.float32, device="cuda")
    rew = torch.empty((n,), dtype=torch.float32, device="cuda"); done = torch.empty((n,), dtype=torch.uint8, device="cuda")
    gen = torch.Generator(device="cuda"); gen.manual_seed(1)
    for _ in range(5):
        eng.step(torch.rand((n, 6), generator=gen, device="cuda") * 2 - 1, obs, rew, done, True)
    torch.cuda.synchronize()
    for p in peers:
        assert torch.equal(p[n:2 * n], obs) and (p[:n] == -7.0).all() and (p[2 * n:] == -7.0).all()
    eng.set_obs_peers([], 0)
    peers[0].fill_(-7.0)
    eng.step(torch.zeros((n, 6), device="cuda"), obs, rew, done, True)
    torch.cuda.synchronize()
    assert (peers[0] == -7.0).all()
    eng.close()


@pytest.mark.parametrize("f64", [True, False])
def test_per_world_body_params_match_oracle(f64):
    """SURVEY 8f.2 dynamics randomisation (snake_7link.py:115-120: bodynodes[i].set_mass / set_friction_coeff): with
    dartb_set_body_params every world steps with ITS masses and friction coefficients — compared, world by world, with an
    oracle world that had orc_set_mass / orc_set_friction applied; clearing returns to the shared model and its kernel."""
    from oracle import oracle as orc
    env_id = "DartWalker2d-v1"
    g = np.load(os.path.join(GOLD, "walker2d.npz"))
    idx = np.concatenate([np.where(g["sub_ncontact"] > 0)[0][:40], np.arange(24)])
    n = len(idx)
    env = _make(env_id, num_envs=n, output="numpy", seed=0, auto_reset=False, f64=f64)
    auto_name = env.engine.kernel_name
    nb = len(env.model.bodies)
    rng = np.random.RandomState(11)
    mass0 = np.array([b.mass for b in env.model.bodies])
    fr0 = np.array([b.friction_coeff for b in env.model.bodies])
    mass = np.clip(mass0 + rng.uniform(-1.5, 1.5, (n, nb)), 0.05, None)
    fric = np.clip(fr0 + rng.uniform(-0.5, 0.5, (n, nb)), 0.0, None)
    env.set_body_params(mass, fric)
    assert "loop:generic" in env.engine.kernel_name
    env.set_state(g["sub_q"][idx], g["sub_dq"][idx])
    env.do_simulation(g["sub_tau"][idx], 1)
    s = env.state_vector()
    nd = env.model.n_dofs
    worst = 0.0
    for k, i in enumerate(idx):
        w = orc.OracleWorld(env.model)
        for b in range(nb):
            w.set_mass(b, mass[k, b]); w.set_friction(b, fric[k, b])
        w.set_state(g["sub_q"][i], g["sub_dq"][i]); w.set_forces(g["sub_tau"][i]); w.step()
        oq, odq = w.get_state()
        if not f64 and (g["sub_contact_margin"][i] < 1e-4 or g["sub_limit_margin"][i] < 1e-4 or g["sub_tie_margin"][i] < 1e-4):
            continue   # the golden's tagged event-margin samples (fp32 can take the other side of a threshold)
        worst = max(worst, float((np.abs(s[k, nd:] - odq) / (1 + np.abs(odq))).max()), float(np.abs(s[k, :nd] - oq).max()))
    assert worst < (1e-8 if f64 else 5e-4), worst   # (the CPU build of the same kernel source stays below 1e-4 on these samples)
    # the fused env.step() reads the same table
    env.set_state(g["sub_q"][idx], g["sub_dq"][idx])
    ob, rew, done, _ = env.step(np.zeros((n, env.act_dim), dtype=np.float32))
    assert np.isfinite(ob).all() and np.isfinite(rew).all()
    # both None: back to the shared model on the automatically chosen kernel, bit-identical to a fresh env
    env.set_body_params(None, None)
    assert env.engine.kernel_name == auto_name
    env2 = _make(env_id, num_envs=n, output="numpy", seed=0, auto_reset=False, f64=f64)
    for e in (env, env2):
        e.set_state(g["sub_q"][idx], g["sub_dq"][idx])
        e.do_simulation(g["sub_tau"][idx], 1)
    assert np.array_equal(env.state_vector(), env2.state_vector())
    # bad input fails loudly and leaves the engine usable
    from dart_env_b200.capi import DartbError
    with pytest.raises(DartbError):
        env.set_body_params(-np.ones((n, nb)), None)
    env.do_simulation(g["sub_tau"][idx], 1)
    env.close(); env2.close()


def _expected_redraw(base, half_range, seed, world, episode, idx0, lo=0.0, hi=None):
    """what redraw_dynamics (kernels.cuh) writes: base + reset_uniform(seed, world, episode, idx0 + i) * half_range in fp32,
    clipped; the generator is the oracle's bit-exact twin of the kernel's (test_reset_noise_*)"""
    from oracle import oracle as orc
    out = np.empty(len(base), dtype=np.float64)
    for i, b in enumerate(base):
        u = np.float32(orc.reset_uniform(seed, world, episode, idx0 + i))
        v = np.float32(np.float32(b) + np.float32(u * np.float32(half_range)))
        v = max(float(v), lo)
        out[i] = v if hi is None else min(v, hi)
    return out


def test_snake_randomize_dynamics_flag():
    """snake_7link.py:11,20-25,115-120: `randomize_dynamics` (hard-coded off in the reference) redraws every bodynode's
    mass / friction at EVERY reset_model.  Here inside the kernel, from the seeded reset generator: the table after
    reset() is exactly original + 1.5 u (clipped at 0), identically seeded worlds diverge only through their masses."""
    from dart_env_b200.cstructs import PM_MAXB
    from dart_env_b200.envs import DartSnake7LinkEnv
    n = 64
    a = np.random.RandomState(0).uniform(-1, 1, (10, 1, 6)).astype(np.float32).repeat(n, 1)
    outs = []
    for rd in (False, True):
        env = DartSnake7LinkEnv(num_envs=n, output="numpy", seed=5, auto_reset=False, randomize_dynamics=rd)
        assert ("loop:generic" in env.engine.kernel_name) == rd
        if rd:
            env.reset()
            tab = env.engine.get_body_table()
            nb = env.model.n_dofs
            mass0 = np.array(env.bodynode_original_masses)
            assert tab.shape == (4 * nb, n)                      # the snake has no capsule that can touch the ground
            # one consistent episode index for the whole batch (the reset that just ran), one draw per body and world
            hits = [ep for ep in range(4) if all(np.array_equal(tab[:nb, w].astype(np.float32),
                    _expected_redraw(mass0, 1.5, 5, w, ep, 2 * PM_MAXB).astype(np.float32)) for w in (0, 1, n - 1))]
            assert len(hits) == 1, hits
            assert (tab[:nb] >= 0).all() and np.abs(tab[:nb] - mass0[:, None]).max() <= 1.5 + 1e-6
            assert np.abs(tab[:nb] - mass0[:, None]).max() > 1.0             # the range is used
            tab2 = (env.reset(), env.engine.get_body_table())[1]
            assert np.abs(tab2[:nb] - tab[:nb]).max() > 0.1                  # every reset draws again
        env.seed([7] * n)            # from here on every world draws the same reset noise (and the same masses)
        env.reset()
        if rd:
            t3 = env.engine.get_body_table()
            assert np.abs(t3 - t3[:, :1]).max() == 0.0
            env.seed(list(range(100, 100 + n)))                  # distinct draws again, then identical STATE noise is gone too:
            env.reset()                                          # compare the two modes on their spread instead
        ob = None
        for t in range(10):
            ob, rew, done, _ = env.step(a[t])
        assert np.isfinite(ob).all()
        outs.append(ob.copy())
        env.close()
    assert np.abs(outs[0] - outs[0][0]).max() == 0.0          # shared model + one seed: identical worlds stay identical
    assert np.abs(outs[1] - outs[1][0]).max() > 1e-3


@pytest.mark.parametrize("f64", [True, False])
def test_randomize_dynamics_redraws_at_every_auto_reset(f64):
    """DARTB_OPT_RANDOMIZE_MASS / _FRICTION on a contact-rich env with short episodes (Hopper): a world that finishes an
    episode inside step() gets new masses and friction coefficients with its reset state; the others keep theirs; the
    values are the documented function of (seed, world, episode)."""
    from dart_env_b200.capi import DartbError
    from dart_env_b200.cstructs import PM_MAXB
    n = 96
    env = _make("DartHopper-v1", num_envs=n, output="numpy", seed=9, f64=f64)
    nb = env.model.n_dofs
    assert len(env.model.bodies) == nb                       # no welded bodynodes
    env.engine.set_randomize(0.5, 0.3)
    assert "loop:generic" in env.engine.kernel_name
    env.reset()
    t0 = env.engine.get_body_table()
    ns = t0.shape[0] - 4 * nb
    assert ns == 4
    mass0 = np.array([b.mass for b in env.model.bodies])
    ever_done = np.zeros(n, dtype=bool)
    rng = np.random.RandomState(2)
    for _ in range(40):     # (random actions end a Hopper episode after ~4 steps: stop while some worlds are still in their first)
        _, _, done, _ = env.step(rng.uniform(-1, 1, (n, 3)).astype(np.float32))
        ever_done |= done
        if ever_done.sum() >= n // 4:
            break
    assert ever_done.any() and not ever_done.all()
    t1 = env.engine.get_body_table()
    changed = np.abs(t1 - t0).max(0) > 0
    assert np.array_equal(changed, ever_done)
    assert np.array_equal(t1[nb:4 * nb], t0[nb:4 * nb])          # COM offsets and izz are not redrawn (set_mass keeps them)
    assert (t1[:nb] >= 0).all() and np.abs(t1[:nb] - mass0[:, None]).max() <= 0.5 + 1e-6
    mu = t1[4 * nb:]
    assert (mu >= 0.7 - 1e-6).all() and (mu <= 1.0).all()       # base 1.0 -+ 0.3, min(body, ground = 1)
    w = int(np.where(~ever_done)[0][0])                          # a world still in its first episode: drawn by reset()
    hits = [ep for ep in range(4) if np.array_equal(t1[:nb, w].astype(np.float32),
                                                    _expected_redraw(mass0, 0.5, 9, w, ep, 2 * PM_MAXB).astype(np.float32))]
    assert len(hits) == 1, hits
    # a skeleton with welded bodynodes is refused (their masses mix on the host: dartb_set_body_params)
    ch = _make("DartHalfCheetah-v1", num_envs=8, output="numpy", seed=0)
    if len(ch.model.bodies) != ch.model.n_dofs:
        with pytest.raises(DartbError):
            ch.engine.set_randomize(0.5, 0.0)
    ch.close()
    env.engine.set_randomize(0.0, 0.0)
    env.set_body_params(None, None)
    assert "loop:generic" not in env.engine.kernel_name
    env.close()


def test_bodynode_views_set_mass_and_friction():
    """the pydart2 spellings the reference uses (snake_7link.py:24-25,117-120): bn.mass(), bn.friction_coeff(),
    bn.set_mass(m), bn.set_friction_coeff(mu) — a scalar reaches every world, an array one world each"""
    n = 16
    env = _make("DartHopper-v1", num_envs=n, output="numpy", seed=0, auto_reset=False, f64=True)
    bn = env.robot_skeleton.bodynodes[3]
    m0, f0 = bn.mass(), bn.friction_coeff()
    assert isinstance(m0, float) and isinstance(f0, float) and bn.m == m0
    env.reset()
    s0 = env.state_vector().copy()
    tau = np.zeros((n, env.model.n_dofs)); tau[:, 3:] = 30.0
    env.do_simulation(tau, 4)
    base = env.state_vector().copy()
    bn.set_mass(m0 + 1.0)
    assert bn.mass() == m0 + 1.0 and "loop:generic" in env.engine.kernel_name
    env.set_state_vector(s0); env.do_simulation(tau, 4)
    heavier = env.state_vector().copy()
    assert np.abs(heavier - base).max() > 1e-4
    per_world = m0 + np.linspace(0.0, 1.0, n)
    bn.set_mass(per_world)
    assert np.array_equal(bn.mass(), per_world)
    env.set_state_vector(s0); env.do_simulation(tau, 4)
    s = env.state_vector()
    assert np.allclose(s[0], base[0], rtol=1e-9, atol=1e-9)        # world 0 kept the original mass: same as the compiled kernel, fp64
    assert np.allclose(s[-1], heavier[-1], rtol=1e-12, atol=1e-12)   # world n-1 has m0 + 1
    bn.set_friction_coeff(0.25)
    assert bn.friction_coeff() == 0.25 and np.array_equal(bn.mass(), per_world)
    env.set_body_params(None, None)
    assert bn.mass() == m0 and bn.friction_coeff() == f0 and "loop:generic" not in env.engine.kernel_name
    env.close()
