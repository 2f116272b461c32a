"""GPU parity tests (run on the B200 box): the CUDA path, called through the C-ABI, against
the CPU oracle and the committed golden vectors.

Tolerances (stated once here and in DESIGN.md §2b; per north_star "fp32 tolerance; bit-exact contact-pair sets and
done flags").  All errors are MAXIMA over every sample outside the explicitly tagged classes, relative to (1 + |x|):
  fp64 kernel instantiation vs oracle (same algorithm, different formulation):  1e-9
  fp32 product path, single DART step from identical (q, dq, tau):   dq 2e-4 (TOL_SUB_DQ), q 1e-5 (TOL_SUB_Q)
  fp32 env.step (4-5 DART steps + task layer):   obs 2e-4 (TOL_STEP_OBS), reward 1e-3 (TOL_STEP_REW)
  contact-pair index sets, limit sets and done flags: bit-exact.
Tagged classes (the goldens carry the tags; nothing else is excluded):
  margin   a discrete decision (contact on/off, flat-capsule end tie, joint limit on/off, termination threshold)
           sits within MARGIN = 1e-4 of its threshold, so fp32 rounding may legitimately flip it;
  deep     a hand-built contact sweep state with a capsule more than DEEP = 0.03 m inside the ground (30x the
           velocity-correction cap): contact impulses amplify rounding through A^-1, cond(A) -> 1/CFM = 1e5.  These stay
           bounded by TOL_SUB_DQ_DEEP = 2e-2 (q: 2e-4) and keep their bit-exact contact sets.
Every test prints p50 / p99 / max of what it measured (pytest -s, and gpurun_out/parity_errors.log when writable).
"""
import os

import numpy as np
import pytest
import torch

from dart_env_b200.tasks import SPECS

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(__file__), "golden")
FILES = {"DartHopper-v1": "hopper.npz", "DartWalker2d-v1": "walker2d.npz",
         "DartHalfCheetah-v1": "halfcheetah.npz", "DartSnake7Link-v1": "snake7link.npz"}
ENVS = list(SPECS)
MARGIN = 1e-4
DEEP = 0.03
TOL_SUB_DQ, TOL_SUB_Q, TOL_SUB_DQ_DEEP, TOL_SUB_Q_DEEP = 2e-4, 1e-5, 2e-2, 2e-4
TOL_STEP_OBS, TOL_STEP_REW = 2e-4, 1e-3


def _report(name, env_id, err):
    err = np.asarray(err, dtype=np.float64).ravel()
    line = "%-34s %-20s variant=%s n=%4d p50 %.2e p99 %.2e max %.2e" % (
        name, env_id, VARIANT, err.size, np.median(err), np.percentile(err, 99), err.max())
    print(line)
    try:
        out = os.path.join(os.path.dirname(os.path.dirname(__file__)), "gpurun_out")
        if os.path.isdir(out):
            with open(os.path.join(out, "parity_errors.log"), "a") as f:
                f.write(line + "\n")
    except OSError:
        pass


VARIANT = 0   # this module pins the one-world-per-thread kernels; tests/test_gpu_coop.py re-runs its tests with the
              # lane-cooperative kernels (2), tests/test_gpu_quad.py with the quad form (3).  (None = automatic choice.)


def _engine(models, env_id, n, **kw):
    from dart_env_b200.engine import Engine
    if VARIANT is not None:
        kw.setdefault("kernel_variant", VARIANT)
    return Engine(models[env_id], SPECS[env_id].task, n, **kw)


def _make(env_id, **kw):
    from dart_env_b200.envs import make
    if VARIANT is not None:
        kw.setdefault("kernel_variant", VARIANT)
    return make(env_id, **kw)


def _gold(env_id):
    return np.load(os.path.join(GOLD, FILES[env_id]))


def _substep(models, env_id, g, f64):
    dev = torch.device("cuda", 0)
    dt = torch.float64 if f64 else torch.float32
    n = len(g["sub_q"])
    eng = _engine(models, env_id, n, f64=f64)
    eng.set_state(torch.tensor(g["sub_q"], dtype=dt, device=dev), torch.tensor(g["sub_dq"], dtype=dt, device=dev))
    if VARIANT in (2, 3):
        # the cooperative / quad kernels take no external forces (the engine routes those to the per-thread kernel):
        # step everything without them and let the caller look at the samples that have none
        eng.substep(torch.tensor(g["sub_tau"], dtype=dt, device=dev))
    else:
        eng.substep(torch.tensor(g["sub_tau"], dtype=dt, device=dev), torch.tensor(g["sub_fext"], dtype=dt, device=dev).contiguous())
    q2, dq2 = eng.get_state(torch.float64)
    cnt, body, data = eng.contacts()
    torch.cuda.synchronize()
    out = q2.cpu().numpy(), dq2.cpu().numpy(), cnt.cpu().numpy(), body.cpu().numpy(), data.cpu().numpy()
    assert eng.launch_count >= 3 and any(k in eng.kernel_name for k in ("static:", "loop:", "coop:", "quad:"))
    if VARIANT in (2, 3):
        assert ("coop:" if VARIANT == 2 else "quad:") in eng.kernel_name
    eng.close()
    return out


def _no_fext(g):
    """samples the cooperative kernel can be compared on (see _substep)"""
    if VARIANT not in (2, 3):
        return np.ones(len(g["sub_q"]), dtype=bool)
    return np.abs(g["sub_fext"]).reshape(len(g["sub_q"]), -1).max(1) == 0


@pytest.mark.parametrize("env_id", ENVS)
def test_substep_fp64_matches_oracle_tightly(models, env_id):
    g = _gold(env_id)
    q2, dq2, cnt, body, data = _substep(models, env_id, g, True)
    nf = _no_fext(g)
    # (samples whose contact decision sits within 1e-9 of the threshold may flip: excluded from everything)
    ok = nf & (g["sub_contact_margin"] > 1e-9) if VARIANT in (2, 3) else nf
    assert np.allclose(q2[ok], g["sub_q2"][ok], rtol=1e-9, atol=1e-10)
    assert np.allclose(dq2[ok], g["sub_dq2"][ok], rtol=1e-8, atol=1e-8)
    safe = nf & (g["sub_contact_margin"] > 1e-9)
    assert np.array_equal(cnt[safe], g["sub_ncontact"][safe])
    mc = min(body.shape[1], g["sub_contact_body"].shape[1])
    tie_ok = safe & (g["sub_tie_margin"] > 1e-9)
    assert np.array_equal(body[tie_ok, :mc], g["sub_contact_body"][tie_ok, :mc])
    # contact geometry + force (pydart2 contact.force), where the tie rule cannot flip the end
    gd = g["sub_contact_data"][tie_ok][:, :mc]
    assert np.allclose(data[tie_ok][:, :mc, :7], gd[..., :7], atol=1e-6)
    assert np.allclose(data[tie_ok][:, :mc, 7:], gd[..., 7:], rtol=1e-4, atol=1e-2)


@pytest.mark.parametrize("env_id", ENVS)
def test_substep_fp32_within_stated_tolerance(models, env_id):
    g = _gold(env_id)
    q2, dq2, cnt, body, data = _substep(models, env_id, g, False)
    safe = _no_fext(g) & (g["sub_contact_margin"] > MARGIN) & (g["sub_limit_margin"] > MARGIN) & (g["sub_tie_margin"] > MARGIN)
    assert safe.sum() > 0.5 * len(safe)
    # bit-exact discrete outcomes
    assert np.array_equal(cnt[safe], g["sub_ncontact"][safe])
    mc = min(body.shape[1], g["sub_contact_body"].shape[1])
    assert np.array_equal(body[safe, :mc], g["sub_contact_body"][safe, :mc])
    eq = (np.abs(q2 - g["sub_q2"]) / (1 + np.abs(g["sub_q2"]))).max(1)
    ev = (np.abs(dq2 - g["sub_dq2"]) / (1 + np.abs(g["sub_dq2"]))).max(1)
    deep = g["sub_contact_data"][:, :, 6].max(1) > DEEP
    _report("substep dq (not deep)", env_id, ev[safe & ~deep])
    _report("substep q (not deep)", env_id, eq[safe & ~deep])
    assert ev[safe & ~deep].max() < TOL_SUB_DQ, ev[safe & ~deep].max()
    assert eq[safe & ~deep].max() < TOL_SUB_Q, eq[safe & ~deep].max()
    if (safe & deep).any():
        _report("substep dq (deep penetration)", env_id, ev[safe & deep])
        assert ev[safe & deep].max() < TOL_SUB_DQ_DEEP, ev[safe & deep].max()
        assert eq[safe & deep].max() < TOL_SUB_Q_DEEP, eq[safe & deep].max()


def _envstep(models, env_id, g, f64):
    dev = torch.device("cuda", 0)
    dt = torch.float64 if f64 else torch.float32
    spec = SPECS[env_id]
    n = len(g["step_q"])
    eng = _engine(models, env_id, n, f64=f64)
    eng.set_state(torch.tensor(g["step_q"], dtype=dt, device=dev), torch.tensor(g["step_dq"], dtype=dt, device=dev))
    obs = torch.empty((n, spec.task.n_obs), dtype=torch.float32, device=dev)
    rew = torch.empty((n,), dtype=torch.float32, device=dev)
    done = torch.empty((n,), dtype=torch.uint8, device=dev)
    eng.step(torch.tensor(g["step_action"], dtype=torch.float32, device=dev), obs, rew, done, auto_reset=False)
    q2, dq2 = eng.get_state(torch.float64)
    torch.cuda.synchronize()
    out = obs.cpu().numpy().astype(np.float64), rew.cpu().numpy().astype(np.float64), done.cpu().numpy().astype(bool), \
        q2.cpu().numpy(), dq2.cpu().numpy()
    eng.close()
    return out


@pytest.mark.parametrize("env_id", ENVS)
def test_env_step_matches_reference_task_layer(models, env_id):
    """goldens: reference env classes (hopper.py ...) run on the oracle; here the fused kernel."""
    g = _gold(env_id)
    fin = np.isfinite(g["step_obs"]).all(1) & np.isfinite(g["step_reward"])
    # fp64 instantiation: algorithmic equivalence (obs are stored as fp32 at the boundary)
    obs, rew, done, q2, dq2 = _envstep(models, env_id, g, True)
    # (actions cross the boundary as fp32, so even the fp64 kernel sees tau rounded to 6e-8 relative)
    assert np.allclose(q2[fin], g["step_q2"][fin], rtol=1e-6, atol=1e-7)
    assert np.allclose(dq2[fin], g["step_dq2"][fin], rtol=2e-5, atol=2e-5)
    assert np.allclose(obs[fin], g["step_obs"][fin], rtol=1e-5, atol=1e-5)
    assert np.allclose(rew[fin], g["step_reward"][fin], rtol=1e-5, atol=1e-4)
    safe = g["step_margin"] > 1e-7
    assert np.array_equal(done[safe], g["step_done"][safe].astype(bool))
    # fp32 product path
    obs, rew, done, q2, dq2 = _envstep(models, env_id, g, False)
    safe = fin & (g["step_margin"] > MARGIN)
    assert np.array_equal(done[safe], g["step_done"][safe].astype(bool))
    # maxima over every sample whose discrete decisions (contact / tie / limit, in any of the frame_skip DART steps)
    # are not within MARGIN of flipping
    ok = fin & (g["step_event_margin"] > MARGIN)
    assert ok.sum() > 0.4 * len(ok)
    eo = (np.abs(obs - g["step_obs"]) / (1 + np.abs(g["step_obs"])))[ok].max(1)
    er = (np.abs(rew - g["step_reward"]) / (1 + np.abs(g["step_reward"])))[ok]
    _report("env.step obs", env_id, eo)
    _report("env.step reward", env_id, er)
    assert eo.max() < TOL_STEP_OBS, eo.max()
    assert er.max() < TOL_STEP_REW, er.max()


@pytest.mark.parametrize("env_id", ENVS)
def test_reset_noise_bit_exact_and_sharding_independent(models, env_id):
    from oracle import oracle as orc
    spec = SPECS[env_id]
    n = 64
    eng = _engine(models, env_id, n, seed=1234, world_offset=1000)
    obs = eng.reset().cpu().numpy()
    q, dq = eng.get_state(torch.float64)
    q, dq = q.cpu().numpy(), dq.cpu().numpy()
    for w in (0, 1, 17, 63):
        e = orc.OracleEnv(models[env_id], spec.task, seed=1234, world_id=1000 + w)
        o = e.reset()
        oq, odq = e.world.get_state()
        assert np.array_equal(q[w], oq) and np.array_equal(dq[w], odq)  # bit-exact noise
        assert np.allclose(obs[w], o, atol=1e-6)
    assert np.abs(q - models[env_id].q_init()).max() <= 0.005 * (1 + 1e-6)
    # a second engine holding only worlds 1032.. reproduces the same states (sharding independence)
    eng2 = _engine(models, env_id, 32, seed=1234, world_offset=1032)
    eng2.reset()
    q2, _ = eng2.get_state(torch.float64)
    assert np.array_equal(q2.cpu().numpy(), q[32:])
    # second episode differs from the first; masked reset only touches masked worlds
    mask = torch.zeros(n, dtype=torch.uint8, device="cuda")
    mask[::2] = 1
    eng.reset(mask)
    q3, _ = eng.get_state(torch.float64)
    q3 = q3.cpu().numpy()
    assert np.array_equal(q3[1::2], q[1::2]) and not np.array_equal(q3[::2], q[::2])
    eng.close(); eng2.close()


@pytest.mark.parametrize("env_id,sweep", [("DartWalker2d-v1", (1, 4, 30)), ("DartSnake7Link-v1", (1, 2, 4, 8, 16, 30, 50))])
def test_pgs_mode_matches_oracle_pgs(models, env_id, sweep):
    """contact-rich PGS path (BASELINE config 3) and the Snake PGS iteration sweep (config 5: its rows are
    joint-limit rows only): same sweep count -> same iterate as the oracle's PGS."""
    from oracle import oracle as orc
    g = _gold(env_id)
    has_rows = (g["sub_ncontact"] > 0) if env_id != "DartSnake7Link-v1" else (np.abs(g["sub_limit_active"]).sum(1) > 0)
    idx = np.where(_no_fext(g) & has_rows & (g["sub_contact_margin"] > MARGIN) & (g["sub_tie_margin"] > MARGIN)
                   & (g["sub_limit_margin"] > MARGIN))[0][:40]
    assert len(idx) >= 8
    dev = torch.device("cuda", 0)
    for iters in sweep:
        ref = []
        w = orc.OracleWorld(models[env_id])
        w.set_option(1, 1); w.set_option(2, iters)
        for i in idx:
            w.set_state(g["sub_q"][i], g["sub_dq"][i]); w.set_forces(g["sub_tau"][i]); w.step()
            ref.append(np.concatenate(w.get_state()))
        ref = np.array(ref)
        eng = _engine(models, env_id, len(idx), f64=True)
        eng.set_lcp(1, iters)
        eng.set_state(torch.tensor(g["sub_q"][idx], device=dev), torch.tensor(g["sub_dq"][idx], device=dev))
        eng.substep(torch.tensor(g["sub_tau"][idx], device=dev))
        q2, dq2 = eng.get_state(torch.float64)
        got = torch.cat([q2, dq2], 1).cpu().numpy()
        assert np.allclose(got, ref, rtol=1e-8, atol=1e-8), iters
        eng.close()


def test_full_size_properties_hopper_4096(models):
    """BASELINE config 2 size: determinism, batch == singles, auto-reset semantics, flags."""
    env_id, n = "DartHopper-v1", 4096
    spec = SPECS[env_id]
    dev = torch.device("cuda", 0)
    gen = torch.Generator(device=dev); gen.manual_seed(7)
    acts = [torch.rand((n, 3), generator=gen, device=dev) * 2.6 - 1.3 for _ in range(30)]

    def run(nw, off):
        eng = _engine(models, env_id, nw, seed=5, world_offset=off)
        obs = eng.reset()
        rew = torch.empty((nw,), dtype=torch.float32, device=dev); done = torch.empty((nw,), dtype=torch.uint8, device=dev)
        tot_done = 0
        hist = []
        for a in acts:
            eng.step(a[off:off + nw].contiguous(), obs, rew, done, True)
            hist.append((obs.clone(), rew.clone(), done.clone()))
            tot_done += int(done.sum())
        q, dq = eng.get_state()
        eng.close()
        return hist, q, dq, tot_done

    h1, q1, dq1, nd1 = run(n, 0)
    h2, q2, dq2, nd2 = run(n, 0)
    assert torch.equal(q1, q2) and torch.equal(dq1, dq2) and nd1 == nd2 and nd1 > n  # many episodes ended
    for (o1, r1, d1), (o2, r2, d2) in zip(h1, h2):
        assert torch.equal(o1, o2) and torch.equal(r1, r2) and torch.equal(d1, d2)
    # shard of worlds [1024, 1024+96) stepped alone gives bit-identical results
    h3, q3, dq3, _ = run(96, 1024)
    assert torch.equal(q3, q1[1024:1120]) and torch.equal(dq3, dq1[1024:1120])
    assert torch.equal(h3[-1][0], h1[-1][0][1024:1120])
    # everything finite, obs of done worlds is a reset obs (|q - q0| <= noise)
    o, r, d = h1[-1]
    assert torch.isfinite(o).all() and torch.isfinite(r).all()
    dm = d.bool()
    if dm.any():
        assert (o[dm][:, 1:5].abs() <= 0.005001).all() and ((o[dm][:, 0] - 1.25).abs() < 0.011).all()


def test_time_limit_truncation(models):
    env = _make("DartSnake7Link-v1", num_envs=8, output="numpy", seed=0, max_episode_steps=5)
    env.reset()
    for t in range(5):
        obs, rew, done, info = env.step(np.zeros((8, 6), dtype=np.float32))
        assert done.all() == (t == 4)
    assert info["TimeLimit.truncated"].all()
    obs, rew, done, info = env.step(np.zeros((8, 6), dtype=np.float32))
    assert not done.any()
    env.close()


def test_gym_surface_single_env_types(models):
    """test_envs.py:10-37 shape/type contract for the N = 1 adapter (BASELINE config 1 plumbing)."""
    for env_id in ENVS:
        env = _make(env_id, seed=0)
        ob = env.reset()
        assert env.observation_space.contains(ob) and ob.dtype == np.float64
        env.action_space.seed(0)
        a = env.action_space.sample()
        assert env.action_space.contains(a)
        ob, r, d, info = env.step(a)
        assert env.observation_space.contains(ob) and np.isscalar(r) and isinstance(d, bool) and isinstance(info, dict)
        assert env.dt == pytest.approx(SPECS[env_id].dt * SPECS[env_id].task.frame_skip)
        s = env.state_vector()
        assert s.shape == (2 * env.model.n_dofs,)
        # determinism (test_determinism.py): same seed, same actions -> identical obs
        env2 = _make(env_id, seed=0)
        ob2 = env2.reset()
        ob2, r2, d2, _ = env2.step(a)
        assert np.array_equal(ob, ob2) and r == r2 and d == d2
        # do_simulation == frame_skip explicit sub-steps (dart_env.py:158-175)
        env.set_state(s[:env.model.n_dofs], s[env.model.n_dofs:])
        tau = np.zeros(env.model.n_dofs)
        env.do_simulation(tau, env.frame_skip)
        assert np.isfinite(env.state_vector()).all()
        env.close(); env2.close()


def test_contacts_readback_walker(models):
    env = _make("DartWalker2d-v1", num_envs=256, output="torch", seed=3, contacts=True)
    env.reset()
    for _ in range(80):
        env.step(torch.zeros((256, 6), device="cuda"))
    cnt, body, data = env.contacts()
    assert int(cnt.max()) >= 1
    w = int(torch.argmax(cnt))
    c = int(cnt[w])
    assert (body[w, :c] >= 2).all() and (body[w, c:] == -1).all()
    f = data[w, :c, 7:10]
    assert (f[:, 1] >= -1e-3).all() and f[:, 1].sum() > 1.0  # ground pushes up
    assert torch.allclose(data[w, :c, 3:6].norm(dim=1), torch.ones(c, device="cuda"), atol=1e-5)
    env.close()


@pytest.mark.parametrize("env_id,n", [("DartWalker2d-v1", 16384), ("DartHalfCheetah-v1", 16384), ("DartSnake7Link-v1", 4096)])
def test_full_size_properties_other_configs(models, env_id, n):
    """BASELINE configs 3-5 at their per-GPU sizes: replay determinism, shard == batch (bit-exact),
    finite outputs, PGS and exact modes agree on done flags for the first steps."""
    spec = SPECS[env_id]
    dev = torch.device("cuda", 0)
    gen = torch.Generator(device=dev); gen.manual_seed(11)
    acts = [torch.rand((n, spec.task.n_act), generator=gen, device=dev) * 2 - 1 for _ in range(12)]

    def run(nw, off, lcp=None):
        eng = _engine(models, env_id, nw, seed=9, world_offset=off)
        if lcp is not None:
            eng.set_lcp(*lcp)
        obs = eng.reset()
        rew = torch.empty((nw,), dtype=torch.float32, device=dev); done = torch.empty((nw,), dtype=torch.uint8, device=dev)
        dones = []
        for a in acts:
            eng.step(a[off:off + nw].contiguous(), obs, rew, done, True)
            dones.append(done.clone())
        q, dq = eng.get_state()
        out = (obs.clone(), rew.clone(), q, dq, dones)
        eng.close()
        return out

    o1, r1, q1, dq1, d1 = run(n, 0)
    o2, r2, q2, dq2, d2 = run(n, 0)
    assert torch.equal(q1, q2) and torch.equal(dq1, dq2) and torch.equal(o1, o2) and torch.equal(r1, r2)
    assert torch.isfinite(o1).all() and torch.isfinite(r1).all() and torch.isfinite(q1).all()
    off = n // 2 + 32
    o3, r3, q3, dq3, d3 = run(64, off)
    assert torch.equal(q3, q1[off:off + 64]) and torch.equal(o3, o1[off:off + 64]) and torch.equal(r3, r1[off:off + 64])
    # PGS(30) is an approximation of the same LCP: the first env step's done flags agree almost everywhere
    o4, r4, q4, dq4, d4 = run(2048, 0, lcp=(1, 30))
    agree = (d4[0] == d1[0][:2048]).float().mean().item()
    assert agree > 0.98


@pytest.mark.parametrize("env_id", ["DartHopper-v1", "DartWalker2d-v1"])
def test_rollout_statistics_match_oracle(models, env_id):
    """Long rollouts diverge chaotically after contact, so compare STATISTICS under the same random
    policy: mean episode length and mean per-step reward of the fp32 GPU engine vs the fp64 oracle."""
    from oracle import oracle as orc
    spec = SPECS[env_id]
    rng = np.random.RandomState(3)
    # oracle: 48 worlds x 250 steps
    lens, rews = [], []
    for w in range(48):
        e = orc.OracleEnv(models[env_id], spec.task, seed=21, world_id=w)
        e.reset()
        cur = 0
        for t in range(250):
            o, r, d = e.step(rng.uniform(-1, 1, spec.task.n_act))
            rews.append(r); cur += 1
            if d:
                lens.append(cur); cur = 0
                e.reset()
    o_len, o_rew = np.mean(lens), np.mean(rews)
    n = 4096
    dev = torch.device("cuda", 0)
    eng = _engine(models, env_id, n, seed=21)
    obs = eng.reset()
    rew = torch.empty((n,), dtype=torch.float32, device=dev); done = torch.empty((n,), dtype=torch.uint8, device=dev)
    gen = torch.Generator(device=dev); gen.manual_seed(5)
    tot_r, tot_d, steps = 0.0, 0, 250
    for t in range(steps):
        a = torch.rand((n, spec.task.n_act), generator=gen, device=dev) * 2 - 1
        eng.step(a, obs, rew, done, True)
        tot_r += float(rew.double().mean()); tot_d += int(done.sum())
    g_len, g_rew = n * steps / max(tot_d, 1), tot_r / steps
    eng.close()
    assert abs(g_len - o_len) / o_len < 0.15, (g_len, o_len)
    assert abs(g_rew - o_rew) < 0.15 * max(1.0, abs(o_rew)), (g_rew, o_rew)


def test_pydart2_shaped_views_match_the_oracle(models):
    """env.dart_world / env.robot_skeleton answer the pydart2 calls the reference's env classes make
    (q, dq, ndofs, set_positions / set_velocities / set_forces, q_lower, bodynodes[i].com() /
    to_world() / com_spatial_velocity(), world.dt / step() / skeletons / collision_result.contacts) with the
    oracle's values for the same state."""
    from oracle import oracle as orc
    env_id = "DartWalker2d-v1"
    spec = SPECS[env_id]
    env = _make(env_id, seed=0)
    env.reset()
    rs = env.robot_skeleton
    assert rs.ndofs == 9 and rs.q.shape == (9,) and rs.dq.shape == (9,)
    assert np.array_equal(rs.q_lower, models[env_id].q_lower()) and env.dart_world.dt == spec.dt
    assert env.dart_world.skeletons[-1] is rs
    rng = np.random.RandomState(0)
    q = models[env_id].q_init() + rng.uniform(-0.2, 0.2, 9); dq = rng.uniform(-1, 1, 9)
    q[1] += 0.3                                   # lifted off the ground: a contact-free comparison step
    rs.set_positions(q); rs.set_velocities(dq)
    assert np.allclose(rs.q, q, atol=1e-6) and np.allclose(rs.dq, dq, atol=1e-6)
    w = orc.OracleWorld(models[env_id])
    w.set_state(rs.q, rs.dq)
    loc = np.array([0.1, -0.2, 0.0])
    for i in (2, 5, 8):
        T = w.body_transform(i)
        assert np.allclose(rs.bodynodes[i].com(), w.body_com(i), atol=1e-9)
        assert np.allclose(rs.bodynodes[i].to_world(loc), T[:3, :3] @ loc + T[:3, 3], atol=1e-9)
        assert np.allclose(rs.bodynodes[i].T, T, atol=1e-9)
        assert np.allclose(rs.bodynodes[i].com_spatial_velocity(), w.body_com_spatial_velocity(i), atol=1e-9)
    # set_forces + world.step() == one DART step of the oracle (fp32 engine vs fp64 oracle)
    tau = np.zeros(9); tau[3:] = rng.uniform(-20, 20, 6)
    rs.set_forces(tau); env.dart_world.step()
    w.set_forces(tau); w.step()
    oq, odq = w.get_state()
    assert np.allclose(rs.dq, odq, rtol=2e-4, atol=2e-4) and np.allclose(rs.q, oq, atol=1e-5)
    # falling onto the ground produces contacts with upward force
    for _ in range(300):
        env.dart_world.step()
    cs = env.dart_world.collision_result.contacts
    assert len(cs) >= 1 and all(c.force[1] >= -1e-3 for c in cs)
    # batched env: leading [N] axis
    benv = _make(env_id, num_envs=4, seed=0, output="numpy")
    benv.reset()
    assert benv.robot_skeleton.q.shape == (4, 9) and benv.robot_skeleton.bodynodes[2].com().shape == (4, 3)
    env.close(); benv.close()


def test_step_is_cuda_graph_capturable(models):
    """dartb_step enqueues exactly one kernel on the caller's stream and never synchronises, so a caller can
    capture it in a CUDA graph (e.g. together with its policy network) and replay it: bit-identical to eager."""
    env_id, n = "DartHopper-v1", 512
    dev = torch.device("cuda", 0)
    gen = torch.Generator(device=dev); gen.manual_seed(2)
    acts = [torch.rand((n, 3), generator=gen, device=dev) * 2 - 1 for _ in range(6)]

    def fresh():
        eng = _engine(models, env_id, n, seed=8)
        obs = eng.reset()
        rew = torch.empty((n,), dtype=torch.float32, device=dev); done = torch.empty((n,), dtype=torch.uint8, device=dev)
        return eng, obs, rew, done

    eng, obs, rew, done = fresh()
    for a in acts:
        eng.step(a, obs, rew, done, True)
    q_ref, dq_ref = eng.get_state()
    obs_ref = obs.clone()
    eng.close()

    eng, obs, rew, done = fresh()
    a_static = torch.zeros((n, 3), dtype=torch.float32, device=dev)
    a_static.copy_(acts[0])
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        eng.step(a_static, obs, rew, done, True)          # warm-up on the side stream (step 0)
    torch.cuda.current_stream().wait_stream(s)
    g = torch.cuda.CUDAGraph()
    a_static.copy_(acts[1])
    with torch.cuda.graph(g):
        eng.step(a_static, obs, rew, done, True)          # captured, not executed
    for a in acts[1:]:
        a_static.copy_(a)
        g.replay()
    torch.cuda.synchronize()
    q, dq = eng.get_state()
    assert torch.equal(q, q_ref) and torch.equal(dq, dq_ref) and torch.equal(obs, obs_ref)
    eng.close()
