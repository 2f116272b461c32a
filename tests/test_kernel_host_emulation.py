"""No-GPU regression test of the KERNEL SOURCE: tools/host_emu compiles the very same device code
(dart_env_b200/csrc/planar_kernels.cuh + lower.h) as plain C++ and steps the golden (q, dq, tau)
triples on the CPU.  The fp64 instantiation must agree with the 3-D fp64 oracle to 1e-9 (the
planar/weld-merged formulation is algebraically the same step); the fp32 instantiation must
stay inside the tolerance the GPU test states.  This is a test tool, not a product fallback:
libdartb.so contains no host path."""
import os

import numpy as np
import pytest

from dart_env_b200.tasks import SPECS
from tools.host_emu import emu

GOLD = os.path.join(os.path.dirname(__file__), "golden")
FILES = {"DartHopper-v1": "hopper.npz", "DartWalker2d-v1": "walker2d.npz",
         "DartHalfCheetah-v1": "halfcheetah.npz", "DartSnake7Link-v1": "snake7link.npz"}


DEEP, TOL_SUB_DQ, TOL_SUB_Q, TOL_SUB_DQ_DEEP, TOL_SUB_Q_DEEP = 0.03, 2e-4, 1e-5, 2e-2, 2e-4   # == tests/test_gpu_parity.py (stated there)
VARIANTS = [0, 1]  # 0 = unrolled per-topology kernel, 1 = loop / topology-generic kernel
# (2 = lane-cooperative kernel: its own tests below, under the SIMT emulator of tools/host_emu/simt.h)


def _run(models, env_id, f64, **kw):
    g = np.load(os.path.join(GOLD, FILES[env_id]))
    mc = 8
    out = emu.substep(models[env_id], SPECS[env_id].task, g["sub_q"], g["sub_dq"], g["sub_tau"], g["sub_fext"],
                      f64=f64, maxc=mc, **kw)
    return g, out


@pytest.mark.parametrize("variant", VARIANTS)
@pytest.mark.parametrize("env_id", list(SPECS))
def test_kernel_source_fp64_equals_oracle(models, env_id, variant):
    g, (q2, dq2, cnt, body, data) = _run(models, env_id, True, variant=variant)
    assert np.allclose(q2, g["sub_q2"], rtol=1e-9, atol=1e-10)
    assert np.allclose(dq2, g["sub_dq2"], rtol=1e-8, atol=1e-8)
    safe = g["sub_contact_margin"] > 1e-9
    assert np.array_equal(cnt[safe], g["sub_ncontact"][safe])
    tie_ok = safe & (g["sub_tie_margin"] > 1e-9)
    assert np.array_equal(body[tie_ok], g["sub_contact_body"][tie_ok])
    assert np.allclose(data[tie_ok][..., :7], g["sub_contact_data"][tie_ok][..., :7], atol=1e-6)
    assert np.allclose(data[tie_ok][..., 7:], g["sub_contact_data"][tie_ok][..., 7:], rtol=1e-4, atol=1e-2)


@pytest.mark.parametrize("variant", VARIANTS)
@pytest.mark.parametrize("env_id", list(SPECS))
def test_kernel_source_fp32_within_tolerance(models, env_id, variant):
    g, (q2, dq2, cnt, body, data) = _run(models, env_id, False, variant=variant)
    safe = (g["sub_contact_margin"] > 1e-4) & (g["sub_limit_margin"] > 1e-4) & (g["sub_tie_margin"] > 1e-4)
    assert np.array_equal(cnt[safe], g["sub_ncontact"][safe])
    assert np.array_equal(body[safe], g["sub_contact_body"][safe])
    ev = (np.abs(dq2 - g["sub_dq2"]) / (1 + np.abs(g["sub_dq2"]))).max(1)
    eq = (np.abs(q2 - g["sub_q2"]) / (1 + np.abs(g["sub_q2"]))).max(1)
    deep = g["sub_contact_data"][:, :, 6].max(1) > DEEP     # tagged class: capsule > 3 cm inside the ground
    # the stated fp32 tolerances of tests/test_gpu_parity.py, as MAXIMA over the untagged samples
    assert ev[safe & ~deep].max() < TOL_SUB_DQ and eq[safe & ~deep].max() < TOL_SUB_Q
    assert not (safe & deep).any() or (ev[safe & deep].max() < TOL_SUB_DQ_DEEP and eq[safe & deep].max() < TOL_SUB_Q_DEEP)


def test_kernel_source_pgs_equals_oracle_pgs(models):
    from oracle import oracle as orc
    env_id = "DartWalker2d-v1"
    g = np.load(os.path.join(GOLD, FILES[env_id]))
    idx = np.where(g["sub_ncontact"] > 0)[0][:30]
    for iters in (1, 5, 30):
        w = orc.OracleWorld(models[env_id])
        w.set_option(1, 1); w.set_option(2, iters)
        ref = []
        for i in idx:
            w.set_state(g["sub_q"][i], g["sub_dq"][i]); w.set_forces(g["sub_tau"][i]); w.step()
            ref.append(np.concatenate(w.get_state()))
        q2, dq2, *_ = emu.substep(models[env_id], SPECS[env_id].task, g["sub_q"][idx], g["sub_dq"][idx], g["sub_tau"][idx],
                                  f64=True, lcp_mode=1, pgs_iters=iters)
        assert np.allclose(np.concatenate([q2, dq2], 1), np.array(ref), rtol=1e-8, atol=1e-8)


def test_reset_noise_generator_matches_oracle():
    from oracle import oracle as orc
    L = emu.lib()
    for seed, w, ep, i in [(0, 0, 0, 0), (1234, 1017, 3, 7), (2 ** 40 + 5, 2 ** 33, 9, 17)]:
        assert L.emu_reset_uniform(seed, w, ep, i) == orc.reset_uniform(seed, w, ep, i)


def _random_contact_lcp(rng, nc, nl, mu=1.0, rank_def=False):
    """LCP with the kernel's row layout: (normal, tangent) per contact, then limit rows."""
    n = 2 * nc + nl
    G = rng.normal(size=(n, n + 2 if not rank_def else max(2, n - 2)))
    A = G @ G.T
    A += np.diag(np.diag(A)) * 1e-5 + 1e-9 * np.eye(n)   # the CFM the kernel applies
    b = rng.normal(size=n) * 2
    lo, hi, fi = np.zeros(n), np.full(n, np.inf), -np.ones(n, dtype=np.int32)
    for c in range(nc):
        lo[2 * c + 1], hi[2 * c + 1], fi[2 * c + 1] = -mu, mu, 2 * c
    for k in range(nl):
        if rng.random() < 0.5:
            lo[2 * nc + k], hi[2 * nc + k] = -np.inf, 0.0
    return A, b, lo, hi, fi


@pytest.mark.parametrize("mode,maxn", [(0, 20), (4, 4), (5, 6), (1, 8), (2, 20), (3, 20), (6, 4), (7, 6), (8, 8), (9, 20)])
def test_kernel_lcp_solvers_equal_oracle_dantzig(mode, maxn):
    """Every LCP code path of the kernel (register block pivoting <4>/<6>/<8>, thread-local block
    pivoting, Dantzig, and the dispatch) returns the oracle's Dantzig solution: the two-stage
    boxed LCP has a unique solution for positive-definite A."""
    from oracle import oracle as orc
    rng = np.random.default_rng(100 + mode)
    worst = 0.0
    for trial in range(300):
        nc = int(rng.integers(0, 5))
        nl = int(rng.integers(0 if nc else 1, 4))
        if 2 * nc + nl > maxn:
            continue
        A, b, lo, hi, fi = _random_contact_lcp(rng, nc, nl, mu=float(rng.choice([1.0, 0.5, 0.05])), rank_def=(trial % 7 == 0))
        xr, *_rest = orc.solve_lcp_dantzig(A, b, lo.copy(), hi.copy(), fi)
        x, rc = emu.lcp(A, b, lo, hi, fi, mode=mode, f64=True)
        assert rc == 0
        # compare the physically meaningful quantity A x (x itself is not unique when A is singular up to CFM)
        err = np.abs(A @ (x - xr)).max() / (1 + np.abs(A @ xr).max())
        worst = max(worst, err)
    assert worst < 1e-6, worst


# ------------------------------------------------------------------ lane-cooperative kernel (planar_coop.cuh)
def _run_coop(models, env_id, f64, idx=None, **kw):
    g = np.load(os.path.join(GOLD, FILES[env_id]))
    nf = np.abs(g["sub_fext"]).reshape(len(g["sub_q"]), -1).max(1) == 0   # the cooperative kernel takes no external forces
    sel = np.where(nf)[0] if idx is None else np.asarray(idx)
    out = emu.substep(models[env_id], SPECS[env_id].task, g["sub_q"][sel], g["sub_dq"][sel], g["sub_tau"][sel], None,
                      f64=f64, maxc=8, variant=2, **kw)
    return g, sel, out


@pytest.mark.parametrize("env_id", list(SPECS))
def test_cooperative_kernel_fp64_equals_oracle(models, env_id):
    """G lanes per world (prefix-sum kinematics, CRBA + sparse LTL, distributed LCP) == the 3-D fp64 oracle."""
    g, sel, (q2, dq2, cnt, body, data) = _run_coop(models, env_id, True)
    safe = g["sub_contact_margin"][sel] > 1e-9    # exact-touch samples may flip: a different (but valid) summation order
    assert safe.mean() > 0.95
    assert np.allclose(q2[safe], g["sub_q2"][sel][safe], rtol=1e-9, atol=1e-10)
    assert np.allclose(dq2[safe], g["sub_dq2"][sel][safe], rtol=1e-8, atol=1e-8)
    assert np.array_equal(cnt[safe], g["sub_ncontact"][sel][safe])
    tie_ok = safe & (g["sub_tie_margin"][sel] > 1e-9)
    assert np.array_equal(body[tie_ok], g["sub_contact_body"][sel][tie_ok])
    assert np.allclose(data[tie_ok][..., :7], g["sub_contact_data"][sel][tie_ok][..., :7], atol=1e-6)
    assert np.allclose(data[tie_ok][..., 7:], g["sub_contact_data"][sel][tie_ok][..., 7:], rtol=1e-4, atol=1e-2)


@pytest.mark.parametrize("env_id", list(SPECS))
def test_cooperative_kernel_fp32_within_tolerance(models, env_id):
    """fp32 engine with the fp64 mass-matrix core: TIGHTER than the per-thread fp32 kernel's stated tolerance."""
    g, sel, (q2, dq2, cnt, body, data) = _run_coop(models, env_id, False)
    safe = (g["sub_contact_margin"][sel] > 1e-4) & (g["sub_limit_margin"][sel] > 1e-4) & (g["sub_tie_margin"][sel] > 1e-4)
    assert np.array_equal(cnt[safe], g["sub_ncontact"][sel][safe])
    assert np.array_equal(body[safe], g["sub_contact_body"][sel][safe])
    ev = (np.abs(dq2 - g["sub_dq2"][sel]) / (1 + np.abs(g["sub_dq2"][sel]))).max(1)
    eq = (np.abs(q2 - g["sub_q2"][sel]) / (1 + np.abs(g["sub_q2"][sel]))).max(1)
    deep = g["sub_contact_data"][sel][:, :, 6].max(1) > DEEP
    assert np.median(ev[safe]) < 2e-6
    assert ev[safe & ~deep].max() < TOL_SUB_DQ and eq[safe & ~deep].max() < TOL_SUB_Q
    assert not (safe & deep).any() or (ev[safe & deep].max() < TOL_SUB_DQ_DEEP and eq[safe & deep].max() < TOL_SUB_Q_DEEP)


def test_cooperative_kernel_world_independent_of_warp_neighbours(models):
    """A warp picks ONE LCP column class from its largest row count; a world's result must not depend on the
    class its neighbours put it in (shard == batch, bit for bit)."""
    env_id = "DartWalker2d-v1"
    g = np.load(os.path.join(GOLD, FILES[env_id]))
    nf = np.abs(g["sub_fext"]).reshape(len(g["sub_q"]), -1).max(1) == 0
    rows = g["sub_lcp_rows"]
    small = np.where(nf & (rows >= 1) & (rows <= 4))[0][:6]
    big = np.where(nf & (rows >= 5))[0][:6]
    assert len(small) == 6 and len(big) >= 2
    for f64 in (True, False):
        _, _, (qa, dqa, *_r) = _run_coop(models, env_id, f64, idx=small)                      # warps of small worlds only
        mixed = np.array([v for pair in zip(small, np.resize(big, 6)) for v in pair])        # every warp holds a big one
        _, _, (qm, dqm, *_r) = _run_coop(models, env_id, f64, idx=mixed)
        assert np.array_equal(qa, qm[0::2]) and np.array_equal(dqa, dqm[0::2])


def test_cooperative_kernel_pgs_equals_oracle_pgs(models):
    from oracle import oracle as orc
    env_id = "DartWalker2d-v1"
    g = np.load(os.path.join(GOLD, FILES[env_id]))
    nf = np.abs(g["sub_fext"]).reshape(len(g["sub_q"]), -1).max(1) == 0
    idx = np.where(nf & (g["sub_ncontact"] > 0))[0][:30]
    for iters in (1, 5, 30):
        w = orc.OracleWorld(models[env_id])
        w.set_option(1, 1); w.set_option(2, iters)
        ref = []
        for i in idx:
            w.set_state(g["sub_q"][i], g["sub_dq"][i]); w.set_forces(g["sub_tau"][i]); w.step()
            ref.append(np.concatenate(w.get_state()))
        _, _, (q2, dq2, *_r) = _run_coop(models, env_id, True, idx=idx, lcp_mode=1, pgs_iters=iters)
        assert np.allclose(np.concatenate([q2, dq2], 1), np.array(ref), rtol=1e-8, atol=1e-8)


def test_many_contact_states_large_column_classes(models):
    """A fallen walker (13-16 LCP rows: six capsules on the ground + limits) drives the 16-column class of the
    cooperative LCP and the thread-local pivoting of the per-thread kernel.  States are harvested from oracle
    rollouts started in random tumbled poses; samples where the oracle's own Dantzig gave up are excluded."""
    from oracle import oracle as orc
    env_id = "DartWalker2d-v1"
    m, task = models[env_id], SPECS[env_id].task
    rng = np.random.default_rng(1)
    S = []
    for trial in range(8):
        w = orc.OracleWorld(m)
        q = np.array(m.q_init(), dtype=float); dq = np.zeros_like(q)
        q[2] = rng.uniform(-2.0, 2.0); q[1] += rng.uniform(-0.3, 0.1); q[3:] += rng.uniform(-0.6, 0.6, size=len(q) - 3)
        w.set_state(q, dq)
        for t in range(300):
            tau = np.zeros(len(q)); tau[3:] = rng.uniform(-1, 1, size=len(q) - 3) * 30
            s = w.get_state()
            w.set_forces(tau); w.step()
            n = 2 * len(w.contacts()) + int((w.limit_active() != 0).sum())
            if n >= 13 and not w.lcp_failed() and t % 3 == 0:
                S.append((s[0].copy(), s[1].copy(), tau.copy(), w.get_state()[1].copy()))
    assert len(S) >= 10
    q, dq, tau, ref = (np.array([s[k] for s in S]) for k in range(4))
    for variant in (0, 2, 3):
        _, dq2, *_r = emu.substep(m, task, q, dq, tau, None, f64=True, maxc=8, variant=variant)
        assert np.allclose(dq2, ref, rtol=1e-7, atol=1e-7), variant
        _, dq2, *_r = emu.substep(m, task, q, dq, tau, None, f64=False, maxc=8, variant=variant)
        ev = (np.abs(dq2 - ref) / (1 + np.abs(ref))).max(1)
        # (fp32: A is rank-deficient up to the CFM here, cond ~ 1e5; the cooperative kernel runs these classes in fp64)
        assert np.median(ev) < (2e-5 if variant == 2 else 1e-3) and ev.max() < 0.2, (variant, np.median(ev), ev.max())


# ------------------------------------------------------------------ quad form of the per-thread kernels (substep<..., G = 4>)
QUAD_VARIANTS = {3: 4, 4: 2, 5: 8}   # kernel variant -> lanes per world


def _run_quad(models, env_id, f64, idx=None, variant=3, **kw):
    g = np.load(os.path.join(GOLD, FILES[env_id]))
    nf = np.abs(g["sub_fext"]).reshape(len(g["sub_q"]), -1).max(1) == 0
    sel = np.where(nf)[0] if idx is None else np.asarray(idx)
    out = emu.substep(models[env_id], SPECS[env_id].task, g["sub_q"][sel], g["sub_dq"][sel], g["sub_tau"][sel], None,
                      f64=f64, maxc=8, variant=variant, **kw)
    return g, sel, out


@pytest.mark.parametrize("variant", list(QUAD_VARIANTS))
@pytest.mark.parametrize("env_id", list(SPECS))
def test_quad_kernel_fp64_equals_oracle(models, env_id, variant):
    """four lanes per world: redundant K1-K4, constraint rows dealt out over the lanes, GroupLcp<4> == the 3-D fp64 oracle;
    the emulation also checks that the lanes of a group hold the same bits (NaN marks otherwise)"""
    g, sel, (q2, dq2, cnt, body, data) = _run_quad(models, env_id, True, variant=variant)
    assert not np.isnan(q2).any()
    safe = g["sub_contact_margin"][sel] > 1e-9
    assert np.allclose(q2[safe], g["sub_q2"][sel][safe], rtol=1e-9, atol=1e-10)
    assert np.allclose(dq2[safe], g["sub_dq2"][sel][safe], rtol=1e-8, atol=1e-8)
    assert np.array_equal(cnt[safe], g["sub_ncontact"][sel][safe])
    tie_ok = safe & (g["sub_tie_margin"][sel] > 1e-9)
    assert np.array_equal(body[tie_ok], g["sub_contact_body"][sel][tie_ok])
    assert np.allclose(data[tie_ok][..., :7], g["sub_contact_data"][sel][tie_ok][..., :7], atol=1e-6)
    assert np.allclose(data[tie_ok][..., 7:], g["sub_contact_data"][sel][tie_ok][..., 7:], rtol=1e-4, atol=1e-2)


@pytest.mark.parametrize("variant", list(QUAD_VARIANTS))
@pytest.mark.parametrize("env_id", list(SPECS))
def test_quad_kernel_fp32_within_tolerance(models, env_id, variant):
    g, sel, (q2, dq2, cnt, body, data) = _run_quad(models, env_id, False, variant=variant)
    assert not np.isnan(q2).any()
    safe = (g["sub_contact_margin"][sel] > 1e-4) & (g["sub_limit_margin"][sel] > 1e-4) & (g["sub_tie_margin"][sel] > 1e-4)
    assert np.array_equal(cnt[safe], g["sub_ncontact"][sel][safe])
    assert np.array_equal(body[safe], g["sub_contact_body"][sel][safe])
    ev = (np.abs(dq2 - g["sub_dq2"][sel]) / (1 + np.abs(g["sub_dq2"][sel]))).max(1)
    eq = (np.abs(q2 - g["sub_q2"][sel]) / (1 + np.abs(g["sub_q2"][sel]))).max(1)
    deep = g["sub_contact_data"][sel][:, :, 6].max(1) > DEEP
    assert ev[safe & ~deep].max() < TOL_SUB_DQ and eq[safe & ~deep].max() < TOL_SUB_Q
    assert not (safe & deep).any() or (ev[safe & deep].max() < TOL_SUB_DQ_DEEP and eq[safe & deep].max() < TOL_SUB_Q_DEEP)


def test_quad_kernel_pgs_equals_oracle_pgs(models):
    from oracle import oracle as orc
    env_id = "DartWalker2d-v1"
    g = np.load(os.path.join(GOLD, FILES[env_id]))
    nf = np.abs(g["sub_fext"]).reshape(len(g["sub_q"]), -1).max(1) == 0
    idx = np.where(nf & (g["sub_ncontact"] > 0))[0][:30]
    for iters in (1, 5, 30):
        w = orc.OracleWorld(models[env_id])
        w.set_option(1, 1); w.set_option(2, iters)
        ref = []
        for i in idx:
            w.set_state(g["sub_q"][i], g["sub_dq"][i]); w.set_forces(g["sub_tau"][i]); w.step()
            ref.append(np.concatenate(w.get_state()))
        _, _, (q2, dq2, *_r) = _run_quad(models, env_id, True, idx=idx, lcp_mode=1, pgs_iters=iters)
        assert np.allclose(np.concatenate([q2, dq2], 1), np.array(ref), rtol=1e-8, atol=1e-8)


def test_quad_kernel_world_independent_of_warp_neighbours(models):
    env_id = "DartWalker2d-v1"
    g = np.load(os.path.join(GOLD, FILES[env_id]))
    nf = np.abs(g["sub_fext"]).reshape(len(g["sub_q"]), -1).max(1) == 0
    rows = g["sub_lcp_rows"]
    small = np.where(nf & (rows >= 1) & (rows <= 4))[0][:8]
    big = np.where(nf & (rows >= 9))[0][:8]      # oracle rows (3 per contact): sends the warp to the 8-column class or the per-thread path
    assert len(small) == 8 and len(big) >= 2
    for f64 in (True, False):
        _, _, (qa, dqa, *_r) = _run_quad(models, env_id, f64, idx=small)
        mixed = np.array([v for pair in zip(small, np.resize(big, 8)) for v in pair])
        _, _, (qm, dqm, *_r) = _run_quad(models, env_id, f64, idx=mixed)
        assert np.array_equal(qa, qm[0::2]) and np.array_equal(dqa, dqm[0::2])


# ------------------------------------------------------------------------ per-world dynamics parameters
@pytest.mark.parametrize("env_id", ["DartWalker2d-v1", "DartSnake7Link-v1", "DartHalfCheetah-v1"])
def test_per_world_body_params_equal_oracle_set_mass_set_friction(models, env_id):
    """bodynodes[i].set_mass / set_friction_coeff per world (snake_7link.py:115-120): the loop kernel reading the table
    dartb_set_body_params uploads == one oracle world per sample with orc_set_mass / orc_set_friction applied."""
    from oracle import oracle as orc
    g = np.load(os.path.join(GOLD, FILES[env_id]))
    has_c = np.where(g["sub_ncontact"] > 0)[0][:12]
    idx = np.concatenate([has_c, np.arange(8)]) if len(has_c) else np.arange(16)
    m = models[env_id]
    nb = len(m.bodies)
    rng = np.random.RandomState(5)
    mass0 = np.array([b.mass for b in m.bodies])
    fr0 = np.array([b.friction_coeff for b in m.bodies])
    mass = np.clip(mass0 + rng.uniform(-1.5, 1.5, (len(idx), nb)), 0.0, None) * (mass0 > 0)   # massless helper bodies stay massless
    fric = np.clip(fr0 + rng.uniform(-0.5, 0.5, (len(idx), nb)), 0.0, None)
    ref = []
    for k, i in enumerate(idx):
        w = orc.OracleWorld(m)
        for b in range(nb):
            w.set_mass(b, mass[k, b]); w.set_friction(b, fric[k, b])
        w.set_state(g["sub_q"][i], g["sub_dq"][i]); w.set_forces(g["sub_tau"][i]); w.step()
        ref.append(np.concatenate(w.get_state()))
    ref = np.array(ref)
    q2, dq2, *_ = emu.substep_body_params(m, SPECS[env_id].task, g["sub_q"][idx], g["sub_dq"][idx], g["sub_tau"][idx],
                                          mass=mass, friction=fric, f64=True)
    got = np.concatenate([q2, dq2], 1)
    assert np.allclose(got, ref, rtol=1e-7, atol=1e-7), np.abs(got - ref).max()
    # and the parameters matter: the shared-model step differs
    q0, dq0, *_ = emu.substep(m, SPECS[env_id].task, g["sub_q"][idx], g["sub_dq"][idx], g["sub_tau"][idx], f64=True, variant=1)
    assert np.abs(dq0 - dq2).max() > 1e-3
    # mass only / friction only go through the same table
    q3, dq3, *_ = emu.substep_body_params(m, SPECS[env_id].task, g["sub_q"][idx], g["sub_dq"][idx], g["sub_tau"][idx],
                                          mass=np.tile(mass0, (len(idx), 1)), friction=None, f64=True)
    assert np.allclose(dq3, dq0, rtol=1e-12, atol=1e-12)


def test_branch_free_sincos_accuracy():
    """Num<float>::sincos_ (Cody-Waite reduction by pi/2 + minimax polynomials, no branches) against double precision:
    <= 2 ulp like libm's sincosf over the range joint angles can reach, exact identities at the special points."""
    import ctypes as C
    L = emu.lib()
    rng = np.random.RandomState(0)
    fp = C.POINTER(C.c_float)
    for R in (3.2, 100.0, 1e4, 1e5):
        x = rng.uniform(-R, R, 400000).astype(np.float32)
        s, c = np.empty_like(x), np.empty_like(x)
        L.emu_sincos(len(x), x.ctypes.data_as(fp), s.ctypes.data_as(fp), c.ctypes.data_as(fp))
        xs = x.astype(np.float64)
        rs, rc = np.sin(xs), np.cos(xs)
        ulp_s = np.abs(s - rs) / np.spacing(np.abs(rs).astype(np.float32))
        ulp_c = np.abs(c - rc) / np.spacing(np.abs(rc).astype(np.float32))
        assert ulp_s.max() <= 2.0 and ulp_c.max() <= 2.0, (R, ulp_s.max(), ulp_c.max())
        assert np.abs(s - rs).max() < 1.2e-7 and np.abs(c - rc).max() < 1.2e-7
        assert np.abs(s.astype(np.float64) ** 2 + c.astype(np.float64) ** 2 - 1).max() < 4e-7
    x = np.array([0.0, -0.0, 1e-30, -1e-30, np.pi / 2, -np.pi / 2, np.pi, 2 * np.pi, 1e-4], dtype=np.float32)
    s, c = np.empty_like(x), np.empty_like(x)
    L.emu_sincos(len(x), x.ctypes.data_as(fp), s.ctypes.data_as(fp), c.ctypes.data_as(fp))
    assert s[0] == 0 and c[0] == 1 and s[1] == 0 and c[1] == 1 and s[2] == x[2] and s[3] == x[3] and c[2] == 1
    assert np.allclose(s, np.sin(x.astype(np.float64)), atol=1e-7) and np.allclose(c, np.cos(x.astype(np.float64)), atol=1e-7)
