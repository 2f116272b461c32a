"""Self-checks of the CPU oracle (no external oracle exists: SURVEY.md §7 step 2).

ABA vs an independent mass-matrix solve, energy conservation, closed forms, LCP residuals,
replay determinism."""
import numpy as np
import pytest

from dart_env_b200.skel import (JOINT_PRISMATIC, JOINT_REVOLUTE, Body, Model, Shape, SHAPE_BOX,
                                SHAPE_CAPSULE, make_transform)
from dart_env_b200.tasks import SPECS
from oracle import oracle as orc


def _random_state(m, rng, scale_q=0.5, scale_v=2.0, lift=1.0):
    nd = m.n_dofs
    q = rng.uniform(-scale_q, scale_q, nd)
    q[1] += lift  # keep it airborne
    dq = rng.uniform(-scale_v, scale_v, nd)
    return q, dq


@pytest.mark.parametrize("env_id", list(SPECS))
def test_aba_matches_mass_matrix_solve(models, env_id):
    """(M + dt*D + dt^2*K) ddq = tau - c - d*dq - k*(q - rest + dt*dq): ABA with implicit joint
    damping/spring (Appendix B.4) equals the joint-space solve with an independently built M."""
    m = models[env_id]
    rng = np.random.default_rng(0)
    w = orc.OracleWorld(m)
    nd, dt = m.n_dofs, m.dt
    dbs = m.dof_bodies()
    d = np.array([m.bodies[i].damping for i in dbs])
    k = np.array([m.bodies[i].spring_k for i in dbs])
    for trial in range(5):
        q, dq = _random_state(m, rng)
        tau = rng.uniform(-50, 50, nd)
        w.set_state(q, dq)
        M = w.mass_matrix()
        assert np.allclose(M, M.T, atol=1e-12)
        assert np.all(np.linalg.eigvalsh(M) > 0)
        # bias c(q,dq) from ABA itself with damping-free algebra: c = M*0 ... use tau=0, undamped identity
        w.set_forces(np.zeros(nd))
        dd0 = w.forward_dynamics()
        # (M+dtD+dt2K) dd0 = -c - d dq - k(q + dt dq)  =>  c
        Mi = M + np.diag(dt * d + dt * dt * k)
        c = -(Mi @ dd0) - d * dq - k * (q + dt * dq)
        w.set_forces(tau)
        dd = w.forward_dynamics()
        rhs = tau - c - d * dq - k * (q + dt * dq)
        assert np.allclose(Mi @ dd, rhs, rtol=1e-9, atol=1e-8)
        # and linearity in tau: dd - dd0 = Mi^-1 tau  (independent of the bias)
        assert np.allclose(dd - dd0, np.linalg.solve(Mi, tau), rtol=1e-9, atol=1e-9)


def _pendulum(damping=0.0, gravity=(0, -9.81, 0), length=1.0, mass=2.0):
    I = np.diag([0.1, 0.2, 0.3])
    b = Body(name="bob", parent=-1, joint_name="j", joint_type=JOINT_REVOLUTE, dof=0,
             T_parent_joint=np.eye(4), T_child_joint=np.eye(4), axis=np.array([0.0, 0.0, 1.0]),
             damping=damping, mass=mass, com=np.array([0.0, -length, 0.0]), inertia=I)
    return Model("pend", 0.001, np.array(gravity, dtype=float), [b], [], [])


def test_pendulum_small_angle_period():
    L, mass = 1.0, 2.0
    m = _pendulum(length=L, mass=mass)
    w = orc.OracleWorld(m)
    Izz = 0.3 + mass * L * L
    omega = np.sqrt(mass * 9.81 * L / Izz)
    th0 = 1e-3
    w.set_state([th0], [0.0])
    n = int(round(2 * np.pi / omega / m.dt))
    for _ in range(n):
        w.step()
    q, dq = w.get_state()
    assert abs(q[0] - th0) < 2e-5 * th0 * 100  # one period later back at th0 (symplectic Euler)


@pytest.mark.parametrize("env_id", ["DartHopper-v1", "DartWalker2d-v1", "DartSnake7Link-v1"])
def test_energy_conserved_without_damping_or_contact(models, env_id):
    import copy
    m = copy.deepcopy(models[env_id])
    for b in m.bodies:
        b.damping = 0.0
        b.spring_k = 0.0
        b.limit_enforced = False
    m.ground = []
    m.dt = 1e-4
    w = orc.OracleWorld(m)
    rng = np.random.default_rng(1)
    q, dq = _random_state(m, rng, scale_v=1.0)
    w.set_state(q, dq)
    e0 = w.energy()
    for _ in range(2000):
        w.step()
    e1 = w.energy()
    assert abs(e1 - e0) < 2e-3 * max(1.0, abs(e0))


def test_free_fall_closed_form(models):
    m = models["DartHopper-v1"]
    w = orc.OracleWorld(m)
    w.reset()
    n = 20  # airborne: foot bottom starts 4 cm above the ground
    for _ in range(n):
        w.step()
    q, dq = w.get_state()
    assert abs(dq[1] - (-9.81 * n * m.dt)) < 1e-9
    # symplectic Euler: y_n = -g dt^2 n(n+1)/2
    assert abs(q[1] - (-9.81 * m.dt ** 2 * n * (n + 1) / 2)) < 1e-9
    assert np.allclose(np.delete(q, 1), 0, atol=1e-9)


def _check_lcp(A, x, b, lo, hi, tol=1e-8):
    w = A @ x - b
    for i in range(len(x)):
        assert lo[i] - tol <= x[i] <= hi[i] + tol
        if x[i] > lo[i] + tol and x[i] < hi[i] - tol:
            assert abs(w[i]) < tol * (1 + np.abs(A[i]).sum())
        elif abs(x[i] - lo[i]) <= tol and not abs(x[i] - hi[i]) <= tol:
            assert w[i] > -tol * (1 + np.abs(A[i]).sum())
        elif abs(x[i] - hi[i]) <= tol and not abs(x[i] - lo[i]) <= tol:
            assert w[i] < tol * (1 + np.abs(A[i]).sum())


def test_dantzig_random_boxed_lcp():
    rng = np.random.default_rng(3)
    for trial in range(200):
        n = rng.integers(1, 13)
        G = rng.normal(size=(n, n + 2))
        A = G @ G.T + 1e-3 * np.eye(n)
        b = rng.normal(size=n) * 3
        lo = np.where(rng.random(n) < 0.5, 0.0, -rng.random(n))
        hi = np.where(rng.random(n) < 0.5, np.inf, rng.random(n) + 0.1)
        lo = np.where(rng.random(n) < 0.1, -np.inf, lo)
        x, w, lo2, hi2, fail = orc.solve_lcp_dantzig(A, b, lo, hi, -np.ones(n, dtype=np.int32))
        assert not fail
        _check_lcp(A, x, b, lo, hi)


def test_dantzig_friction_two_stage():
    """ODE semantics: friction bounds = mu * (normal impulse of the frictionless solve)."""
    rng = np.random.default_rng(4)
    for trial in range(100):
        nc = rng.integers(1, 4)
        n = 3 * nc
        G = rng.normal(size=(n, n + 1))
        A = G @ G.T + 1e-2 * np.eye(n)
        b = rng.normal(size=n) * 2
        lo, hi, fi = np.zeros(n), np.zeros(n), -np.ones(n, dtype=np.int32)
        mu = 0.7
        for c in range(nc):
            lo[3 * c], hi[3 * c] = 0, np.inf
            for r in (1, 2):
                lo[3 * c + r], hi[3 * c + r], fi[3 * c + r] = -mu, mu, 3 * c
        x, w, lo2, hi2, fail = orc.solve_lcp_dantzig(A, b, lo, hi, fi)
        assert not fail
        # stage 1: normal rows only
        idx = np.arange(0, n, 3)
        xn, *_ = orc.solve_lcp_dantzig(A[np.ix_(idx, idx)], b[idx], np.zeros(nc), np.full(nc, np.inf),
                                       -np.ones(nc, dtype=np.int32))
        for c in range(nc):
            for r in (1, 2):
                assert abs(hi2[3 * c + r] - mu * xn[c]) < 1e-9
        _check_lcp(A, x, b, lo2, hi2)


def test_pgs_converges_to_dantzig_without_friction():
    rng = np.random.default_rng(5)
    n = 6
    G = rng.normal(size=(n, n + 3))
    A = G @ G.T + 0.5 * np.eye(n)
    b = rng.normal(size=n)
    lo, hi, fi = np.zeros(n), np.full(n, np.inf), -np.ones(n, dtype=np.int32)
    xd, *_ = orc.solve_lcp_dantzig(A, b, lo, hi, fi)
    xp = orc.solve_lcp_pgs(A, b, lo, hi, fi, 2000)
    assert np.allclose(xd, xp, atol=1e-8)


@pytest.mark.parametrize("env_id", list(SPECS))
def test_step_lcp_solution_is_valid_and_deterministic(models, env_id):
    m = models[env_id]
    spec = SPECS[env_id]
    rng = np.random.default_rng(7)
    e1, e2 = orc.OracleEnv(m, spec.task, 0, 0), orc.OracleEnv(m, spec.task, 0, 0)
    o1, o2 = e1.reset(), e2.reset()
    assert np.array_equal(o1, o2)
    saw_contact = False
    for t in range(300):
        a = rng.uniform(-1, 1, spec.task.n_act)
        r1 = e1.step(a)
        r2 = e2.step(a)
        assert np.array_equal(r1[0], r2[0]) and r1[1] == r2[1] and r1[2] == r2[2]
        L = e1.world.lcp()
        if len(L["x"]) and e1.world.contacts():
            saw_contact = True
            A = L["A"]
            assert np.allclose(A, A.T, atol=1e-9 * np.abs(A).max())
            live = np.diag(A) > 1e-14
            _check_lcp(A[np.ix_(live, live)], L["x"][live], L["b"][live], L["lo"][live], L["hi"][live], tol=1e-7)
        assert not e1.world.lcp_failed()
        if r1[2]:
            e1.reset(), e2.reset()
    if env_id != "DartSnake7Link-v1":
        assert saw_contact
    else:
        assert not saw_contact  # 1 mm permanent gap, ODE has no margin


def test_capsule_box_single_contact_and_tie_rule(models):
    """ODE dCollideCapsuleBox: one contact; a capsule exactly parallel to the face -> endpoint p1."""
    m = models["DartHopper-v1"]
    w = orc.OracleWorld(m)
    q = np.zeros(6)
    q[1] = -0.05  # foot capsule (r = 0.06, centre y = 0.1) now 1 cm into the ground, exactly flat
    w.set_state(q, np.zeros(6))
    w.step()
    cs = w.contacts()
    assert len(cs) == 1 and cs[0]["body"] == 5
    assert abs(cs[0]["depth"] - 0.01) < 1e-12
    assert np.allclose(cs[0]["normal"], [0, 1, 0], atol=1e-12)
    # p1 = centre + (h/2) * axis, axis = R ez with Ry(pi/2): +x end of the foot
    assert abs(cs[0]["point"][0] - (0.065 + 0.195)) < 1e-9
    assert abs(cs[0]["point"][1] - (-0.005)) < 1e-12  # pl - n (r + d)/2
