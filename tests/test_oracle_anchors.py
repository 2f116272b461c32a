"""Anchors for the physics oracle that do NOT share code with it (round 2; the physics stays "parity unpinned" until
oracle/replay_with_pydart2.py runs against real pydart2, these narrow what could still be wrong):

 1. Euler-Lagrange equations by automatic differentiation: L = T - V built from the skeleton GEOMETRY only (torch fp64
    forward kinematics of dart_env_b200/kinematics.py: positions and orientations as functions of q), M(q) and the
    Coriolis / gravity bias from autograd derivatives of it.  The oracle's articulated-body recursion (a force-based
    algorithm in C) must return the same accelerations, including DART's implicit joint damping and springs
    ((M + dt D + dt^2 K) ddq = tau - c - K (q - rest + dt dq) - D dq).  Cart-pole, double pendulum, reacher and the
    airborne hopper / walker / cheetah / snake.
 2. Newton's second law for the whole skeleton: over one DART step the change of total linear momentum equals gravity
    plus the reported contact forces times dt, on every contact-rich golden sample (joint limits, damping and springs
    are internal forces).  Ties the LCP impulses, M^-1 J^T and the contact read-back to the body inertias.
 3. A brute-force boxed-LCP solver (enumerates every active set; numpy) on the LCPs the oracle assembled for the golden
    sub-steps, two-stage friction bounds included: the oracle's Dantzig pivoting must return the same solution.
"""
import itertools
import os

import numpy as np
import pytest
import torch

from dart_env_b200.kinematics import body_transforms
from dart_env_b200.skel import JOINT_REVOLUTE, load_model
from dart_env_b200.tasks import SPECS
from oracle import oracle as orc

GOLD = os.path.join(os.path.dirname(__file__), "golden")
FILES = {"DartHopper-v1": "hopper.npz", "DartWalker2d-v1": "walker2d.npz", "DartHalfCheetah-v1": "halfcheetah.npz"}
CONTACT_FREE = {"cartpole": ("cartpole.skel", 0.02), "swingup": ("cartpole_swingup.skel", 0.01),
                "double_pendulum": ("inverted_double_pendulum.skel", 0.01), "reacher2d": ("reacher2d.skel", 0.01)}


# ------------------------------------------------------------------ 1. Lagrangian by autodiff
def _mass_matrix(model, q):
    """M(q) = sum_b m J_v^T J_v + J_w^T (R I_c R^T) J_w with Jacobians taken by autograd of the forward kinematics."""
    nd = model.n_dofs

    def com_and_rot(qq):
        R, p = body_transforms(model, qq[None, :])
        loc = torch.tensor(np.array([b.com for b in model.bodies]), dtype=torch.float64)
        return ((R[0] @ loc[:, :, None]).squeeze(-1) + p[0]), R[0]

    Jc = torch.autograd.functional.jacobian(lambda qq: com_and_rot(qq)[0], q, create_graph=True)   # [nb,3,nd]
    JR = torch.autograd.functional.jacobian(lambda qq: com_and_rot(qq)[1], q, create_graph=True)   # [nb,3,3,nd]
    _, R = com_and_rot(q)
    M = torch.zeros((nd, nd), dtype=torch.float64)
    for i, b in enumerate(model.bodies):
        Jv = Jc[i]                                                       # [3,nd]
        W = torch.einsum("abk,cb->ack", JR[i], R[i])                     # (dR/dq_k) R^T = [w_k]x
        Jw = torch.stack([W[2, 1], W[0, 2], W[1, 0]], 0)                 # vee -> [3,nd]
        Ic = R[i] @ torch.tensor(np.asarray(b.inertia, dtype=np.float64).reshape(3, 3)) @ R[i].T
        M = M + b.mass * Jv.T @ Jv + Jw.T @ Ic @ Jw
    return M


def _potential(model, q):
    R, p = body_transforms(model, q[None, :])
    loc = torch.tensor(np.array([b.com for b in model.bodies]), dtype=torch.float64)
    com = (R[0] @ loc[:, :, None]).squeeze(-1) + p[0]
    g = torch.tensor(model.gravity, dtype=torch.float64)
    m = torch.tensor([b.mass for b in model.bodies], dtype=torch.float64)
    return -(m[:, None] * com * g[None, :]).sum()


def _lagrange_ddq(model, q, dq, tau):
    q = torch.tensor(q, dtype=torch.float64)
    dq_t = torch.tensor(dq, dtype=torch.float64)
    M = _mass_matrix(model, q)
    dM = torch.autograd.functional.jacobian(lambda qq: _mass_matrix(model, qq), q)                 # [nd,nd,nd] dM_ij/dq_k
    dV = torch.autograd.functional.jacobian(lambda qq: _potential(model, qq), q)
    c = torch.einsum("ijk,j,k->i", dM, dq_t, dq_t) - 0.5 * torch.einsum("jki,j,k->i", dM, dq_t, dq_t) + dV
    dt = model.dt
    dofb = [model.bodies[bi] for bi in model.dof_bodies()]
    D = torch.tensor([b.damping for b in dofb], dtype=torch.float64)
    K = torch.tensor([b.spring_k for b in dofb], dtype=torch.float64)
    rest = torch.tensor([b.spring_rest for b in dofb], dtype=torch.float64)
    rhs = torch.tensor(tau, dtype=torch.float64) - c - K * (q - rest + dt * dq_t) - D * dq_t
    A = M.detach() + torch.diag(dt * D + dt * dt * K)
    return torch.linalg.solve(A, rhs.detach()).numpy(), M.detach().numpy()


def _check_model(model, rng, n_samples, airborne):
    nd = model.n_dofs
    w = orc.OracleWorld(model)
    worst = 0.0
    for _ in range(n_samples):
        q = np.array(model.q_init(), dtype=float) + rng.uniform(-0.6, 0.6, nd)
        if airborne:
            q[1] += 2.0                      # far above the ground: no contact rows
        for d, bi in enumerate(model.dof_bodies()):   # inside the joint limits: no limit rows
            b = model.bodies[bi]
            if b.limit_enforced:
                q[d] = np.clip(q[d], b.q_lo + 0.05, b.q_hi - 0.05)
        dq = rng.uniform(-3, 3, nd)
        tau = rng.uniform(-20, 20, nd)
        w.set_state(q, dq)
        w.set_forces(tau)
        ddq_aba = w.forward_dynamics()
        ddq_lag, M = _lagrange_ddq(model, q, dq, tau)
        assert np.allclose(M, w.mass_matrix(), rtol=1e-9, atol=1e-10)
        worst = max(worst, float(np.abs(ddq_aba - ddq_lag).max() / (1 + np.abs(ddq_lag).max())))
    return worst


@pytest.mark.parametrize("name", list(CONTACT_FREE))
def test_aba_equals_autodiff_lagrangian_contact_free(name):
    skel, dt = CONTACT_FREE[name]
    m = load_model(skel, dt)
    worst = _check_model(m, np.random.default_rng(1), 3, airborne=False)
    assert worst < 1e-9, worst


@pytest.mark.parametrize("env_id", list(SPECS))
def test_aba_equals_autodiff_lagrangian_locomotion(models, env_id):
    worst = _check_model(models[env_id], np.random.default_rng(2), 1, airborne=True)
    assert worst < 1e-9, worst


def test_set_mass_equals_lagrangian_with_that_mass_and_the_original_moment(models):
    """bodynode.set_mass (snake_7link.py:117; DART 6 Inertia::setMass: the mass changes, the moment of inertia about the COM
    stays): the oracle world after set_mass must move like the autodiff Lagrangian of a skeleton with those masses and
    the ORIGINAL moments — the semantics the per-world tables of dartb_set_body_params are lowered with."""
    import copy
    model = models["DartWalker2d-v1"]
    rng = np.random.default_rng(4)
    changed = copy.deepcopy(model)
    w = orc.OracleWorld(model)
    for i, b in enumerate(changed.bodies):
        if b.mass > 0:
            b.mass = float(b.mass + rng.uniform(-1.5, 1.5))
            w.set_mass(i, b.mass)
    nd = model.n_dofs
    q = np.array(model.q_init(), dtype=float) + rng.uniform(-0.4, 0.4, nd)
    q[1] += 2.0
    for d, bi in enumerate(model.dof_bodies()):
        b = model.bodies[bi]
        if b.limit_enforced:
            q[d] = np.clip(q[d], b.q_lo + 0.05, b.q_hi - 0.05)
    dq, tau = rng.uniform(-3, 3, nd), rng.uniform(-20, 20, nd)
    w.set_state(q, dq)
    w.set_forces(tau)
    ddq_lag, M = _lagrange_ddq(changed, q, dq, tau)
    assert np.allclose(M, w.mass_matrix(), rtol=1e-9, atol=1e-10)
    assert np.abs(w.forward_dynamics() - ddq_lag).max() / (1 + np.abs(ddq_lag).max()) < 1e-9
    # and it is not the original skeleton's motion
    w0 = orc.OracleWorld(model)
    w0.set_state(q, dq); w0.set_forces(tau)
    assert np.abs(w0.forward_dynamics() - ddq_lag).max() > 1e-2


# ------------------------------------------------------------------ 2. momentum balance of the contact step
def _total_momentum(w, model):
    P = np.zeros(3)
    for i, b in enumerate(model.bodies):
        T = w.body_transform(i)                       # 3x4
        v_body = w.body_com_spatial_velocity(i)[3:]   # COM velocity, body coordinates
        P += b.mass * (T[:3, :3] @ v_body)
    return P


@pytest.mark.parametrize("env_id", list(FILES))
def test_contact_step_obeys_newtons_second_law(models, env_id):
    """P = A(q) dq is the total linear momentum.  DART's step updates dq at FIXED q0 (positions integrate afterwards), so
    A(q0) (dq' - dq) / dt = m g + sum(contact forces) - (dA/dt) dq, the last term being the momentum's velocity-product
    part, taken here by a central difference of P along dq.  Everything else that acts (joint limits, damping, springs,
    actuator torques) is internal to the skeleton."""
    m = models[env_id]
    g = np.load(os.path.join(GOLD, FILES[env_id]))
    w, w2 = orc.OracleWorld(m), orc.OracleWorld(m)
    mtot = sum(b.mass for b in m.bodies)
    grav = np.array(m.gravity)
    idx = np.where((g["sub_ncontact"] > 0) & (np.abs(g["sub_fext"]).reshape(len(g["sub_q"]), -1).max(1) == 0))[0]
    assert len(idx) > 50
    worst, eps = 0.0, 1e-6

    def P(q, dq):
        w2.set_state(q, dq)
        return _total_momentum(w2, m)

    for i in idx:
        q0, dq0 = g["sub_q"][i], g["sub_dq"][i]
        assert np.all(g["sub_tau"][i][:3] == 0)      # generalized forces on the actuated joints only: internal
        w.set_state(q0, dq0)
        w.set_forces(g["sub_tau"][i])
        w.step()
        if w.lcp_failed():
            continue
        dq1 = w.get_state()[1]
        F = np.sum([c["force"] for c in w.contacts()], axis=0)
        Adot_dq = (P(q0 + eps * dq0, dq0) - P(q0 - eps * dq0, dq0)) / (2 * eps)
        resid = (P(q0, dq1) - P(q0, dq0)) / m.dt - (mtot * grav + F) + Adot_dq
        worst = max(worst, float(np.abs(resid).max() / (mtot * 9.81 + np.abs(F).max())))
    assert worst < 1e-6, worst


# ------------------------------------------------------------------ 3. brute-force boxed LCP
def _solve_fixed(A, b, lo, hi, state):
    """x for a given active set: state[i] in {0 free, 1 at lo, 2 at hi}; returns (x, w) or None if singular"""
    n = len(b)
    x = np.where(state == 1, lo, np.where(state == 2, hi, 0.0))
    F = np.where(state == 0)[0]
    if len(F):
        rhs = b[F] - A[np.ix_(F, np.where(state != 0)[0])] @ x[state != 0]
        try:
            x[F] = np.linalg.solve(A[np.ix_(F, F)], rhs)
        except np.linalg.LinAlgError:
            return None
    return x, A @ x - b


def _enumerate_lcp(A, b, lo, hi, rows, tol=1e-9):
    """all solutions of the boxed LCP restricted to `rows` (other rows held at x = 0)"""
    sols = []
    A_, b_, lo_, hi_ = A[np.ix_(rows, rows)], b[rows], lo[rows], hi[rows]
    choices = [[0] + ([1] if np.isfinite(lo_[i]) else []) + ([2] if np.isfinite(hi_[i]) else []) for i in range(len(rows))]
    for st in itertools.product(*choices):
        st = np.array(st)
        r = _solve_fixed(A_, b_, lo_, hi_, st)
        if r is None:
            continue
        x, wv = r
        scale = 1 + np.abs(x).max() + np.abs(b_).max()
        ok = np.all(x >= lo_ - tol * scale) and np.all(x <= hi_ + tol * scale)
        ok = ok and np.all(wv[st == 1] >= -tol * scale) and np.all(wv[st == 2] <= tol * scale)
        if ok:
            sols.append(x)
    return sols


def _two_stage_bruteforce(A, b, lo, hi, fidx, mu):
    """ODE dSolveLCP semantics.  Stage 1: the non-friction rows alone (friction rows at 0).  Stage 2: every row, with the
    friction bounds fixed at +-|mu x_n| from stage 1.  `lo` / `hi` are what the oracle reports AFTER its solve, i.e. the
    friction rows already carry their fixed stage-2 bounds: stage 1 is enumerated independently and must reproduce them."""
    n = len(b)
    nonf = [i for i in range(n) if fidx[i] < 0]
    s1 = _enumerate_lcp(A, b, lo, hi, nonf)
    assert len(s1) >= 1
    x1 = np.zeros(n)
    x1[nonf] = s1[0]
    for i in range(n):
        if fidx[i] >= 0:
            assert abs(hi[i] - abs(mu[i] * x1[fidx[i]])) <= 1e-7 * (1 + abs(hi[i])) and lo[i] == -hi[i], (i, hi[i], mu[i] * x1[fidx[i]])
    s2 = _enumerate_lcp(A, b, lo, hi, list(range(n)))
    assert len(s2) >= 1
    return s2[0], len(s1), len(s2)


@pytest.mark.parametrize("env_id", list(FILES))
def test_dantzig_equals_bruteforce_enumeration_on_golden_lcps(models, env_id):
    m = models[env_id]
    g = np.load(os.path.join(GOLD, FILES[env_id]))
    w = orc.OracleWorld(m)
    done, worst = 0, 0.0
    for i in np.where(g["sub_lcp_rows"] > 0)[0]:
        w.set_state(g["sub_q"][i], g["sub_dq"][i])
        w.set_forces(g["sub_tau"][i])
        w.step()
        L = w.lcp()
        A, x, b, lo, hi, fi = L["A"], L["x"], L["b"], L["lo"], L["hi"], L["findex"]
        live = [k for k in range(len(b)) if A[k, k] > 1e-12]     # inert rows (zero Jacobian: out-of-plane tangents) carry x = 0
        if len(live) == 0 or len(live) > 8 or w.lcp_failed():
            continue
        # friction coefficient of each friction row: rows (3k, 3k+1, 3k+2) belong to contact k; mu = min(body, ground = 1)
        cs = w.contacts()
        mu_row = np.zeros(len(b))
        for k, c in enumerate(cs):
            mu_row[3 * k + 1] = mu_row[3 * k + 2] = min(m.bodies[c["body"]].friction_coeff, 1.0)
        remap = {k: j for j, k in enumerate(live)}
        fi2 = np.array([remap.get(fi[k], -1) if fi[k] >= 0 else -1 for k in live])
        Al = A[np.ix_(live, live)]
        xb, n1, n2 = _two_stage_bruteforce(Al, b[live], lo[live], hi[live], fi2, mu_row[live])
        err = np.abs(Al @ (xb - x[live])).max() / (1 + np.abs(b[live]).max())
        worst = max(worst, float(err))
        done += 1
        if done >= 60:
            break
    assert done >= 20 and worst < 1e-7, (done, worst)
