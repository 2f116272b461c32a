#!/usr/bin/env python
"""Mint the golden vectors in tests/golden/*.npz  (run HERE, in the build container).

The task layer (action clamp/scale, frame-skip loop, obs, reward, done, snake fluid force) is
executed by the reference's OWN env classes, imported unmodified from /root/reference
(gym/envs/dart/{hopper,walker2d,half_cheetah,snake_7link}.py on top of dart_env.py), with the
pydart2 import satisfied by oracle/pydart2_shim (the fp64 CPU oracle).  pydart2/DART themselves
are not installable here, so the PHYSICS in these vectors is the oracle's
restatement ("parity unpinned"); the TASK LAYER is pinned to the reference code.

The format is engine-agnostic ((q, dq, tau/action) -> results) so real pydart2 could replay it.

  python tests/golden/make_golden.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle", "pydart2_shim"), "/root/reference"]

import warnings

import numpy as np

warnings.filterwarnings("ignore")

from dart_env_b200.skel import load_model  # noqa: E402
from dart_env_b200.tasks import SPECS  # noqa: E402
from oracle.oracle import OracleWorld  # noqa: E402

REF_CLASSES = {
    "DartHopper-v1": ("gym.envs.dart.hopper", "DartHopperEnv"),
    "DartWalker2d-v1": ("gym.envs.dart.walker2d", "DartWalker2dEnv"),
    "DartHalfCheetah-v1": ("gym.envs.dart.half_cheetah", "DartHalfCheetahEnv"),
    "DartSnake7Link-v1": ("gym.envs.dart.snake_7link", "DartSnake7LinkEnv"),
}
# SURVEY 8f.1 contact-free envs (host-side task layer in dart_env_b200/envs_contact_free.py)
REF_CLASSES_CF = {
    "DartCartPole-v1": ("gym.envs.dart.cart_pole", "DartCartPoleEnv", "cartpole.skel", 0.02),
    "DartCartPoleSwingUp-v1": ("gym.envs.dart.cartpole_swingup", "DartCartPoleSwingUpEnv", "cartpole_swingup.skel", 0.01),
    "DartDoubleInvertedPendulumEnv-v1": ("gym.envs.dart.inverted_double_pendulum", "DartDoubleInvertedPendulumEnv",
                                         "inverted_double_pendulum.skel", 0.01),
    "DartReacher-v1": ("gym.envs.dart.reacher2d", "DartReacher2dEnv", "reacher2d.skel", 0.01),
}
MAXC = 8


def contact_free_rollouts(env_id, n_steps, seed):
    import importlib
    mod, cls, skel, dt = REF_CLASSES_CF[env_id]
    env = getattr(importlib.import_module(mod), cls)()
    env.seed(seed)
    rng = np.random.RandomState(seed)
    rec = {k: [] for k in ("q", "dq", "action", "obs", "reward", "done", "q2", "dq2", "target")}
    nact = env.action_space.shape[0]
    for scale in (1.0, 0.3, 0.05):
        env.reset()
        for t in range(n_steps):
            a = rng.uniform(-1.2, 1.2, nact) * scale
            rec["target"].append(np.array(getattr(env, "target", np.zeros(3)), dtype=np.float64))
            s0 = env.state_vector()
            ob, r, d, _ = env.step(a)
            s1 = env.state_vector()
            nd = len(s0) // 2
            rec["q"].append(s0[:nd]); rec["dq"].append(s0[nd:]); rec["action"].append(a)
            rec["obs"].append(ob); rec["reward"].append(r); rec["done"].append(d)
            rec["q2"].append(s1[:nd]); rec["dq2"].append(s1[nd:])
            if d:
                env.reset()
    return {"step_" + k: np.array(v) for k, v in rec.items()}


def build_model(env_id):
    spec = SPECS[env_id]
    m = load_model(spec.skel, spec.dt)
    m.enforce_limits()
    if spec.friction_all is not None:
        for b in m.bodies:
            b.friction_coeff = spec.friction_all
    return m


def done_margin(spec, env):
    t = spec.task
    s = env.state_vector()
    mg = [abs(abs(s[2]) - t.ang_max), t.state_bound - np.abs(s[2:]).max()]
    if t.height_body >= 0:
        h = env.robot_skeleton.bodynodes[t.height_body].com()[1]
        mg += [abs(h - t.height_lo), abs(h - t.height_hi)]
    return float(np.min(np.abs(mg)))


def env_rollouts(env_id, n_steps, scales, seed):
    import importlib
    spec = SPECS[env_id]
    mod, cls = REF_CLASSES[env_id]
    env = getattr(importlib.import_module(mod), cls)()
    env.seed(seed)
    rng = np.random.RandomState(seed)
    rec = {k: [] for k in ("q", "dq", "action", "obs", "reward", "done", "q2", "dq2", "margin", "reset_obs", "event_margin")}
    # distance of every DISCRETE decision taken inside an env step (contact on/off per capsule, flat-capsule end tie,
    # joint limit on/off, over all frame_skip DART steps) from its threshold: the fp32 parity tests assert a MAX error
    # over the samples whose decisions are not within rounding of flipping, and tag the rest
    m = build_model(env_id)
    limits = [(d, m.bodies[bi].q_lo, m.bodies[bi].q_hi) for d, bi in enumerate(m.dof_bodies()) if m.bodies[bi].limit_enforced]
    world = env.dart_world
    orig_step = world.step
    ev = {"m": np.inf}

    def step_hook():
        q_pre = world._ow.get_state()[0].copy()
        orig_step()
        gaps, tilts = world._ow.shape_gaps()
        mg = [np.inf]
        if len(gaps):
            mg.append(float(np.min(np.abs(gaps))))
            if (gaps <= 0).any():
                mg.append(float(np.min(tilts[gaps <= 0])))
        for d, lo, hi in limits:
            mg += [abs(q_pre[d] - lo), abs(q_pre[d] - hi)]
        ev["m"] = min(ev["m"], min(mg))

    world.step = step_hook
    for scale in scales:
        ob = env.reset()
        rec["reset_obs"].append(np.concatenate([env.state_vector(), ob]))
        for t in range(n_steps):
            a = rng.uniform(-1.3, 1.3, spec.task.n_act) * scale  # beyond +-1: exercises the clamp
            s0 = env.state_vector()
            ev["m"] = np.inf
            ob, r, d, _ = env.step(a)
            s1 = env.state_vector()
            nd = len(s0) // 2
            rec["q"].append(s0[:nd]); rec["dq"].append(s0[nd:]); rec["action"].append(a)
            rec["obs"].append(ob); rec["reward"].append(r); rec["done"].append(d)
            rec["q2"].append(s1[:nd]); rec["dq2"].append(s1[nd:]); rec["margin"].append(done_margin(spec, env))
            rec["event_margin"].append(ev["m"])
            if d:
                ob = env.reset()
    return {"step_" + k: np.array(v) for k, v in rec.items()}


def substep_record(w, m, q, dq, tau, fext=None):
    nd = m.n_dofs
    w.set_state(q, dq)
    if fext is not None:
        for i in range(m.n_bodies):
            w.add_ext_force(i, fext[i])
    w.set_forces(tau)
    w.step()
    q2, dq2 = w.get_state()
    cs = w.contacts()
    body = -np.ones(MAXC, dtype=np.int32)
    data = np.zeros((MAXC, 10))
    for i, c in enumerate(cs):
        body[i] = c["body"]
        data[i] = np.concatenate([c["point"], c["normal"], [c["depth"]], c["force"]])
    gaps, tilts = w.shape_gaps()
    lim = w.limit_active()
    lim_margin = np.inf
    for d, bi in enumerate(m.dof_bodies()):
        b = m.bodies[bi]
        if b.limit_enforced:
            lim_margin = min(lim_margin, abs(q[d] - b.q_lo), abs(q[d] - b.q_hi))
    contact_margin = float(np.min(np.abs(gaps))) if len(gaps) else np.inf
    touching = gaps <= 0
    tie_margin = float(np.min(tilts[touching])) if touching.any() else np.inf
    return dict(q=q, dq=dq, tau=tau, fext=np.zeros((m.n_bodies, 3)) if fext is None else fext, q2=q2, dq2=dq2,
                ncontact=len(cs), contact_body=body, contact_data=data, limit_active=lim,
                contact_margin=contact_margin, tie_margin=tie_margin, limit_margin=lim_margin,
                lcp_rows=len(w.lcp()["x"]))


def substeps(env_id, roll, seed):
    spec = SPECS[env_id]
    m = build_model(env_id)
    w = OracleWorld(m)
    rng = np.random.RandomState(seed + 100)
    nd = m.n_dofs
    recs = []
    scale = np.zeros(nd)
    scale[list(spec.task.act_dof)] = spec.task.act_scale
    # (a) states harvested from the reference rollouts (post-step states are contact-rich)
    idx = rng.choice(len(roll["step_q2"]), size=min(120, len(roll["step_q2"])), replace=False)
    for i in idx:
        q, dq = roll["step_q2"][i], roll["step_dq2"][i]
        if not (np.all(np.isfinite(q)) and np.all(np.isfinite(dq))):
            continue
        tau = rng.uniform(-1, 1, nd) * scale
        recs.append(substep_record(w, m, q, dq, tau))
    # (b) hand-built edge cases
    q0, dq0 = m.q_init(), m.dq_init()
    for d, bi in enumerate(m.dof_bodies()):
        b = m.bodies[bi]
        if not b.limit_enforced:
            continue
        for lim in (b.q_lo, b.q_hi):
            for off in (-0.02, -1e-3, 0.0, 1e-3, 0.02):
                q = q0 + rng.uniform(-0.05, 0.05, nd)
                q[1] += 0.5  # airborne: isolate the limit rows
                q[d] = lim + off
                dq = rng.uniform(-3, 3, nd)
                recs.append(substep_record(w, m, q, dq, rng.uniform(-1, 1, nd) * scale))
    if env_id != "DartSnake7Link-v1":
        # ground contact sweep: lower the root until several capsules touch, with tilt
        for drop in np.linspace(0.0, 0.6, 25):
            for ang in (0.0, 0.3, -0.5, 1.2):
                q = q0 + rng.uniform(-0.1, 0.1, nd)
                q[1] = -drop
                q[2] = ang + rng.uniform(-0.05, 0.05)
                for d, bi in enumerate(m.dof_bodies()):  # keep inside limits
                    b = m.bodies[bi]
                    if b.limit_enforced:
                        q[d] = np.clip(q[d], b.q_lo + 0.01, b.q_hi - 0.01)
                dq = rng.uniform(-2, 2, nd)
                recs.append(substep_record(w, m, q, dq, rng.uniform(-1, 1, nd) * scale))
        # exactly flat foot just touching / penetrating (tie rule: contact at endpoint p1)
        for pen in (-1e-3, 0.0, 1e-3, 5e-3):
            q = q0.copy()
            q[1] = {"DartHopper-v1": -0.04, "DartWalker2d-v1": -0.04}.get(env_id, -0.2) - pen
            recs.append(substep_record(w, m, q, np.zeros(nd), np.zeros(nd)))
    # (c) external forces at body origins (bn.add_ext_force path)
    for _ in range(10):
        q = q0 + rng.uniform(-0.3, 0.3, nd)
        q[1] += 0.5
        fext = rng.uniform(-30, 30, (m.n_bodies, 3))
        recs.append(substep_record(w, m, q, rng.uniform(-2, 2, nd), rng.uniform(-1, 1, nd) * scale, fext))
    out = {}
    for k in recs[0]:
        out["sub_" + k] = np.array([r[k] for r in recs])
    return out


def main():
    outdir = os.path.dirname(os.path.abspath(__file__))
    for env_id in SPECS:
        roll = env_rollouts(env_id, n_steps=60, scales=(1.0, 0.3, 0.1, 0.0, 0.3), seed=11)
        sub = substeps(env_id, roll, seed=11)
        name = env_id.replace("-v1", "").replace("Dart", "").lower()
        path = os.path.join(outdir, name + ".npz")
        np.savez_compressed(path, env_id=env_id, **roll, **sub)
        print(env_id, "->", path, "steps", len(roll["step_q"]), "done", int(roll["step_done"].sum()),
              "substeps", len(sub["sub_q"]), "with contact", int((sub["sub_ncontact"] > 0).sum()),
              "max contacts", int(sub["sub_ncontact"].max()), "limit rows", int((sub["sub_limit_active"] != 0).sum()),
              "max lcp rows", int(sub["sub_lcp_rows"].max()), os.path.getsize(path) // 1024, "KiB")


def main_contact_free():
    outdir = os.path.dirname(os.path.abspath(__file__))
    for env_id in REF_CLASSES_CF:
        roll = contact_free_rollouts(env_id, n_steps=80, seed=5)
        name = {"DartCartPole-v1": "cartpole", "DartCartPoleSwingUp-v1": "cartpole_swingup",
                "DartDoubleInvertedPendulumEnv-v1": "double_pendulum", "DartReacher-v1": "reacher2d"}[env_id]
        path = os.path.join(outdir, name + ".npz")
        np.savez_compressed(path, env_id=env_id, **roll)
        print(env_id, "->", path, "steps", len(roll["step_q"]), "done", int(roll["step_done"].sum()), os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    if "--contact-free" not in sys.argv:
        main()
    main_contact_free()
