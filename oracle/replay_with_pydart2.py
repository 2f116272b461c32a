#!/usr/bin/env python
"""TEST INFRASTRUCTURE (not a product path): pin the oracle against REAL pydart2 / DART 6.

The reference's physics lives in pydart2 + DART 6 + ODE, none of which is under /root/reference or
installable here, so the fp64 oracle of this repo is "parity unpinned".  This script
is the pin for whoever has a box with pydart2: it replays the committed, engine-agnostic golden
triples (tests/golden/*.npz: (q, dq, tau) -> (q', dq', contact set) for single DART steps, and
(q, dq, action) -> (obs, reward, done) for whole env steps of the reference's own env classes) through
the real engine and reports the deltas against what the oracle minted.

    python oracle/replay_with_pydart2.py [--skel-dir /path/to/gym/envs/dart/assets] [--tol 1e-6]

Exit status: 0 = pydart2 unavailable (prints why) or every sample within tolerance; 1 = deltas found.
"""
import argparse
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
SKELS = {"hopper.npz": ("hopper_capsule.skel", 0.002), "walker2d.npz": ("walker2d.skel", 0.002),
         "halfcheetah.npz": ("half_cheetah.skel", 0.01), "snake7link.npz": ("snake_7link.skel", 0.002)}


def real_pydart2():
    """import the REAL pydart2 (never oracle/pydart2_shim, which is this repo's oracle in disguise)"""
    sys.path[:] = [p for p in sys.path if "pydart2_shim" not in p]
    try:
        mod = importlib.import_module("pydart2")
    except Exception as exc:  # ModuleNotFoundError, or a broken native build
        return None, "import pydart2 failed: %s" % exc
    if "pydart2_shim" in (getattr(mod, "__file__", "") or ""):
        return None, "only the oracle's pydart2 shim is importable"
    return mod, ""


def replay_substeps(pydart, skel_path, dt, g, snake):
    pydart.init()
    world = pydart.World(dt, skel_path)
    robot = world.skeletons[-1]
    for jt in robot.joints:  # dart_env.py:64-67
        for dof in range(len(jt.dofs)):
            if jt.has_position_limit(dof):
                jt.set_position_limit_enforced(True)
    try:
        world.set_collision_detector(3)  # hopper.py:14-18: ODE, else Bullet
    except Exception:
        world.set_collision_detector(2)
    if snake:
        for bn in robot.bodynodes:  # snake_7link.py:29-31
            bn.set_friction_coeff(0.0)
    eq, ev, ec = [], [], 0
    for i in range(len(g["sub_q"])):
        world.reset()
        robot.set_positions(g["sub_q"][i]); robot.set_velocities(g["sub_dq"][i])
        fext = g["sub_fext"][i]
        for b, f in enumerate(fext):
            if np.any(f != 0):
                robot.bodynodes[b].add_ext_force(f)
        robot.set_forces(g["sub_tau"][i])
        world.step()
        q2, dq2 = np.array(robot.q), np.array(robot.dq)
        eq.append(np.abs(q2 - g["sub_q2"][i]).max() / (1 + np.abs(g["sub_q2"][i]).max()))
        ev.append(np.abs(dq2 - g["sub_dq2"][i]).max() / (1 + np.abs(g["sub_dq2"][i]).max()))
        nc = len(world.collision_result.contacts)
        if g["sub_contact_margin"][i] > 1e-9 and nc != int(g["sub_ncontact"][i]):
            ec += 1
    return np.array(eq), np.array(ev), ec


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--skel-dir", default=os.environ.get("DART_ENV_ASSETS", "/root/reference/gym/envs/dart/assets"))
    ap.add_argument("--tol", type=float, default=1e-6)
    args = ap.parse_args()
    pydart, why = real_pydart2()
    if pydart is None:
        print("replay_with_pydart2: unavailable (%s); the oracle stays unpinned for the physics" % why)
        return 0
    bad = 0
    for f, (skel, dt) in SKELS.items():
        g = np.load(os.path.join(GOLD, f))
        eq, ev, ec = replay_substeps(pydart, os.path.join(args.skel_dir, skel), dt, g, f.startswith("snake"))
        print("%-16s %4d DART steps: q err median %.2e max %.2e | dq err median %.2e max %.2e | contact-count mismatches %d"
              % (f, len(eq), np.median(eq), eq.max(), np.median(ev), ev.max(), ec))
        bad += int((eq > args.tol).sum() + (ev > args.tol * 100).sum() + ec)
    print("replay_with_pydart2: %s" % ("all samples within tolerance: oracle pinned" if bad == 0 else "%d samples differ" % bad))
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
