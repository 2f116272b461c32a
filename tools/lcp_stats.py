"""Developer tool (CPU): LCP workload statistics of the KERNEL SOURCE on closed-loop rollouts.

Worlds are driven by the oracle env (random actions, auto-reset); at every env step the pre-step
states go through tools/host_emu (the product's device code compiled for the CPU, fp32, with the
per-world hint word carried across the frame_skip DART steps) and the emulation's counters are
read back: row-count histogram, pivoting iterations per solve, which solver path answered.
Not a product path."""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from dart_env_b200.skel import load_model  # noqa: E402
from dart_env_b200.tasks import SPECS  # noqa: E402
from oracle import oracle as orc  # noqa: E402
from tools.host_emu import emu  # noqa: E402


def run(env_id, n_worlds=64, n_steps=150, seed=0):
    spec = SPECS[env_id]
    m = load_model(spec.skel, spec.dt)
    m.enforce_limits()
    if spec.friction_all is not None:
        for b in m.bodies:
            b.friction_coeff = spec.friction_all
    task = spec.task
    rng = np.random.default_rng(seed)
    envs = [orc.OracleEnv(m, task, seed=seed, world_id=w) for w in range(n_worlds)]
    for e in envs:
        e.reset()
    L = emu.lib()
    cnt = (C.c_long * 8)()
    hist = (C.c_long * 32)()
    L.emu_counters(cnt, 1)
    L.emu_hist(hist, 1)
    nd = m.n_dofs
    nact = len(task.act_dof)
    hints = np.full(n_worlds, np.iinfo(np.uint64).max, dtype=np.uint64)
    err = []
    for step in range(n_steps):
        a = rng.uniform(-1, 1, size=(n_worlds, nact))
        q = np.array([e.world.get_state()[0] for e in envs])
        dq = np.array([e.world.get_state()[1] for e in envs])
        tau = np.zeros((n_worlds, nd))
        tau[:, task.act_dof] = np.clip(a, -1, 1) * np.array(task.act_scale)
        qe, dqe = q, dq
        for _ in range(task.frame_skip):
            qe, dqe, *_rest = emu.substep(m, task, qe, dqe, tau, f64=False, hints=hints)
        for w, e in enumerate(envs):
            _, _, d = e.step(a[w])
            if d or (step + 1) % spec.max_episode_steps == 0:
                e.reset()
                hints[w] = np.iinfo(np.uint64).max
        if not task.fluid_force:
            q2 = np.array([e.world.get_state()[0] for e in envs])
            err.append(np.abs(qe - q2).max())
    L.emu_counters(cnt, 0)
    L.emu_hist(hist, 0)
    c = list(cnt)
    h = np.array(list(hist))
    print(f"{env_id}: exact-LCP calls {c[0]} of {n_worlds * n_steps * task.frame_skip} DART steps; "
          f"small<4/6> {c[1]} small<8> {c[2]} bpp_local {c[5]} dantzig {c[3]} small-fail {c[6]}; "
          f"pivot iterations {c[4]} ({c[4] / max(1, c[0]):.2f}/solve)")
    print("  rows histogram:", {i: int(v) for i, v in enumerate(h) if v})
    return c, h


if __name__ == "__main__":
    ids = sys.argv[1:] or ["DartHopper-v1", "DartWalker2d-v1", "DartHalfCheetah-v1"]
    for env_id in ids:
        run(env_id)
