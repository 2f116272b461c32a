"""Developer probe (run under gpurun): error statistics of the CUDA engine against the golden
vectors, plus a quick timing.  Not a test; tests/test_gpu_parity.py asserts the tolerances."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

from dart_env_b200.engine import Engine
from dart_env_b200.skel import load_model
from dart_env_b200.tasks import SPECS

FILES = {"DartHopper-v1": "hopper.npz", "DartWalker2d-v1": "walker2d.npz",
         "DartHalfCheetah-v1": "halfcheetah.npz", "DartSnake7Link-v1": "snake7link.npz"}


def model_for(env_id):
    spec = SPECS[env_id]
    m = load_model(spec.skel, spec.dt)
    m.enforce_limits()
    if spec.friction_all is not None:
        for b in m.bodies:
            b.friction_coeff = spec.friction_all
    return m, spec


def main():
    dev = torch.device("cuda", 0)
    print(torch.cuda.get_device_name(0))
    for env_id in SPECS:
        g = np.load(os.path.join(ROOT, "tests", "golden", FILES[env_id]))
        m, spec = model_for(env_id)
        n = len(g["sub_q"])
        for f64 in (True, False):
            dt = torch.float64 if f64 else torch.float32
            eng = Engine(m, spec.task, n, f64=f64)
            q = torch.tensor(g["sub_q"], dtype=dt, device=dev)
            dq = torch.tensor(g["sub_dq"], dtype=dt, device=dev)
            tau = torch.tensor(g["sub_tau"], dtype=dt, device=dev)
            fext = torch.tensor(g["sub_fext"], dtype=dt, device=dev).contiguous()
            eng.set_state(q, dq)
            eng.substep(tau, fext)
            q2, dq2 = eng.get_state(torch.float64)
            torch.cuda.synchronize()
            eq = (q2.cpu().numpy() - g["sub_q2"])
            ev = (dq2.cpu().numpy() - g["sub_dq2"])
            cnt, body, data = eng.contacts()
            cnt = cnt.cpu().numpy()
            body = body.cpu().numpy()
            ncm = (cnt != g["sub_ncontact"])
            mc = body.shape[1]
            bm = (body[:, :min(mc, 8)] != g["sub_contact_body"][:, :min(mc, 8)]).any(1)
            scale_v = 1 + np.abs(g["sub_dq2"])
            worst = np.argmax(np.abs(ev / scale_v).max(1))
            print("%-20s %s substep: n=%d max|dq|=%.3e max|dq_rel|=%.3e (row %d, ncontact %d, rows %d, margins c=%.1e t=%.1e l=%.1e) max|q|=%.3e  ncontact mismatch %d body mismatch %d  [%s]"
                  % (env_id, "f64" if f64 else "f32", n, np.abs(ev).max(), np.abs(ev / scale_v).max(), worst,
                     g["sub_ncontact"][worst], g["sub_lcp_rows"][worst], g["sub_contact_margin"][worst],
                     g["sub_tie_margin"][worst], g["sub_limit_margin"][worst], np.abs(eq).max(), ncm.sum(), bm.sum(),
                     eng.kernel_name))
            if ncm.sum():
                idx = np.where(ncm)[0][:5]
                print("   contact-count mismatches at", idx, "margins", g["sub_contact_margin"][idx])
            # percentile view
            rel = np.abs(ev / scale_v).max(1)
            print("   dq rel err percentiles 50/90/99/100: %.2e %.2e %.2e %.2e" % tuple(np.percentile(rel, [50, 90, 99, 100])))
            # contact data
            d = data.cpu().numpy()
            gd = g["sub_contact_data"][:, :min(mc, 8)]
            ok = ~ncm
            if ok.any() and mc > 0:
                dd = d[:, :gd.shape[1]] - gd
                print("   contact point/normal/depth err %.2e, force err %.2e (max |force| %.1f)"
                      % (np.abs(dd[ok][..., :7]).max(), np.abs(dd[ok][..., 7:]).max(), np.abs(gd[..., 7:]).max()))
            eng.close()
        # env step
        n = len(g["step_q"])
        for f64 in (True, False):
            dt = torch.float64 if f64 else torch.float32
            eng = Engine(m, spec.task, n, f64=f64)
            eng.set_state(torch.tensor(g["step_q"], dtype=dt, device=dev), torch.tensor(g["step_dq"], dtype=dt, device=dev))
            act = torch.tensor(g["step_action"], dtype=torch.float32, device=dev)
            obs = torch.empty((n, spec.task.n_obs), dtype=torch.float32, device=dev)
            rew = torch.empty((n,), dtype=torch.float32, device=dev)
            done = torch.empty((n,), dtype=torch.uint8, device=dev)
            eng.step(act, obs, rew, done, auto_reset=False)
            q2, dq2 = eng.get_state(torch.float64)
            torch.cuda.synchronize()
            fin = np.isfinite(g["step_obs"]).all(1)
            eo = np.abs(obs.cpu().numpy().astype(np.float64) - g["step_obs"])[fin]
            er = np.abs(rew.cpu().numpy() - g["step_reward"])[fin]
            dm = (done.cpu().numpy().astype(bool) != g["step_done"])
            ev = np.abs(dq2.cpu().numpy() - g["step_dq2"])[fin] / (1 + np.abs(g["step_dq2"][fin]))
            print("%-20s %s envstep: n=%d max obs err %.3e  reward err %.3e  dq rel %.3e (p50 %.1e p99 %.1e)  done mismatch %d (margins %s)"
                  % (env_id, "f64" if f64 else "f32", n, eo.max(), er.max(), ev.max(), np.percentile(ev.max(1), 50),
                     np.percentile(ev.max(1), 99), dm.sum(), g["step_margin"][dm][:5]))
            eng.close()
    # timing
    for env_id, n in (("DartHopper-v1", 4096), ("DartWalker2d-v1", 16384), ("DartHalfCheetah-v1", 16384), ("DartSnake7Link-v1", 4096)):
        m, spec = model_for(env_id)
        eng = Engine(m, spec.task, n, seed=1)
        obs = eng.reset()
        rew = torch.empty((n,), dtype=torch.float32, device=dev)
        done = torch.empty((n,), dtype=torch.uint8, device=dev)
        gen = torch.Generator(device=dev)
        gen.manual_seed(1234)
        acts = [torch.rand((n, spec.task.n_act), generator=gen, device=dev) * 2 - 1 for _ in range(16)]
        for i in range(50):
            eng.step(acts[i % 16], obs, rew, done)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        K = 500
        e0.record()
        for i in range(K):
            eng.step(acts[i % 16], obs, rew, done)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / K
        print("%-20s n=%6d  %.1f us/step  %.3e env-steps/s   done frac %.3f" % (env_id, n, ms * 1e3, n / (ms * 1e-3), done.float().mean().item()))
        eng.close()


if __name__ == "__main__":
    main()
