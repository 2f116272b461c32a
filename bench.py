#!/usr/bin/env python
"""bench.py — env steps/sec of the batched DART stepper (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--config 2|3|4|5] [--no-extras]
                    [--env ID --worlds N --lcp exact|pgs --pgs-iters I --allgather]   (custom workload)

One "step" = one env.step() of every world on this rank: action -> frame_skip DART time steps
-> obs / reward / done -> auto-reset.  The line's workload is BASELINE configs[1] (`--config 2`:
DartHopper-v1, 4096 worlds per GPU, fp32, frame_skip 4, random actions U(-1,1)) unless another
config is asked for; the other GPU configs of BASELINE.json are measured in the same run, shorter,
and reported under "configs" (per-GPU shard sizes; the NCCL obs all-gather of config 4 INSIDE the
timed region when there is more than one rank; the PGS iteration sweep of config 5):

    2  DartHopper-v1       4096 worlds/GPU   exact LCP
    3  DartWalker2d-v1     16384 worlds      PGS LCP path (30 sweeps), exact LCP beside it
    4  DartHalfCheetah-v1  16384 worlds/GPU  exact LCP + all_gather_into_tensor(obs) per step
    5  DartSnake7Link-v1   4096 worlds/GPU   PGS iteration sweep k in {1,2,4,8,16,30,50}

For N>1 launch with torchrun (one rank per GPU); worlds are sharded per GPU (weak scaling), no
data-path collective except config 4's all-gather.

Timing: W warm-up steps, then K steps; each timed step is bracketed by CUDA events on the
launching stream and L2 is flushed (a 256 MiB write, untimed) between steps because the whole
working set is far smaller than L2; max over ranks.  `value` is device-resident throughput;
`e2e` goes through the public host API (DartEnv.step with numpy arrays: the kernel reads the
actions from and writes obs / float64 rewards / bool dones to page-locked host memory, one launch
+ one sync per step) and is the headline against `--impl reference`, which times the CPU
restatement of the reference path (oracle/, "port": pydart2/DART are not installable here) on all
host threads for at least 10 s of CPU work regardless of --steps.  `e2e.value` is the MEAN over
every call (wall clock, L2 flushed between calls, cyclic GC off inside the loop like `timeit`, the
nvidia-smi clocks poll stopped once the device-timed regions end); `e2e.us_per_call_rank0` adds the
p50 / p99 / max of the same calls.
"""
import argparse
import gc
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
os.environ.setdefault("MASTER_ADDR", "127.0.0.1")

import numpy as np  # noqa: E402

UNIT = "env-steps/s"
CONFIGS = {
    2: dict(env="DartHopper-v1", worlds=4096, lcp="exact", pgs_iters=30, allgather=False),
    3: dict(env="DartWalker2d-v1", worlds=16384, lcp="pgs", pgs_iters=30, allgather=False),
    4: dict(env="DartHalfCheetah-v1", worlds=16384, lcp="exact", pgs_iters=30, allgather=True),
    5: dict(env="DartSnake7Link-v1", worlds=4096, lcp="pgs", pgs_iters=30, allgather=False),
}
PGS_SWEEP = (1, 2, 4, 8, 16, 30, 50)


def build_model(env_id):
    from dart_env_b200.skel import load_model
    from dart_env_b200.tasks import SPECS
    spec = SPECS[env_id]
    m = load_model(spec.skel, spec.dt)
    m.enforce_limits()
    if spec.friction_all is not None:
        for b in m.bodies:
            b.friction_coeff = spec.friction_all
    return m, spec


def metric_name(env_id):
    return "env steps/sec (batched worlds) %s" % env_id


def workload_string(env_id, worlds, frame_skip, lcp, pgs_iters):
    """the SAME string in both arms (the driver compares them)"""
    return "%s, %d worlds/GPU, fp32 engine, frame_skip %d, random actions U(-1,1), auto-reset, lcp=%s" % (
        env_id, worlds, frame_skip, lcp if lcp == "exact" else "pgs(%d)" % pgs_iters)


def algo_bytes(nd, nact, nobs):
    """4*(2nd [q,dq in] + nact + 2nd [q,dq out] + nobs + 1 [reward]) + 1 [done] per env step"""
    return 4 * (4 * nd + nact + nobs + 1) + 1


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


class GpuCpuAffinity:
    """Run the GPU phases on the CPUs NVML names as local to the GPU (nvmlDeviceSetCpuAffinity: the NUMA node the GPU's
    PCIe root hangs off), like NCCL does for its own threads: the end-to-end step is a launch + sync round trip plus
    zero-copy PCIe traffic to page-locked memory, both sensitive to a remote socket.  restore() gives the process its
    full mask back before the CPU baseline runs on all host threads.  BENCH_CPU_AFFINITY=0 switches it off."""

    def __init__(self, torch, local_rank):
        self.note, self.full = "off", None
        if os.environ.get("BENCH_CPU_AFFINITY", "1") == "0":
            return
        try:
            import pynvml
            self.full = os.sched_getaffinity(0)
            pynvml.nvmlInit()
            uuid = str(torch.cuda.get_device_properties(local_rank).uuid)
            h = pynvml.nvmlDeviceGetHandleByUUID(uuid if uuid.startswith("GPU-") else "GPU-" + uuid)
            pynvml.nvmlDeviceSetCpuAffinity(h)
            now = os.sched_getaffinity(0)
            self.note = "nvml ideal cpus: %d of %d" % (len(now), len(self.full))
        except Exception as e:   # (no NVML / no permission: run unbound and say so)
            self.note = "unavailable (%s)" % type(e).__name__

    def restore(self):
        if self.full:
            try:
                os.sched_setaffinity(0, self.full)
            except Exception:
                pass


# ------------------------------------------------------------------------------ clocks sampler
class ClockSampler:
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown," \
        "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index):
        self.gpu, self.proc, self.lines = gpu_index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "20"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if self.proc is None:
            return None
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return None
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------ CPU arm
def cpu_arm(env_id, worlds, threads, budget_s, block_steps=25):
    """The CPU restatement of the reference path (oracle 'port'): all `worlds` worlds of the workload, all host
    threads, auto-reset, the same random-action distribution; blocks of `block_steps` env steps are repeated
    until about `budget_s` seconds of CPU work have been timed (so the figure does not depend on --steps)."""
    from oracle import oracle as orc
    m, spec = build_model(env_id)
    orc.cpu_bench(m, spec.task, max(threads * 8, 64), 5, threads, seed=1)  # page in, spin the threads up
    steps, wall = 0, 0.0
    while wall < budget_s and steps < 1000000:
        t0 = time.perf_counter()
        orc.cpu_bench(m, spec.task, worlds, block_steps, threads, seed=3 + steps)
        wall += time.perf_counter() - t0
        steps += block_steps
    return {"value": worlds * steps / wall, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": "%d worlds x %d env steps (%.1f s of CPU work, blocks of %d steps), fp64 scalar C restatement of "
                      "pydart2 World.step (Dantzig LCP) + task layer, %d pthreads" % (worlds, steps, wall, block_steps, threads)}, wall, steps


def run_reference(args, cfg, rank):
    if rank != 0:
        return
    threads = host_threads()
    m, spec = build_model(cfg["env"])
    cb, wall, steps = cpu_arm(cfg["env"], cfg["worlds"], threads, budget_s=10.0, block_steps=max(1, min(args.steps, 25)))
    extras = {}
    if not args.no_extras:
        for cid, c in CONFIGS.items():
            if c["env"] == cfg["env"]:
                continue
            e, _, _ = cpu_arm(c["env"], min(c["worlds"], 4096), threads, budget_s=4.0)
            extras[str(cid)] = {"env": c["env"], "value": e["value"], "unit": UNIT, "cpu_baseline": e}
    line = {"impl": "reference", "metric": metric_name(cfg["env"]), "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * wall / steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_string(cfg["env"], cfg["worlds"], spec.task.frame_skip, cfg["lcp"], cfg["pgs_iters"]),
                       "timed": "%d env steps of all %d worlds = %.1f s on %d host threads (--steps only sets the block size)"
                                % (steps, cfg["worlds"], wall, threads)},
            "cpu_baseline": cb,
            "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0, "configs": extras}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------ GPU arm
def load_profile():
    for name in ("r2_ncu_summary.json", "r1_ncu_summary.json"):
        try:
            return json.load(open(os.path.join(ROOT, "profiles", name))), "profiles/" + name
        except Exception:
            continue
    return {}, None


class Runner:
    """one workload on this rank: engine, action pool, timed loops"""

    def __init__(self, torch, dist, cfg, rank, local_rank, world_size, seed):
        from dart_env_b200.envs import make
        self.torch, self.dist, self.cfg, self.rank, self.world_size = torch, dist, cfg, rank, world_size
        self.dev = torch.device("cuda", local_rank)
        n = self.n = cfg["worlds"]
        self.m, self.spec = build_model(cfg["env"])
        self.nact, self.nobs, self.nd = self.spec.task.n_act, self.spec.task.n_obs, self.m.n_dofs
        self.env = make(cfg["env"], num_envs=n, output="torch", device=local_rank, seed=seed, world_offset=rank * n, batched=True)
        self.eng = self.env.engine
        if cfg["lcp"] == "pgs":
            self.eng.set_lcp(1, cfg["pgs_iters"])
        gen = torch.Generator(device=self.dev)
        gen.manual_seed(1234 + rank)
        # actions: a pool of 64 pre-generated U(-1,1) batches cycled through (BASELINE.md regenerates them on the device
        # every step; a torch RNG kernel inside the timed region would not be this repo's kernel)
        self.pool = [(torch.rand((n, self.nact), generator=gen, device=self.dev) * 2 - 1).contiguous() for _ in range(64)]
        self.gather, self.fused = None, None
        if cfg.get("fused") and world_size > 1:
            # the all-gather inside the step kernel: peer stores over NVLink into every rank's (symmetric-memory) buffer
            from dart_env_b200.parallel import FusedObsGather
            self.fused = FusedObsGather(self.env)
        elif cfg["allgather"] and world_size > 1:
            self.gather = torch.empty((world_size * n, self.nobs), dtype=torch.float32, device=self.dev)
        self.env.reset()
        self.i = 0

    def one_step(self):
        env = self.env
        if self.fused is not None:
            self.fused.step(self.pool[self.i % 64])
        else:
            self.eng.step(self.pool[self.i % 64], env._obs, env._rew, env._done, True)
            if self.gather is not None:
                self.dist.all_gather_into_tensor(self.gather, env._obs)
        self.i += 1

    def warm(self, W):
        for _ in range(W):
            self.one_step()
        self.torch.cuda.synchronize()

    def timed_flushed(self, K, flush):
        torch = self.torch
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
        l0 = self.eng.launch_count
        for i in range(K):
            flush.fill_(float(i & 1))
            ev[i][0].record()
            self.one_step()
            ev[i][1].record()
        torch.cuda.synchronize()
        return float(sum(a.elapsed_time(b) for a, b in ev)), self.eng.launch_count - l0

    def timed_warm(self, K):
        torch = self.torch
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(K):
            self.one_step()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1)

    def done_fraction(self):
        return float((self.env._done != 0).float().mean().item())

    def close(self):
        self.env.close()


def reduce_max(torch, dist, dev, world_size, vals):
    if world_size == 1:
        return list(vals)
    t = torch.tensor(list(vals), dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(x) for x in t.tolist()]


def roofline(r, kern_ms, peak, peak_src, prof, prof_src):
    bpl = algo_bytes(r.nd, r.nact, r.nobs) * r.n
    achieved = bpl / (kern_ms * 1e-3) / 1e9
    roof = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
            "peak_source": peak_src, "kernel": r.eng.kernel_name, "algorithmic_bytes_per_env_step": bpl // r.n,
            "note": "fused per-world stepper: only the API-boundary I/O crosses HBM, so the HBM fraction is small by "
                    "construction; the kernel is bound by dependent-issue latency / instruction issue (issue_active_pct, "
                    "avg_active_lanes from the committed ncu capture of this kernel)"}
    form = "coop" if "coop:" in r.eng.kernel_name else ("quad" if "quad:" in r.eng.kernel_name else ("loop" if "loop:" in r.eng.kernel_name else "static"))
    key = "%s/%d/%s" % (r.cfg["env"], r.n, form)
    p = prof.get(key) or prof.get(r.cfg["env"] if "coop:" in r.eng.kernel_name else r.cfg["env"] + "/static")
    if isinstance(p, dict):
        roof["traffic"] = p.get("dram_bytes_per_launch")
        for k in ("fp32", "fp64", "issue_active_pct", "warps_active_pct", "avg_active_lanes", "stall_pct", "duration_us"):
            if k in p:
                roof[k if k != "duration_us" else "ncu_duration_us"] = p[k]
        roof["profile"] = "%s[%s]" % (prof_src, key)
    return roof


def run_ours(args, cfg, rank, local_rank, world_size):
    import torch
    import torch.distributed as dist

    from dart_env_b200.envs import make

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the engine has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    affinity = GpuCpuAffinity(torch, local_rank)
    if world_size > 1:
        dist.init_process_group("nccl", device_id=dev)
    K, W = args.steps, max(args.warmup, 3)
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)
    prof, prof_src = load_profile()
    peaks, peak_src = None, "fallback (B200_PROFILING.md)"
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        peak_src = "MEASURED_PEAKS.json"
    except Exception:
        pass
    peak = float(peaks["hbm_gbs"]) if peaks else 6650.0

    def barrier():
        torch.cuda.synchronize()
        if world_size > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- headline workload
    r = Runner(torch, dist, cfg, rank, local_rank, world_size, args.seed)
    n, nact, nobs = r.n, r.nact, r.nobs
    sampler = ClockSampler(local_rank)
    sampler.start()
    r.warm(W)
    t_w = time.perf_counter()
    while not sampler.lines and sampler.proc is not None and time.perf_counter() - t_w < 2.0:
        r.one_step()   # keep the GPU under load until nvidia-smi delivers its first sample
        torch.cuda.synchronize()
    barrier()
    t_wall0 = time.perf_counter()
    dev_ms, launches = r.timed_flushed(K, flush)
    t_wall = time.perf_counter() - t_wall0
    barrier()
    warm_ms = r.timed_warm(K)
    done_frac = r.done_fraction()
    # (sampled over the two device-timed regions.  It stops HERE: every nvidia-smi query takes driver locks that a kernel
    # launch waits on, and the end-to-end figure below is wall clock per call — with the 20 ms poll running, one call in
    # ~350 took milliseconds and the mean moved by 2-30 us from run to run on the same box, gpurun_out/r2v_probe.log)
    keep_sampler = os.environ.get("BENCH_SAMPLE_DURING_E2E", "0") == "1"   # (A/B switch for the sentence above)
    clocks = None if keep_sampler else sampler.stop()

    # ---- end to end through the public host API (numpy in / numpy out, page-locked output slots)
    henv = make(cfg["env"], num_envs=n, output="numpy", device=local_rank, seed=args.seed, world_offset=rank * n, batched=True)
    if cfg["lcp"] == "pgs":
        henv.engine.set_lcp(1, cfg["pgs_iters"])
    henv.reset()
    rng = np.random.RandomState(99 + rank)
    hpool = [rng.uniform(-1, 1, (n, nact)).astype(np.float32) for _ in range(16)]
    for i in range(W):
        henv.step(hpool[i % 16])
    barrier()
    Ke2e = max(K, 200)      # (wall-clock per call: at least 200 calls so that a short --steps run is not one scheduler hiccup)
    e2e_t = np.empty(Ke2e)
    # (like `timeit`: the cyclic garbage collector is off inside the timed loop.  A generation-2 pass over the heap torch
    # leaves behind takes 4-15 ms and lands in one call of ~350 — measured as the whole run-to-run spread of this figure,
    # 53-71 us around a p50 of 50 us, gpurun_out/r2w_probe.log; the percentiles are reported next to the mean)
    gc.collect()
    gc.disable()
    for i in range(Ke2e):
        flush.fill_(float(i & 1))
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        o, rw, d, _ = henv.step(hpool[i % 16])
        e2e_t[i] = time.perf_counter() - t0
    gc.enable()
    e2e_s = float(e2e_t.sum())   # the figure is the MEAN over every call; the percentiles say how it is distributed
    e2e_pct = {"p50": float(np.percentile(e2e_t, 50)) * 1e6, "p99": float(np.percentile(e2e_t, 99)) * 1e6, "max": float(e2e_t.max()) * 1e6}
    assert o.dtype == np.float32 and rw.dtype == np.float64 and d.dtype == np.bool_
    h2d, d2h = n * nact * 4, n * nobs * 4 + n * 8 + n
    henv.close()
    if keep_sampler:
        clocks = sampler.stop()
    dev_ms, warm_ms, e2e_s = reduce_max(torch, dist, dev, world_size, (dev_ms, warm_ms, e2e_s))
    total_worlds = n * world_size
    roof = roofline(r, dev_ms / K, peak, peak_src, prof, prof_src) if rank == 0 else None
    kernel_name = r.eng.kernel_name
    r.close()

    # ---- the other GPU configs of BASELINE.json, shorter, in the same run
    extras = {}
    if not args.no_extras:
        Ke, We = max(20, min(K, 200)), max(3, min(W, 20))
        for cid, c in CONFIGS.items():
            if c["env"] == cfg["env"] and c["worlds"] == cfg["worlds"]:
                continue
            variants = [c]
            if cid == 3:
                variants = [c, dict(c, lcp="exact")]
            if cid == 4 and world_size > 1:
                variants = [c, dict(c, fused=True)]
            out = {"env": c["env"], "worlds_per_gpu": c["worlds"], "runs": []}
            for v in variants:
                try:
                    x = Runner(torch, dist, v, rank, local_rank, world_size, args.seed)
                except Exception as exc:   # (symmetric memory unavailable on this box: the fused gather is reported as such)
                    out["runs"].append({"lcp": v["lcp"], "error": "%s: %s" % (type(exc).__name__, str(exc)[:200])})
                    continue
                x.warm(We)
                barrier()
                ms, nl = x.timed_flushed(Ke, flush)
                barrier()
                wm = x.timed_warm(Ke)
                ms, wm = reduce_max(torch, dist, dev, world_size, (ms, wm))
                run = {"lcp": v["lcp"] if v["lcp"] == "exact" else "pgs(%d)" % v["pgs_iters"],
                       "value": x.n * world_size * Ke / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms / Ke,
                       "value_l2_warm": x.n * world_size * Ke / (wm * 1e-3), "ms_per_step_l2_warm": wm / Ke, "steps": Ke,
                       "gpu_launches": int(nl), "done_fraction": x.done_fraction(),
                       "parallelism": "worlds sharded x%d%s" % (world_size,
                           ", obs all-gather FUSED into the step kernel (peer stores over NVLink into every rank's [%d,%d] buffer + one device-side barrier) inside every timed step"
                           % (x.n * world_size, x.nobs) if x.fused is not None else
                           (", NCCL all_gather_into_tensor(obs [%d,%d] fp32) inside every timed step" % (x.n * world_size, x.nobs) if x.gather is not None else
                            (", obs all-gather is a no-op on one rank" if v["allgather"] else ", no collective")))}
                if rank == 0:
                    run["roofline"] = roofline(x, ms / Ke, peak, peak_src, prof, prof_src)
                out["runs"].append(run)
                if cid == 5 and v is variants[-1]:
                    sweep = []
                    for k in PGS_SWEEP:
                        x.eng.set_lcp(1, k)
                        x.warm(5)
                        barrier()
                        wk = reduce_max(torch, dist, dev, world_size, (x.timed_warm(max(20, Ke // 2)),))[0] / max(20, Ke // 2)
                        sweep.append({"pgs_iters": k, "us_per_step_l2_warm": 1e3 * wk, "value": x.n * world_size / (wk * 1e-3)})
                    out["pgs_sweep"] = sweep
                x.close()
            extras[str(cid)] = out

    line = None
    affinity.restore()   # the CPU baseline runs on every host thread the process was given
    if rank == 0:
        cb, _, _ = cpu_arm(cfg["env"], n, host_threads(), budget_s=10.0)
        if not args.no_extras:
            for cid, c in CONFIGS.items():
                if str(cid) in extras:
                    extras[str(cid)]["cpu_baseline"] = cpu_arm(c["env"], min(c["worlds"], 4096), host_threads(), budget_s=4.0)[0]
        line = {"metric": metric_name(cfg["env"]), "value": total_worlds * K / (dev_ms * 1e-3), "unit": UNIT, "n_gpus": world_size,
                "steps": K, "warmup": W, "ms_per_step": dev_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic",
                "config": {"workload": workload_string(cfg["env"], n, r.spec.task.frame_skip, cfg["lcp"], cfg["pgs_iters"]),
                           "worlds_per_gpu": n,
                           "parallelism": "worlds sharded x%d, %s" % (world_size, "obs all_gather_into_tensor inside the timed step"
                                                                     if (cfg["allgather"] and world_size > 1) else "no data-path collective"),
                           "actions": "pool of 64 pre-generated U(-1,1) batches per rank, cycled (not regenerated per step)",
                           "l2": "flushed between timed steps (256 MiB write, untimed); working set %.1f MB << L2"
                                 % (algo_bytes(r.nd, nact, nobs) * n / 1e6),
                           "timing": "sum of per-step CUDA-event durations on the launching stream, max over ranks",
                           "cpu_affinity": affinity.note},
                "value_l2_warm": total_worlds * K / (warm_ms * 1e-3), "ms_per_step_l2_warm": warm_ms / K,
                "wall_s_timed_loop": t_wall, "done_fraction": done_frac, "kernel": kernel_name,
                "e2e": {"value": total_worlds * Ke2e / e2e_s, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "ms_per_step": 1e3 * e2e_s / Ke2e, "steps": Ke2e, "us_per_call_rank0": e2e_pct,
                        "api": "DartEnv.step(numpy float32 [N,nact]) -> (float32 obs, float64 rewards, bool dones): one launch + one "
                               "sync; the kernel reads / writes page-locked host memory itself"},
                "gpu_launches": int(launches), "roofline": roof, "cpu_baseline": cb, "clocks": clocks, "configs": extras}
        print(json.dumps(line), flush=True)
    if world_size > 1:
        dist.barrier()
        dist.destroy_process_group()
    return line


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=50)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=2, choices=sorted(CONFIGS), help="BASELINE.json config the line is quoted on")
    ap.add_argument("--no-extras", action="store_true", help="skip the other configs (reported under 'configs')")
    ap.add_argument("--env", default=None, help="custom workload: env id (overrides --config)")
    ap.add_argument("--worlds", type=int, default=None, help="worlds per GPU")
    ap.add_argument("--allgather", action="store_true")
    ap.add_argument("--lcp", default=None, choices=["exact", "pgs"])
    ap.add_argument("--pgs-iters", type=int, default=None)
    ap.add_argument("--seed", type=int, default=0)
    args = ap.parse_args()
    cfg = dict(CONFIGS[args.config])
    if args.env:
        cfg.update(env=args.env, lcp="exact", allgather=False, worlds=4096)
        args.no_extras = True
    for k, v in (("worlds", args.worlds), ("lcp", args.lcp), ("pgs_iters", args.pgs_iters)):
        if v is not None:
            cfg[k] = v
    if args.allgather:
        cfg["allgather"] = True
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world_size = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, cfg, rank)
        return
    if world_size == 1 and args.gpus > 1:
        raise SystemExit("bench.py: --gpus %d needs torchrun (python -m torch.distributed.run --nproc-per-node %d ...)"
                         % (args.gpus, args.gpus))
    run_ours(args, cfg, rank, local_rank, world_size)


if __name__ == "__main__":
    main()
