"""Developer sweep (under gpurun): step time vs block size / world count.  Each configuration runs
in a fresh process because DARTB_BLOCK is read once."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CODE = r'''
import sys, os
sys.path.insert(0, %r)
import torch
from dart_env_b200.envs import make
env_id, n = sys.argv[1], int(sys.argv[2])
env = make(env_id, num_envs=n, output="torch", seed=1, batched=True)
eng = env.engine
if os.environ.get("SWEEP_PGS"): eng.set_lcp(1, int(os.environ["SWEEP_PGS"]))
dev = eng.device
obs = env.reset()
gen = torch.Generator(device=dev); gen.manual_seed(1234)
nact = eng.n_act
acts = [torch.rand((n, nact), generator=gen, device=dev) * 2 - 1 for _ in range(16)]
for i in range(50): eng.step(acts[i %% 16], env._obs, env._rew, env._done, True)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
K = 300
e0.record()
for i in range(K): eng.step(acts[i %% 16], env._obs, env._rew, env._done, True)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / K
print("%%-20s n=%%7d wpw=%%4s block=%%4s cw=%%s lcp=%%s %%8.1f us/step  %%.3e env-steps/s  [%%s]" %% (env_id, n, os.environ.get("DARTB_WPW", "auto"), os.environ.get("DARTB_BLOCK", "auto"), os.environ.get("DARTB_COOP_WARPS", "-"), os.environ.get("SWEEP_PGS", "exact"), ms * 1e3, n / (ms * 1e-3), eng.kernel_name))
''' % ROOT

cfgs = []
mode = sys.argv[1] if len(sys.argv) > 1 else "block"
if mode == "variant":
    for v in ("0", "1"):
        for env_id, n in (("DartHopper-v1", 4096), ("DartHopper-v1", 65536), ("DartWalker2d-v1", 16384),
                          ("DartHalfCheetah-v1", 16384), ("DartSnake7Link-v1", 4096)):
            for b in (("32", "128") if n <= 4096 else ("128",)):
                cfgs.append((env_id, n, b, v))
elif mode == "main":
    for env_id, n in (("DartHopper-v1", 4096), ("DartHopper-v1", 65536), ("DartWalker2d-v1", 16384),
                      ("DartHalfCheetah-v1", 16384), ("DartSnake7Link-v1", 4096)):
        cfgs.append((env_id, n, "32" if n <= 4096 else "128", "0"))
elif mode == "coop":   # per-thread (0) vs lane-cooperative (2) kernels across batch sizes
    for env_id, sizes in (("DartHopper-v1", (1024, 4096, 16384, 65536)), ("DartWalker2d-v1", (4096, 16384)),
                          ("DartHalfCheetah-v1", (4096, 16384)), ("DartSnake7Link-v1", (4096, 32768))):
        for n in sizes:
            for v in ("0", "2"):
                cfgs.append((env_id, n, "32" if n <= 4096 else "128", v))
elif mode == "coopw":   # cooperative kernel: warps per block, and the crossover batch sizes
    for cw in ("1", "2", "4"):
        for env_id, n in (("DartHopper-v1", 4096), ("DartHalfCheetah-v1", 4096)):
            cfgs.append((env_id, n, "32", "2", "", "", cw))
    for env_id, sizes in (("DartHopper-v1", (2048, 6144, 8192, 12288)), ("DartWalker2d-v1", (2048, 8192)), ("DartHalfCheetah-v1", (2048, 8192, 12288)),
                          ("DartSnake7Link-v1", (512, 1024, 2048))):
        for n in sizes:
            for v in ("0", "2"):
                cfgs.append((env_id, n, "32" if n <= 4096 else "128", v))
elif mode == "coopq":   # quick check of the cooperative kernel after a change
    for env_id, n in (("DartHopper-v1", 1024), ("DartHopper-v1", 4096), ("DartWalker2d-v1", 4096), ("DartHalfCheetah-v1", 4096),
                      ("DartSnake7Link-v1", 2048)):
        cfgs.append((env_id, n, "32", "2"))
elif mode == "coopbig":   # cooperative kernel at large batches (register-cap experiments)
    for env_id, n in (("DartHopper-v1", 4096), ("DartHopper-v1", 16384), ("DartWalker2d-v1", 16384), ("DartHalfCheetah-v1", 16384),
                      ("DartHalfCheetah-v1", 4096)):
        cfgs.append((env_id, n, "32" if n <= 4096 else "128", "2"))
elif mode == "big":   # per-thread kernels at their large-batch sizes
    for env_id, n in (("DartHopper-v1", 65536), ("DartHopper-v1", 16384), ("DartWalker2d-v1", 16384), ("DartHalfCheetah-v1", 16384),
                      ("DartSnake7Link-v1", 32768)):
        cfgs.append((env_id, n, "128", "0"))
elif mode == "xover":   # both forms right at the automatic crossover sizes
    for env_id, n in (("DartHopper-v1", 4736), ("DartWalker2d-v1", 6144), ("DartHalfCheetah-v1", 12288), ("DartSnake7Link-v1", 2368)):
        for v in ("0", "2"):
            cfgs.append((env_id, n, "128", v))
elif mode == "r2big":   # round 2: the large-batch per-thread kernels, launch shapes x exact / PGS
    for env_id, n in (("DartWalker2d-v1", 16384), ("DartHalfCheetah-v1", 16384)):
        for b, wpw, pgs in (("32", "", ""), ("128", "", ""), ("32", "16", ""), ("32", "8", ""), ("128", "", "30")):
            cfgs.append((env_id, n, b, "0", pgs, wpw))
    cfgs.append(("DartHopper-v1", 65536, "128", "0"))
    cfgs.append(("DartHopper-v1", 4096, "32", "0"))
elif mode == "r2pgs":   # round 2: the PGS path (BASELINE configs 3 and 5) next to the exact solver
    for env_id, n in (("DartWalker2d-v1", 16384), ("DartHalfCheetah-v1", 16384), ("DartSnake7Link-v1", 4096), ("DartHopper-v1", 65536)):
        for pgs in ("", "30", "8", "1"):
            cfgs.append((env_id, n, "128" if n > 4096 else "32", "0", pgs))
    for pgs in ("", "30"):
        cfgs.append(("DartSnake7Link-v1", 4096, "32", "2", pgs))
elif mode == "r2rc":   # round 2: per-thread kernels at their large-batch sizes, exact and PGS(30)
    for env_id, n in (("DartWalker2d-v1", 16384), ("DartHalfCheetah-v1", 16384), ("DartHopper-v1", 65536), ("DartHopper-v1", 16384)):
        for pgs in ("", "30"):
            cfgs.append((env_id, n, "128", "0", pgs))
elif mode == "r2quad":   # round 2: quad form (3) vs one world per thread (0) vs lane-cooperative (2)
    for env_id, sizes in (("DartWalker2d-v1", (16384, 8192, 4096)), ("DartHalfCheetah-v1", (16384, 8192, 4096)), ("DartHopper-v1", (65536, 16384, 4096)),
                          ("DartSnake7Link-v1", (32768, 4096))):
        for n in sizes:
            for v in ("0", "3") + (("2",) if n <= 8192 else ()):
                for pgs in ("", "30"):
                    if pgs and v == "2":
                        continue
                    cfgs.append((env_id, n, "128", v, pgs))
elif mode == "r2group":   # round 2: group forms 2 / 4 / 8 lanes per world (variants 4 / 3 / 5) around the crossover sizes
    for env_id, sizes in (("DartHopper-v1", (2048, 4096, 8192, 16384)), ("DartWalker2d-v1", (4096, 8192, 16384)), ("DartHalfCheetah-v1", (4096, 8192, 16384)),
                          ("DartSnake7Link-v1", (2048, 4096, 8192))):
        for n in sizes:
            for v in ("4", "3", "5"):
                cfgs.append((env_id, n, "128", v))
elif mode == "r2xover":   # round 2: the three forms around the automatic crossovers
    for env_id, sizes in (("DartHopper-v1", (1024, 2048, 3072, 10240, 12288)), ("DartWalker2d-v1", (2048, 3072, 10240, 12288)),
                          ("DartHalfCheetah-v1", (10240, 12288)), ("DartSnake7Link-v1", (1024, 12288, 16384))):
        for n in sizes:
            for v in (("2", "3") if n <= 3072 else (("0", "3") if env_id != "DartHalfCheetah-v1" else ("0", "2"))):
                cfgs.append((env_id, n, "128", v))
elif mode == "r2q4mid":
    for env_id, n in (("DartHopper-v1", 4096), ("DartHopper-v1", 8192), ("DartWalker2d-v1", 4096), ("DartWalker2d-v1", 8192), ("DartSnake7Link-v1", 4096), ("DartHalfCheetah-v1", 8192)):
        cfgs.append((env_id, n, "128", "3"))
elif mode == "r2final":   # the automatic choice at every config size, exact and PGS(30)
    for env_id, sizes in (("DartHopper-v1", (1024, 4096, 8192, 16384, 65536)), ("DartWalker2d-v1", (2048, 4096, 8192, 16384)),
                          ("DartHalfCheetah-v1", (4096, 8192, 16384)), ("DartSnake7Link-v1", (1024, 4096, 8192, 32768))):
        for n in sizes:
            cfgs.append((env_id, n, "128", "-1"))
    for env_id, n in (("DartWalker2d-v1", 16384), ("DartHalfCheetah-v1", 16384), ("DartSnake7Link-v1", 4096), ("DartHopper-v1", 4096)):
        cfgs.append((env_id, n, "128", "-1", "30"))
    for env_id, n, v in (("DartHopper-v1", 4096, "2"), ("DartHopper-v1", 4096, "0"), ("DartWalker2d-v1", 4096, "2"), ("DartHalfCheetah-v1", 8192, "3"), ("DartWalker2d-v1", 16384, "3")):
        cfgs.append((env_id, n, "128", v))
elif mode == "r2p":   # after the branch-free solvers: the forms around every automatic crossover again, and the loop kernel
    for env_id, sizes in (("DartHopper-v1", (1536, 2048, 2560, 3072, 8192, 10240, 12288)), ("DartWalker2d-v1", (1536, 2048, 2560, 3072, 8192, 10240, 12288)),
                          ("DartHalfCheetah-v1", (8192, 10240, 12288, 16384)), ("DartSnake7Link-v1", (512, 768, 8192, 12288, 16384))):
        for n in sizes:
            for v in (("2", "3") if n <= 3072 else (("0", "3") if env_id != "DartHalfCheetah-v1" else ("0", "2", "3"))):
                cfgs.append((env_id, n, "128", v))
    for env_id, n in (("DartHopper-v1", 4096), ("DartWalker2d-v1", 4096), ("DartSnake7Link-v1", 4096), ("DartHalfCheetah-v1", 4096)):
        cfgs.append((env_id, n, "128", "1"))
elif mode == "r2cf":   # the contact-free envs (fused task kinds on the loop kernels)
    for env_id in ("DartCartPole-v1", "DartCartPoleSwingUp-v1", "DartDoubleInvertedPendulumEnv-v1", "DartReacher-v1"):
        for n in (4096, 65536):
            cfgs.append((env_id, n, "128", "-1"))
elif mode == "r2s":   # quick A/B set: the automatic choice at the bench sizes
    for env_id, n in (("DartHopper-v1", 4096), ("DartHopper-v1", 65536), ("DartWalker2d-v1", 4096), ("DartWalker2d-v1", 16384),
                      ("DartHalfCheetah-v1", 8192), ("DartHalfCheetah-v1", 16384), ("DartSnake7Link-v1", 4096)):
        cfgs.append((env_id, n, "128", "-1"))
    for env_id, n in (("DartWalker2d-v1", 16384), ("DartSnake7Link-v1", 4096)):
        cfgs.append((env_id, n, "128", "-1", "30"))
elif mode == "r2y":
    for env_id, n in (("DartHopper-v1", 1024), ("DartHopper-v1", 16384), ("DartWalker2d-v1", 4096), ("DartWalker2d-v1", 16384)):
        cfgs.append((env_id, n, "128", "-1"))
elif mode == "r2quadonly":   # register-cap builds of the quad form
    for env_id, n in (("DartWalker2d-v1", 16384), ("DartHalfCheetah-v1", 16384), ("DartHopper-v1", 65536), ("DartHopper-v1", 16384)):
        for pgs in ("", "30"):
            cfgs.append((env_id, n, "128", "3", pgs))
elif mode == "lcp":
    for v in ("0", "1"):
        for pgs in ("", "1"):
            for n in (4096, 65536):
                cfgs.append(("DartHopper-v1", n, "32" if n == 4096 else "128", v, pgs))
            cfgs.append(("DartHalfCheetah-v1", 16384, "128", v, pgs))
elif mode == "wpw":
    for env_id, n in (("DartHopper-v1", 4096), ("DartWalker2d-v1", 16384), ("DartHalfCheetah-v1", 16384),
                      ("DartSnake7Link-v1", 4096), ("DartHopper-v1", 16384), ("DartHopper-v1", 65536)):
        for wpw in ("32", "16", "8", "4", "2"):
            if n // int(wpw) > 16384:
                continue
            for b in ("32", "128"):
                cfgs.append((env_id, n, b, "0", "", wpw))
else:
    for b in ("32", "64", "128", "256"):
        cfgs.append(("DartHopper-v1", 4096, b, "0"))
for cfg in cfgs:
    env_id, n, b, v = cfg[:4]
    env = dict(os.environ, DARTB_BLOCK=b, DARTB_VARIANT=v, DART_ENV_NO_REFERENCE="1")
    if len(cfg) > 4 and cfg[4]:
        env["SWEEP_PGS"] = cfg[4]
    if len(cfg) > 5 and cfg[5]:
        env["DARTB_WPW"] = cfg[5]
    if len(cfg) > 6:
        env["DARTB_COOP_WARPS"] = cfg[6]
    r = subprocess.run([sys.executable, "-c", CODE, env_id, str(n)], env=env, capture_output=True, text=True)
    print("variant=%s " % v + (r.stdout.strip() or r.stderr.strip()[-300:]), flush=True)
