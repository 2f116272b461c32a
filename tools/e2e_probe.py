"""Developer probe (under gpurun): where the end-to-end step time goes."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from dart_env_b200.envs import make

n = 4096
env = make("DartHopper-v1", num_envs=n, output="numpy", seed=0, batched=True)
env.reset()
rng = np.random.RandomState(0)
acts = [rng.uniform(-1, 1, (n, 3)).astype(np.float32) for _ in range(16)]
eng = env.engine
flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device="cuda")


def timeit(fn, K=300, do_flush=False):
    for i in range(20):
        fn(i)
    torch.cuda.synchronize()
    tot = 0.0
    for i in range(K):
        if do_flush:
            flush.fill_(float(i & 1))
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        fn(i)
        tot += time.perf_counter() - t0
    return tot / K * 1e6


obs, rew, done = env._n_obs, env._n_rew, env._n_done
keep = [None]
def hold(i):
    keep[0] = env.step(acts[i % 16])     # the caller keeps the last result: the pool alternates between two slots
print("env.step(numpy), result held warm %.1f us   flushed %.1f us" % (timeit(hold), timeit(hold, do_flush=True)))
print("env.step(numpy)            warm %.1f us   flushed %.1f us" % (timeit(lambda i: env.step(acts[i % 16])), timeit(lambda i: env.step(acts[i % 16]), do_flush=True)))
print("engine.step_host           warm %.1f us   flushed %.1f us" % (timeit(lambda i: eng.step_host(acts[i % 16], obs, rew, done, True)), timeit(lambda i: eng.step_host(acts[i % 16], obs, rew, done, True), do_flush=True)))
pg = np.empty_like(obs); pr = np.empty_like(rew); pd = np.empty_like(done)
print("engine.step_host pageable  warm %.1f us" % timeit(lambda i: eng.step_host(acts[i % 16], pg, pr, pd, True)))
# device-only launch + sync
a_dev = [torch.tensor(a, device="cuda") for a in acts]
def dev(i):
    eng.step(a_dev[i % 16], env._obs, env._rew, env._done, True)
    torch.cuda.synchronize()
print("engine.step (device) + sync warm %.1f us   flushed %.1f us" % (timeit(dev), timeit(dev, do_flush=True)))
def empty(i):
    torch.cuda.synchronize()
print("bare synchronize            %.1f us" % timeit(empty))
h = torch.empty((n, 3), dtype=torch.float32).pin_memory(); d = torch.empty((n, 3), device="cuda")
ho = torch.empty((n, 11), dtype=torch.float32).pin_memory(); do_ = torch.empty((n, 11), device="cuda")
def copies(i):
    d.copy_(h, non_blocking=True); ho.copy_(do_, non_blocking=True); torch.cuda.synchronize()
print("H2D 48KB + D2H 180KB + sync %.1f us" % timeit(copies))
