"""2+ GPUs under torchrun: the observation all-gather fused into the step kernel (parallel.FusedObsGather: peer stores over
NVLink into symmetric memory + one device-side barrier) against NCCL all_gather_into_tensor on the same trajectories —
bit-identical gathered batches — and the time of both per step.  Exit code 0 = identical."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist

from dart_env_b200.parallel import FusedObsGather, gather_batch, make_sharded

rank = int(os.environ["RANK"]); local = int(os.environ["LOCAL_RANK"]); ws = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
env_id = sys.argv[1] if len(sys.argv) > 1 else "DartHalfCheetah-v1"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 16384
ok = True
outs = {}
for mode in ("nccl", "fused"):
    env = make_sharded(env_id, n * ws, output="torch", seed=5)
    env.reset()
    gen = torch.Generator(device=dev); gen.manual_seed(11 + rank)
    acts = [torch.rand((n, env.act_dim), generator=gen, device=dev) * 2 - 1 for _ in range(16)]
    g = FusedObsGather(env) if mode == "fused" else None
    buf = torch.empty((ws * n, env.obs_dim), dtype=torch.float32, device=dev)

    def one(i):
        if g is not None:
            return g.step(acts[i % 16])
        env.engine.step(acts[i % 16], env._obs, env._rew, env._done, True)
        return gather_batch(env._obs, buf)

    hist = []
    for i in range(12):
        hist.append(one(i).clone())
    for i in range(30):
        one(i)
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    K = 200
    e0.record()
    for i in range(K):
        one(i)
    e1.record(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / K * 1e3], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    outs[mode] = (hist, float(t.item()))
    if g is not None:
        g.close()
    env.close()
for a, b in zip(outs["nccl"][0], outs["fused"][0]):
    ok = ok and torch.equal(a, b)
flag = torch.tensor([1 if ok else 0], device=dev)
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print("fused gather == NCCL all_gather (12 steps, %d ranks x %d worlds of %s): %s" % (ws, n, env_id, bool(flag.item())))
    print("us per step incl. gather: NCCL all_gather_into_tensor %.1f   fused peer stores + barrier %.1f" % (outs["nccl"][1], outs["fused"][1]))
dist.barrier(); dist.destroy_process_group()
sys.exit(0 if flag.item() else 1)
