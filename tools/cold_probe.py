"""Developer probe (under gpurun): what is cold after an L2 flush?  Two engines run the same kernel
on different buffers: flush -> step A (cold code + cold data) -> step B (code warm in L2, data cold)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from dart_env_b200.envs import make

env_id = sys.argv[1] if len(sys.argv) > 1 else "DartHopper-v1"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
A = make(env_id, num_envs=n, output="torch", seed=0, batched=True)
B = make(env_id, num_envs=n, output="torch", seed=1, batched=True)
A.reset(); B.reset()
dev = A.engine.device
gen = torch.Generator(device=dev); gen.manual_seed(1)
acts = [torch.rand((n, A.engine.n_act), generator=gen, device=dev) * 2 - 1 for _ in range(16)]
flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device=dev)


def stepA(i): A.engine.step(acts[i % 16], A._obs, A._rew, A._done, True)
def stepB(i): B.engine.step(acts[(i + 5) % 16], B._obs, B._rew, B._done, True)


for i in range(30):
    stepA(i); stepB(i)
torch.cuda.synchronize()
K = 200
ev = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(K)]
for i in range(K):
    flush.fill_(float(i & 1))
    ev[i][0].record(); stepA(i); ev[i][1].record(); stepB(i); ev[i][2].record(); stepB(i + 1); ev[i][3].record()
torch.cuda.synchronize()
a = np.array([[e[0].elapsed_time(e[1]), e[1].elapsed_time(e[2]), e[2].elapsed_time(e[3])] for e in ev]) * 1e3
print("%s n=%d  after flush: A (cold code+data) %.1f us | B (warm code, cold data) %.1f us | B again (all warm) %.1f us"
      % (env_id, n, *a.mean(axis=0)))
