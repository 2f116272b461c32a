"""Prints a digest of 20 env steps of the lane-cooperative kernels (Hopper 4096 worlds, Walker2d 1024 worlds) so that
runs under different DARTB_COOP_TMA modes (read once per process) can be compared bit for bit.
Used by tests/test_gpu_round2.py::test_tma_staged_prologue_is_bit_identical."""
import hashlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from dart_env_b200.envs import make  # noqa: E402

h = hashlib.sha256()
for env_id, n in (("DartHopper-v1", 4096), ("DartWalker2d-v1", 1024), ("DartSnake7Link-v1", 512)):
    env = make(env_id, num_envs=n, output="torch", seed=3, batched=True, kernel_variant=2)
    assert "coop:" in env.engine.kernel_name
    env.reset()
    gen = torch.Generator(device="cuda"); gen.manual_seed(5)
    for t in range(20):
        a = torch.rand((n, env.act_dim), generator=gen, device="cuda") * 2 - 1
        o, r, d, _ = env.step(a)
        h.update(o.cpu().numpy().tobytes()); h.update(r.cpu().numpy().tobytes()); h.update(d.cpu().numpy().tobytes())
    q, dq = env.engine.get_state(torch.float64)
    h.update(q.cpu().numpy().tobytes()); h.update(dq.cpu().numpy().tobytes())
    env.close()
print("digest", h.hexdigest())
