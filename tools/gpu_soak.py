"""Developer soak (under gpurun): long random-action rollouts on every kernel form (2 lane-cooperative, 3 quad, 0 one
world per thread); everything must stay finite and the forms must agree statistically (episode length, reward)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from dart_env_b200.envs import make

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 1500
for env_id, n in (("DartHopper-v1", 4096), ("DartWalker2d-v1", 4096), ("DartHalfCheetah-v1", 4096), ("DartSnake7Link-v1", 2048)):
    res = {}
    for variant in (2, 3, 0):
        env = make(env_id, num_envs=n, output="torch", seed=11, batched=True, kernel_variant=variant)
        env.reset()
        dev = env.engine.device
        gen = torch.Generator(device=dev); gen.manual_seed(3)
        nact = env.engine.n_act
        bad = 0; tot_r = 0.0; tot_d = 0; amp = 1.0
        for t in range(steps):
            if t == steps // 2:
                amp = 3.0   # out-of-range actions (clamped by the task layer) for the second half
            a = (torch.rand((n, nact), generator=gen, device=dev) * 2 - 1) * amp
            obs, rew, done, _ = env.step(a)
            if t % 50 == 0 or t == steps - 1:
                bad += int((~torch.isfinite(obs)).sum()) + int((~torch.isfinite(rew)).sum())
            tot_r += float(rew.double().mean()); tot_d += int(done.sum())
        q, dq = env.engine.get_state()
        bad += int((~torch.isfinite(q)).sum()) + int((~torch.isfinite(dq)).sum())
        res[variant] = (n * steps / max(tot_d, 1), tot_r / steps, bad, env.engine.kernel_name)
        env.close()
    print("%-20s coop: len %.1f rew %.3f nonfinite %d | quad: len %.1f rew %.3f nonfinite %d | per-thread: len %.1f rew %.3f nonfinite %d   [%s | %s | %s]"
          % (env_id, res[2][0], res[2][1], res[2][2], res[3][0], res[3][1], res[3][2], res[0][0], res[0][1], res[0][2], res[2][3], res[3][3], res[0][3]), flush=True)
