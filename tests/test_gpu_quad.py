"""The GPU parity suite again, on the QUAD form of the per-thread kernels (kernel variant 3: four lanes per world share the
constraint phase, planar_kernels.cuh::substep<..., G = 4>).  Same oracle, same goldens, same stated tolerances as
tests/test_gpu_parity.py; plus a closed-loop comparison with the one-world-per-thread form."""
import numpy as np
import pytest
import torch

from dart_env_b200.tasks import SPECS
import test_gpu_parity as P
from test_gpu_parity import (  # noqa: F401  (collected here, run with VARIANT = 3)
    test_substep_fp64_matches_oracle_tightly, test_substep_fp32_within_stated_tolerance,
    test_env_step_matches_reference_task_layer, test_reset_noise_bit_exact_and_sharding_independent,
    test_pgs_mode_matches_oracle_pgs, test_full_size_properties_hopper_4096, test_time_limit_truncation,
    test_contacts_readback_walker, test_full_size_properties_other_configs, test_rollout_statistics_match_oracle,
    test_step_is_cuda_graph_capturable)

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _quad_kernel():
    P.VARIANT = 3
    yield
    P.VARIANT = 0


@pytest.mark.parametrize("env_id", list(SPECS))
def test_closed_loop_equals_per_thread_kernel(models, env_id):
    """Both forms run the same K1-K4 arithmetic; the constraint phase differs only in how A = J M^-1 J^T is rounded
    (full rows per lane vs a mirrored lower triangle) and in the LCP's linear algebra.  fp64 engines, same seed and
    actions, auto-reset on: identical done flags and reset draws, states equal to rounding."""
    spec = SPECS[env_id]
    n, dev = 200, torch.device("cuda", 0)      # not a multiple of 8: the last warp is ragged
    gen = torch.Generator(device=dev); gen.manual_seed(3)
    acts = [torch.rand((n, spec.task.n_act), generator=gen, device=dev) * 2 - 1 for _ in range(12)]
    out = []
    for variant in (0, 3):
        P.VARIANT = variant
        eng = P._engine(models, env_id, n, seed=4, f64=True)
        obs = eng.reset()
        rew = torch.empty((n,), dtype=torch.float32, device=dev); done = torch.empty((n,), dtype=torch.uint8, device=dev)
        hist = []
        for a in acts:
            eng.step(a, obs, rew, done, True)
            q, dq = eng.get_state(torch.float64)
            hist.append((obs.clone(), rew.clone(), done.clone(), q.clone(), dq.clone()))
        assert ("quad:" in eng.kernel_name) == (variant == 3)
        eng.close()
        out.append(hist)
    P.VARIANT = 3
    for t, ((o0, r0, d0, q0, v0), (o3, r3, d3, q3, v3)) in enumerate(zip(*out)):
        same = (d0 == d3)
        assert same.float().mean() > 0.99
        if t < 4:
            m = same.cpu().numpy()
            assert np.allclose(q0.cpu().numpy()[m], q3.cpu().numpy()[m], rtol=1e-7, atol=1e-8)
            assert np.allclose(o0.cpu().numpy()[m], o3.cpu().numpy()[m], rtol=1e-5, atol=1e-6)


def test_world_independent_of_warp_neighbours(models):
    """a warp picks ONE constraint-phase class from its largest row count: shard == batch, bit for bit"""
    env_id = "DartWalker2d-v1"
    spec = SPECS[env_id]
    dev = torch.device("cuda", 0)
    gen = torch.Generator(device=dev); gen.manual_seed(9)
    full = P._engine(models, env_id, 64, seed=2)
    full.reset()
    obs = torch.empty((64, 17), device=dev); rew = torch.empty((64,), device=dev); done = torch.empty((64,), dtype=torch.uint8, device=dev)
    acts = [torch.rand((64, 6), generator=gen, device=dev) * 2 - 1 for _ in range(40)]
    for a in acts:
        full.step(a, obs, rew, done, True)
    qf, dqf = full.get_state(torch.float64)
    # the same worlds 24..39 alone (different warp neighbours, different warp-level row-count maxima)
    part = P._engine(models, env_id, 16, seed=2, world_offset=24)
    part.reset()
    o2 = torch.empty((16, 17), device=dev); r2 = torch.empty((16,), device=dev); d2 = torch.empty((16,), dtype=torch.uint8, device=dev)
    for a in acts:
        part.step(a[24:40].contiguous(), o2, r2, d2, True)
    qp, dqp = part.get_state(torch.float64)
    assert torch.equal(qf[24:40], qp) and torch.equal(dqf[24:40], dqp)
    full.close(); part.close()
