"""The GPU parity suite again, on the LANE-COOPERATIVE kernels (planar_coop.cuh, kernel variant 2:
8 / 16 lanes step one world together).  Same oracle, same goldens, same stated tolerances as
tests/test_gpu_parity.py; plus a closed-loop comparison of the two kernel forms."""
import numpy as np
import pytest
import torch

from dart_env_b200.tasks import SPECS
import test_gpu_parity as P
from test_gpu_parity import (  # noqa: F401  (collected here, run with VARIANT = 2)
    test_substep_fp64_matches_oracle_tightly, test_substep_fp32_within_stated_tolerance,
    test_env_step_matches_reference_task_layer, test_reset_noise_bit_exact_and_sharding_independent,
    test_pgs_mode_matches_oracle_pgs, test_full_size_properties_hopper_4096, test_time_limit_truncation,
    test_gym_surface_single_env_types, test_contacts_readback_walker, test_full_size_properties_other_configs,
    test_rollout_statistics_match_oracle, test_pydart2_shaped_views_match_the_oracle,
    test_step_is_cuda_graph_capturable)

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _cooperative_kernel():
    P.VARIANT = 2
    yield
    P.VARIANT = 0


@pytest.mark.parametrize("env_id", list(SPECS))
def test_closed_loop_equals_per_thread_kernel_fp64(models, env_id):
    """fp64 engines of both kernel forms, same seed and actions, auto-reset on: identical done flags and
    reset draws, states equal to rounding for the first env steps (before contact chaos amplifies)."""
    spec = SPECS[env_id]
    n, dev = 256, torch.device("cuda", 0)
    gen = torch.Generator(device=dev); gen.manual_seed(3)
    acts = [torch.rand((n, spec.task.n_act), generator=gen, device=dev) * 2 - 1 for _ in range(12)]
    out = []
    for variant in (0, 2):
        P.VARIANT = variant
        eng = P._engine(models, env_id, n, seed=4, f64=True)
        obs = eng.reset()
        rew = torch.empty((n,), dtype=torch.float32, device=dev); done = torch.empty((n,), dtype=torch.uint8, device=dev)
        hist = []
        for a in acts:
            eng.step(a, obs, rew, done, True)
            q, dq = eng.get_state(torch.float64)
            hist.append((obs.clone(), rew.clone(), done.clone(), q.clone(), dq.clone()))
        assert ("coop:" in eng.kernel_name) == (variant == 2)
        eng.close()
        out.append(hist)
    P.VARIANT = 2
    for t, ((o0, r0, d0, q0, v0), (o2, r2, d2, q2, v2)) in enumerate(zip(*out)):
        same = (d0 == d2)
        assert same.float().mean() > 0.99
        if t < 4:
            m = same.cpu().numpy()
            assert np.allclose(q0.cpu().numpy()[m], q2.cpu().numpy()[m], rtol=1e-6, atol=1e-7)
            assert np.allclose(r0.cpu().numpy()[m], r2.cpu().numpy()[m], rtol=1e-4, atol=1e-4)
            assert np.allclose(o0.cpu().numpy()[m], o2.cpu().numpy()[m], rtol=1e-4, atol=1e-5)


def test_automatic_kernel_choice_by_batch_size(models):
    """Small batches run the cooperative form, mid-size ones the quad form, large ones the per-thread form
    (dartb.cu::lower_into); each can be forced.  Bit-exact shard == batch holds within one form; across forms
    trajectories agree to fp32 tolerance and the reset noise (keyed by global world id) is identical."""
    P.VARIANT = None
    try:
        for env_id, sizes in (("DartHopper-v1", ((2048, "coop:"), (4096, "quad:"), (16384, "static:"))),
                              ("DartWalker2d-v1", ((2048, "coop:"), (8192, "quad:"), (16384, "static:"))),
                              ("DartHalfCheetah-v1", ((4096, "coop:"), (12288, "quad:"), (16384, "static:"))),
                              ("DartSnake7Link-v1", ((256, "coop:"), (4096, "quad:"), (32768, "static:")))):
            ref = None
            for n, tag in sizes:
                e = P._engine(models, env_id, n, seed=1)
                assert tag in e.kernel_name, (env_id, n, e.kernel_name)
                e.reset()
                q, _ = e.get_state(torch.float64)
                if ref is None:
                    ref = q
                else:
                    assert torch.equal(ref, q[:ref.shape[0]])      # same reset draws whichever kernel runs
                e.close()
        # PGS mode skips the quad form (its PGS gathers A onto every lane)
        e = P._engine(models, "DartWalker2d-v1", 8192, seed=1)
        e.set_lcp(1, 30)
        assert "static:" in e.kernel_name
        e.close()
    finally:
        P.VARIANT = 2
