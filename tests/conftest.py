import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    """`-m gpu` tests need a CUDA device and the built extension: on a machine without them they are SKIPPED
    (a plain `pytest` stays green), never silently run on something else."""
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="needs a CUDA device (B200 box): the engine has no CPU fallback")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def models():
    from dart_env_b200.skel import load_model
    from dart_env_b200.tasks import SPECS
    out = {}
    for k, spec in SPECS.items():
        m = load_model(spec.skel, spec.dt)
        m.enforce_limits()
        if spec.friction_all is not None:
            for b in m.bodies:
                b.friction_coeff = spec.friction_all
        out[k] = m
    return out
