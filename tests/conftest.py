import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def cuda_device():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


@pytest.fixture(autouse=True)
def _tanh_after_each_test():
    """The binding's default activation is a module variable (engine.drp_set_activation; the library itself takes the id per launch): tests that select
    another one must not leak it into kernel-level tests that expect tanh epilogues."""
    yield
    mod = sys.modules.get("pyrddlgym_jax_b200.core.engine")
    if mod is not None and getattr(mod, "_LIB", None) is not None:
        mod.drp_set_activation("tanh")
