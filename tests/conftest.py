import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def built():
    """Make sure the native libraries exist (nvcc cross-compiles without a GPU)."""
    from toucan_b200 import build

    build.build_all()
    return True


@pytest.fixture(scope="session")
def cuda(built):
    import torch

    if not torch.cuda.is_available():
        pytest.fail("GPU test selected but no CUDA device is visible (there is no CPU fallback)")
    from toucan_b200 import capi

    capi.check(capi.lib().tc_set_device(0), "tc_set_device")
    return torch.device("cuda:0")
