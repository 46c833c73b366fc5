import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.dirname(os.path.abspath(__file__))):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


@pytest.fixture(scope="session")
def kats():
    import helpers
    return helpers.kats()


@pytest.fixture(scope="session")
def oracle():
    from oracle import binding
    binding.build()
    return binding


@pytest.fixture(scope="session")
def capi():
    """The product's C ABI through ctypes (built on demand; raises if it cannot be built — no fallback)."""
    from kitti_motion_compensation_b200 import build, capi as _capi
    build.build()
    _capi.lib()
    return _capi


@pytest.fixture(scope="session")
def cuda():
    """torch, used for device memory and streams only."""
    import torch
    if not torch.cuda.is_available():
        pytest.fail("a -m gpu test ran without a CUDA device: there is no CPU fallback for the deskew path")
    torch.cuda.set_device(0)
    return torch
