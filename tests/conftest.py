import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


@pytest.fixture(scope="session")
def cdfo_so():
    """Path of the in-tree C-ABI library; (re)built here only when nvcc is present and sources are newer."""
    from cdfo_b200.csrc import build as B
    if B.stale() and os.path.isfile(B.NVCC):
        B.build()
    assert os.path.isfile(B.OUT), "libcdfo_b200.so missing and cannot be built"
    return B.OUT


@pytest.fixture(scope="session")
def cuda_dev():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import cdfo_b200
    assert cdfo_b200._lib.lib().cdfo_device_ok(0) == 1, "device 0 is not sm_100"
    return torch.device("cuda:0")
