"""pytest configuration: `-m "not gpu"` runs here (no GPU: oracle vs golden vectors, host logic, C-ABI
exports); `-m gpu` runs on a B200 and is the parity suite proper (CUDA path vs oracle, through the
C-ABI library)."""
import os
import sys
import warnings

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

warnings.filterwarnings("ignore", category=DeprecationWarning)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (sm_100a) GPU; run through gpurun")


@pytest.fixture(scope="session")
def golden():
    import numpy as np
    return np.load(os.path.join(ROOT, "tests", "golden", "bilinear_golden.npz"))


@pytest.fixture(scope="session")
def cuda_lib():
    """The C-ABI library on a GPU box.  Missing library or GPU is a FAILURE for gpu tests, never a skip:
    the product path has no fallback."""
    import torch
    assert torch.cuda.is_available(), "gpu-marked test started without a CUDA device"
    from bilinear_b200 import _lib
    return _lib.lib()
