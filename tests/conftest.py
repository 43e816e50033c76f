import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def golden():
    return np.load(os.path.join(ROOT, "tests", "golden", "hotpath_golden.npz"))


@pytest.fixture(scope="session")
def cuda():
    """torch with a usable CUDA device, plus the loaded C-ABI library (fails loudly if absent)."""
    import torch
    assert torch.cuda.is_available(), "GPU test selected but no CUDA device is visible"
    import newsched_b200 as nb
    nb.lib()
    return torch


def cplx(rng, n):
    return (rng.uniform(-1, 1, n) + 1j * rng.uniform(-1, 1, n)).astype(np.complex64)
