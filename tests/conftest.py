"""pytest configuration: markers and shared fixtures.

`-m "not gpu"`: oracle vs. golden fixtures, host logic, C-ABI symbol check — no GPU needed.
`-m gpu`      : parity of the CUDA path (through the C-ABI) against the oracle and the fixtures.
"""
from __future__ import annotations

import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

GOLDEN = ROOT / "tests" / "golden"

# north_star tolerance: fp32 rtol 1e-5 / atol 1e-6 against the reference on identical inputs
RTOL = 1e-5
ATOL = 1e-6


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def golden():
    def load(name: str):
        return np.load(GOLDEN / name)
    return load


@pytest.fixture(scope="session")
def cuda_lib():
    """The loaded C-ABI library; GPU tests fail loudly if it is missing."""
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from beyond_deep_ensembles_b200 import _lib
    return _lib.get()
