"""Test configuration: markers, path setup and shared fixtures."""
from __future__ import annotations

import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "slow: full BASELINE sizes")


@pytest.fixture(scope="session")
def backend():
    """The libyamb200 backend on cuda:0; GPU tests fail loudly if it cannot be created."""
    from yamimageprocessor_b200.backend import get_backend

    return get_backend(0)


@pytest.fixture()
def rng():
    return np.random.default_rng(1234)
