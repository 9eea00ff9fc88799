import os
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden():
    """Outputs of the reference itself (tests/golden/make_golden.py)."""
    return dict(np.load(ROOT / "tests" / "golden" / "reference_outputs.npz"))


@pytest.fixture(scope="session")
def golden_pm():
    """Reference outputs for the SURVEY §8(f) rows (tests/golden/make_golden_prematch.py)."""
    return dict(np.load(ROOT / "tests" / "golden" / "prematch_outputs.npz"))


def has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.fixture
def small_log():
    """candidate log of 256 slots per (row, segment) for one test: small fixtures then overflow it
    and exercise the exact-kernel fallback; the default (2048) is restored afterwards"""
    from knn_svc_b200 import _lib
    lib = _lib.load()
    _lib.check(lib.knnsvc_set_option(b"log_cap", 256), "set_option")
    yield
    _lib.check(lib.knnsvc_set_option(b"log_cap", 0), "set_option")
