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
