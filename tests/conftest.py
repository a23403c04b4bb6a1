import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "xd-tts_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(scope="session")
def built():
    """Build (or reuse) the CUDA library and the checkers once per session."""
    import __graft_entry__ as g

    g.build_cuda()
    g.build_checkers()
    return g


def rel_rms(a, b):
    """RMS difference relative to the peak of the reference signal."""
    import numpy as np

    s = max(float(np.abs(b).max()), 1e-30)
    return float(np.sqrt(np.mean(((np.asarray(a, np.float64) - np.asarray(b, np.float64)) / s) ** 2)))
