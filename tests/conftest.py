import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


def load_golden(name):
    with np.load(os.path.join(GOLDEN, name + ".npz")) as z:
        return {k: z[k] for k in z.files}


@pytest.fixture(scope="session")
def golden_ops():
    return load_golden("ops")


def relerr(a, b):
    """||a-b|| / max(||a||,||b||) -- the reference's own gradient-check norm
    (gradientcheck/gradientcheck.py:17-18)."""
    a = np.asarray(a, dtype=np.float64).ravel()
    b = np.asarray(b, dtype=np.float64).ravel()
    d = max(np.linalg.norm(a), np.linalg.norm(b))
    return 0.0 if d == 0 else float(np.linalg.norm(a - b) / d)
