import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)


@pytest.fixture(scope="session")
def golden():
    return load_golden


# Float tolerance of the parity contract (SURVEY.md section 8c / BASELINE.json north_star):
# |y - y_ref| <= max(1e-5 * |y_ref|, 1e-5)   i.e. 1e-5 relative or -100 dBFS absolute.
REL_TOL = 1e-5
ABS_TOL = 1e-5


def assert_parity(y, y_ref, what=""):
    y = np.asarray(y, dtype=np.float64)
    y_ref = np.asarray(y_ref, dtype=np.float64)
    assert y.shape == y_ref.shape, f"{what}: shape {y.shape} vs {y_ref.shape}"
    err = np.abs(y - y_ref)
    bound = np.maximum(REL_TOL * np.abs(y_ref), ABS_TOL)
    bad = err > bound
    assert not bad.any(), (f"{what}: {bad.sum()} samples out of tolerance, max err "
                           f"{err.max():.3e} at {np.unravel_index(err.argmax(), err.shape)}")
    return float(err.max())
