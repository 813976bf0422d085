import os
import sys

import numpy as np
import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "ref: needs the unmodified reference tree at /root/reference")


def golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def rel_err(a, b):
    """|a-b|_inf / max(1, |b|_inf): the contract of SURVEY 8(d) (per array)."""
    a = np.asarray(a, dtype=float)
    b = np.asarray(b, dtype=float)
    fin = np.isfinite(b)
    assert np.array_equal(np.isfinite(a), fin), "finite masks differ"
    if not fin.any():
        return 0.0
    return float(np.max(np.abs(a[fin] - b[fin])) / max(1.0, float(np.max(np.abs(b[fin])))))


@pytest.fixture(scope="session")
def fa_ref():
    """Fully-actuated reference as get_fully_actuated_ref() builds it (trajectory_generation.py:511-518)."""
    d = golden("fully_actuated_trajectory")
    u_ref = np.zeros(d["u"].shape)
    u_ref[:, 1] = d["u"][:, 1]
    return d["x"].copy(), 2.0 * u_ref, d["time"].copy()
