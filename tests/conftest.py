import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    from oracle.pyoracle import OracleLib
    return OracleLib()


@pytest.fixture(scope="session")
def ref():
    """The unmodified reference (oracle/_ref/libcafe_ref.so); tests needing it are skipped when it was not built."""
    from oracle import pyoracle
    if os.path.isdir("/root/reference/src"):
        pyoracle.build_ref()
    if not pyoracle.have_ref():
        pytest.skip("oracle/_ref/libcafe_ref.so not built (reference sources absent)")
    return pyoracle.RefLib()


@pytest.fixture(scope="session")
def golden():
    return {n: np.load(os.path.join(GOLDEN, n + ".npz")) for n in ("mammals", "hymenoptera", "matrices", "small")}


def ulp_distance(a, b):
    a = np.ascontiguousarray(a, dtype=np.float64)
    b = np.ascontiguousarray(b, dtype=np.float64)
    return np.abs(a.view(np.int64) - b.view(np.int64))


def max_rel(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    d = np.abs(a - b)
    scale = np.where(b != 0, np.abs(b), 1.0)
    return float(np.max(d / scale)) if d.size else 0.0
