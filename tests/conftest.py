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
    return {n: np.load(os.path.join(GOLDEN, n + ".npz")) for n in ("mammals", "hymenoptera", "matrices", "small", "priors")}


def random_prior(rng, max_root_family_size):
    """One of the three priors the reference can be run with (src/user_data.cpp:176-206), as the float32 table cafe_b200_set_prior
    takes: uniform, a user root distribution with holes and a table shorter than max_root_family_size, or a Poisson prior."""
    from cafe5_b200 import families as fam
    kind = int(rng.integers(0, 3))
    if kind == 0:
        return fam.uniform_prior(max_root_family_size)
    if kind == 1:
        top = int(rng.integers(max(3, max_root_family_size // 3), max_root_family_size))
        rd = {s: int(rng.choice([1, 2, 30, 700, 20000])) for s in range(1, top + 1) if rng.random() > 0.2}
        rd[top] = 5
        return fam.rootdist_prior(rd)
    return fam.poisson_prior(float(rng.uniform(0.5, max_root_family_size / 3.0)), max_root_family_size)


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
