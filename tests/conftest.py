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


def load_fixture(name):
    """Return (scipy csc genes x cells, dict of extra arrays) for a committed fixture."""
    import scipy.sparse as sp
    z = np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)
    dim = tuple(int(v) for v in z["dim"])
    m = sp.csc_matrix((z["x"].astype(np.float64), z["i"].astype(np.int32), z["p"].astype(np.int32)), shape=dim)
    extra = {k: z[k] for k in z.files if k not in ("p", "i", "x", "dim")}
    return m, extra


@pytest.fixture(scope="session")
def a549():
    return load_fixture("a549")


@pytest.fixture(scope="session")
def worm_l2():
    return load_fixture("worm_l2")


@pytest.fixture(scope="session")
def worm_embryo():
    return load_fixture("worm_embryo")
