"""Pins the jaccard_coeff restatement (oracle/monocle3_ref.py) to the REFERENCE'S OWN native code:
oracle/_ref/libm3ref.so is src/clustering.cpp compiled unmodified from the reference tree against a
shim of the Rcpp containers (oracle/Makefile, built by __graft_entry__.build()).  Skipped when the
library has not been built (no reference tree and no prebuilt copy)."""
import ctypes as C
import os

import numpy as np
import pytest

from oracle import monocle3_ref as ref

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "oracle", "_ref", "libm3ref.so")


def native_jaccard(idx, weight):
    lib = C.CDLL(LIB)
    a = np.asfortranarray(idx, dtype=np.float64)
    n, k = a.shape
    out = np.empty((3, n * k))
    dp = C.POINTER(C.c_double)
    lib.m3ref_jaccard_coeff.argtypes = [dp, C.c_int, C.c_int, C.c_int, dp]
    lib.m3ref_jaccard_coeff(a.ctypes.data_as(dp), n, k, int(bool(weight)), out.ctypes.data_as(dp))
    return out.T


needs_lib = pytest.mark.skipif(not os.path.exists(LIB), reason="oracle/_ref/libm3ref.so not built (make -C oracle)")


@needs_lib
@pytest.mark.parametrize("n,k", [(3, 2), (50, 5), (400, 20), (64, 64), (7, 1)])
def test_restatement_equals_reference_native(n, k):
    rng = np.random.default_rng(100 * n + k)
    idx = rng.integers(1, n + 1, (n, k))
    idx[::5, : max(1, k // 2)] = idx[::5, :1]  # repeated ids inside a row
    for weight in (True, False):
        got = ref.jaccard_coeff(idx, weight)
        exp = native_jaccard(idx, weight)
        assert got.shape == exp.shape
        assert np.array_equal(got, exp)  # ids, order of the rows and every weight bit for bit


@needs_lib
def test_reference_native_hand_worked():
    """The hand-worked case of tests/test_oracle_aggregate.py, on the reference's code itself."""
    w = native_jaccard(np.array([[2, 3], [1, 3], [1, 2]]), True)
    assert np.array_equal(w[:, 2], np.ones(6))
    idx = np.array([[3, 4], [3, 4], [4, 4], [1, 2]])
    w = native_jaccard(idx, True)
    exp = {(1, 3): 1 / 3, (1, 4): 1.0, (2, 3): 1 / 3, (2, 4): 1.0, (3, 4): 1.0, (4, 1): 1.0, (4, 2): 1.0}
    for r in range(8):
        assert w[r, 2] == exp[(int(w[r, 0]), int(w[r, 1]))]
