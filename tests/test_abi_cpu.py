"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol
include/m3b.h declares, refuses to run without a GPU (no fallback), and its host-only entry
point (size-factor finish) is bit-exact against the oracle."""
import ctypes
import os
import re

import numpy as np
import pytest

from oracle import monocle3_ref as ref

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    syms = set()
    for fn in os.listdir(os.path.join(ROOT, "include")):
        if fn.endswith(".h"):
            txt = open(os.path.join(ROOT, "include", fn)).read()
            txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
            syms |= set(re.findall(r"\b(m3b_[a-z0-9_]+)\s*\(", txt))
    return syms


@pytest.fixture(scope="module")
def lib():
    from monocle3_b200 import _lib
    return _lib.load()


def test_library_exports_every_declared_symbol(lib):
    from monocle3_b200 import _lib
    declared = _declared_symbols()
    assert len(declared) >= 30
    for s in declared:
        assert hasattr(lib, s), "libm3b.so does not export %s" % s
    # the ctypes binding covers exactly the header
    assert set(_lib.SIGNATURES) == declared


def test_abi_version_and_status_strings(lib):
    assert lib.m3b_abi_version() == 3
    # the first codes mirror irlba's C-path errors
    assert lib.m3b_status_string(-2) == b"didn't converge"
    assert lib.m3b_status_string(-4) == b"starting vector near the null space"
    assert lib.m3b_status_string(-5) == b"linear dependency encountered"


def test_no_cpu_fallback(lib):
    if lib.m3b_device_count() > 0:
        pytest.skip("a GPU is visible")
    h = ctypes.c_void_p()
    assert lib.m3b_create(0, 0, ctypes.byref(h)) == -13  # M3B_E_NODEVICE
    import monocle3_b200 as m3
    from monocle3_b200._lib import M3BError
    m3.set_context(None)
    with pytest.raises(M3BError, match="no CPU fallback"):
        m3.load_a549()  # new_cell_data_set -> estimate_size_factors needs the device


@pytest.mark.parametrize("name", ["a549", "worm_l2", "worm_embryo"])
@pytest.mark.parametrize("method", ["mean-geometric-mean-total", "mean-geometric-mean-log-total"])
def test_size_factor_finish_is_bit_exact(request, name, method):
    """m3b_size_factors_from_totals (host, long double like R's mean()) vs the oracle."""
    from monocle3_b200 import engine, _lib as L
    counts, _ = request.getfixturevalue(name)
    sf_o, tot = ref.estimate_sf_sparse(counts, method=method)
    code = L.SF_TOTAL if method.endswith("-total") and "log" not in method else L.SF_LOG_TOTAL
    sf, any_zero = engine.size_factors_from_totals(tot, code)
    assert not any_zero
    assert np.array_equal(sf, sf_o)


def test_size_factor_finish_edge_cases():
    from monocle3_b200 import engine
    sf, any_zero = engine.size_factors_from_totals(np.array([3.0, 0.0, 5.0]))
    assert any_zero                      # zero-read guard, R/utils.R:32-39
    sf1, _ = engine.size_factors_from_totals(np.array([7.0]))
    assert sf1[0] == 7.0 / np.exp(np.log(7.0))  # a single cell: total / exp(mean(log(total))), in libm arithmetic
    sf2, _ = engine.size_factors_from_totals(np.zeros(0))
    assert sf2.size == 0                 # empty input


def test_host_argument_checks_need_no_device():
    """preprocess_cds validates its arguments like R/preprocess_cds.R:73-98 before touching the GPU."""
    import scipy.sparse as sp
    import monocle3_b200 as m3
    cds = m3.CellDataSet(sp.csc_matrix(np.ones((4, 5))), gene_names=list("abcd"))
    with pytest.raises(ValueError, match="method must be one of"):
        m3.preprocess_cds(cds, method="tSNE")
    with pytest.raises(ValueError, match="norm_method must be one of"):
        m3.preprocess_cds(cds, norm_method="sqrt")
    with pytest.raises(ValueError, match="num_dim"):
        m3.preprocess_cds(cds, num_dim=0)
    with pytest.raises(ValueError, match="use_genes"):
        m3.preprocess_cds(cds, use_genes=["zz"])
    with pytest.raises(ValueError, match="estimate_size_factors"):
        m3.preprocess_cds(cds)          # no Size_Factor column yet
    with pytest.raises(NotImplementedError):
        m3.set_size_factors(cds, np.ones(5))
        m3.preprocess_cds(cds, build_nn_index=True)
    m3.set_size_factors(cds, None)
    with pytest.raises(ValueError, match="reduction_method"):
        m3.preprocess_transform(cds, reduction_method="UMAP")
    m3.set_size_factors(cds, np.ones(5))
    with pytest.raises(ValueError, match="not in the model object"):
        m3.preprocess_transform(cds)


def test_product_rng_matches_r_and_oracle():
    from monocle3_b200.r_rng import rnorm_after_seed
    from oracle.r_rng import RRandom
    np.testing.assert_allclose(rnorm_after_seed(2016, 5),
                               [-0.91474184, 1.00124785, -0.05642291, 0.29664516, -2.79147086], atol=5e-9)
    assert np.array_equal(rnorm_after_seed(2016, 1500), RRandom(2016).rnorm(1500))
    # one stream, several draws (start vector, then irlba's replacement vectors)
    from monocle3_b200.r_rng import RStream
    a, b = RStream(7), RRandom(7)
    assert np.array_equal(a.rnorm(11), b.rnorm(11)) and np.array_equal(a.rnorm(5), b.rnorm(5))


def test_csc_blocks_host_logic():
    """monocle3_b200.CSCBlocks (a matrix beyond the dgCMatrix limit as consecutive cell blocks): shape / nnz
    bookkeeping and the argument checks need no device."""
    import scipy.sparse as sp
    import monocle3_b200 as m3
    rng = np.random.default_rng(0)
    a = sp.random(30, 7, density=0.3, format="csc", random_state=rng)
    b = sp.random(30, 5, density=0.3, format="csc", random_state=rng)
    blk = m3.CSCBlocks([a, b])
    assert blk.shape == (30, 12) and blk.nnz == a.nnz + b.nnz
    cds = m3.CellDataSet(blk)                       # accepted as it is (no coercion to one csc_matrix)
    assert cds.counts is blk and cds.n_cells_total == 12 and len(cds.cell_names) == 12
    with pytest.raises(ValueError, match="same number of genes"):
        m3.CSCBlocks([a, sp.random(31, 5, density=0.3, format="csc", random_state=rng)])
    with pytest.raises(ValueError, match="at least one block"):
        m3.CSCBlocks([])


@pytest.mark.parametrize("method", ["mean-geometric-mean-total", "mean-geometric-mean-log-total"])
def test_size_factor_finish_threads_do_not_change_the_result(method):
    """Above 65 536 cells the element-wise logs of m3b_size_factors_from_totals run on several host threads;
    the mean stays R's sequential long-double loop, so the result must still equal the oracle's bit for bit
    (600 000 cells: 8 threads)."""
    import scipy.sparse as sp
    from monocle3_b200 import engine, _lib as L
    rng = np.random.default_rng(11)
    tot = np.floor(rng.lognormal(7.5, 0.6, 600_000)) + 2.0
    counts = sp.csc_matrix(tot[None, :])            # one gene: the column sums are the totals
    sf_o, tot_o = ref.estimate_sf_sparse(counts, method=method)
    assert np.array_equal(tot_o, tot)
    code = L.SF_TOTAL if "log-total" not in method else L.SF_LOG_TOTAL
    sf, any_zero = engine.size_factors_from_totals(tot, code)
    assert not any_zero
    assert np.array_equal(sf, sf_o)
    tot[123_456] = 0.0
    assert engine.size_factors_from_totals(tot, code)[1]       # the zero is seen whichever thread owns it
