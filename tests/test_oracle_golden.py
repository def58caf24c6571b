"""Pin the oracle (oracle/) against the reference's own golden values.

Every literal below is copied from the reference's testthat suite (file:line in
the comments); the input is the a549 fixture that load_a549() reads
(R/io.R:10-26), committed as tests/golden/a549.npz by oracle/make_golden.py.
"""
import numpy as np
import pytest

from oracle import monocle3_ref as ref
from oracle.r_rng import RRandom


def test_r_rng_known_values():
    # SURVEY.md Appendix A.1 (values printed by R itself)
    np.testing.assert_allclose(RRandom(42).runif(3), [0.9148060, 0.9370754, 0.2861395], atol=5e-8)
    np.testing.assert_allclose(RRandom(42).rnorm(3), [1.37095845, -0.56469817, 0.36312841], atol=5e-9)
    np.testing.assert_allclose(RRandom(2016).rnorm(5),
                               [-0.91474184, 1.00124785, -0.05642291, 0.29664516, -2.79147086], atol=5e-9)


@pytest.fixture(scope="module")
def cds(a549):
    counts, extra = a549
    sf = ref.estimate_size_factors(counts)
    return counts, sf, extra


def test_fixture_shape(cds):
    counts, sf, extra = cds
    assert counts.shape == (198, 500) and counts.nnz == 9824
    # SURVEY.md Appendix C
    np.testing.assert_allclose(sf[:3], [0.60566273, 0.22313890, 2.10388105], rtol=1e-7)


def test_pca_log_nd20(cds):
    counts, sf, _ = cds
    r = ref.preprocess_cds(counts, sf, num_dim=20)
    assert r["PCA"].shape == (500, 20)
    assert r["PCA"][0, 0] == pytest.approx(2.4207391, abs=1e-5)  # test-preprocessing.R:11
    assert r["PCA"][1, 0] == pytest.approx(1.982232, abs=1e-5)  # test-preprocessing.R:75
    # model slots, test-preprocessing.R:47-51
    assert r["svd_v"][0, 0] == pytest.approx(0.1362, abs=1e-3)
    assert r["svd_sdev"][0] == pytest.approx(2.67, abs=1e-2)
    assert r["svd_center"][0] == pytest.approx(2.22475, abs=1e-4)
    assert r["svd_scale"][0] == pytest.approx(0.9393, abs=1e-3)
    assert r["prop_var_expl"][0] == pytest.approx(0.1493, abs=1e-3)
    assert len(r["kept_genes"]) == 179
    assert (r["iter"], r["mprod"]) == (25, 208)


def test_pca_size_only(cds):
    counts, sf, _ = cds
    r = ref.preprocess_cds(counts, sf, num_dim=20, norm_method="size_only")
    assert r["PCA"][0, 0] == pytest.approx(2.222207, abs=1e-5)  # test-preprocessing.R:22


def test_pca_none(cds):
    counts, sf, _ = cds
    r = ref.preprocess_cds(counts, sf, num_dim=20, norm_method="none")
    assert r["PCA"][0, 0] == pytest.approx(-2.42836, abs=1e-5)  # test-preprocessing.R:34


def test_pca_dense_pseudocount_noscale(cds):
    counts, sf, _ = cds
    r = ref.preprocess_cds(counts, sf, num_dim=20, norm_method="size_only", pseudo_count=1.4, scaling=False)
    assert r["PCA"][0, 0] == pytest.approx(-24.30544, abs=1e-5)  # test-preprocessing.R:69


def test_pca_use_genes(cds):
    counts, sf, _ = cds
    r = ref.preprocess_cds(counts, sf, num_dim=20, use_genes_idx=np.arange(100))
    assert r["PCA"][1, 0] == pytest.approx(-0.5347819, abs=1e-5)  # test-preprocessing.R:82


def test_pca_nd50_and_transform(cds):
    counts, sf, _ = cds
    r = ref.preprocess_cds(counts, sf, num_dim=50)
    assert r["PCA"].shape == (500, 50)
    assert r["PCA"][0, 0] == pytest.approx(2.420739, abs=1e-5)  # test-io.R:267
    assert (r["iter"], r["mprod"]) == (35, 326)
    # test-projection.R:28-30: transform of the same cells reproduces the embedding
    gidx = np.full(counts.shape[0], -1)
    gidx[r["kept_genes"]] = np.arange(len(r["kept_genes"]))
    y = ref.preprocess_transform(counts, sf, r, gidx)
    for (a, b) in ((0, 0), (1, 0), (0, 1)):
        assert y[a, b] == pytest.approx(r["PCA"][a, b], abs=1e-6)
    assert np.abs(y - r["PCA"]).max() < 1e-10


def test_normalized_counts(cds):
    counts, sf, _ = cds
    nm = ref.normalized_counts(counts, sf)
    j = 7
    lo, hi = counts.indptr[j], counts.indptr[j + 1]
    np.testing.assert_allclose(nm.data[lo:hi], np.log10(counts.data[lo:hi] / sf[j] + 1), rtol=0, atol=0)
    with pytest.raises(ValueError):
        ref.normalized_counts(counts, sf, pseudocount=2)  # R/utils.R:498
    b = ref.normalized_counts(counts, sf, norm_method="binary")
    assert set(np.unique(b.data)) == {1.0}


def test_zero_read_cell_guard(a549):
    counts, _ = a549
    c = counts.tolil()
    c[:, 3] = 0
    assert ref.estimate_size_factors(c.tocsc()) is None  # R/utils.R:32-39


def test_lsi_goldens(cds):
    """LSI branch: tests/testthat/test-preprocessing.R:13-16,24-28,36-40,54-61."""
    counts, sf, _ = cds
    r = ref.preprocess_cds_lsi(counts, sf, num_dim=20)
    assert r["LSI"].shape == (500, 20)
    assert r["LSI"][0, 0] == pytest.approx(13.73796, abs=1e-5)          # test-preprocessing.R:16
    assert r["col_sums"][0] == pytest.approx(18.7, abs=1e-1)            # :58
    assert r["row_sums"][0] == pytest.approx(467, abs=1)                # :59
    assert r["svd_v"][0, 0] == pytest.approx(0.16318, abs=1e-3)         # :60
    assert r["svd_sdev"][0] == pytest.approx(32.21, abs=1e-2)           # :61
    r2 = ref.preprocess_cds_lsi(counts, sf, num_dim=20, norm_method="size_only")
    assert r2["LSI"][0, 0] == pytest.approx(13.49733, abs=1e-5)         # :28
    r3 = ref.preprocess_cds_lsi(counts, sf, num_dim=20, norm_method="none")
    assert r3["LSI"][0, 0] == pytest.approx(13.49733, abs=1e-5)         # :40
    # test-projection.R:32-34: transform of the same cells reproduces the LSI embedding
    r50 = ref.preprocess_cds_lsi(counts, sf, num_dim=50)
    gidx = np.full(counts.shape[0], -1)
    gidx[r50["kept_genes"]] = np.arange(len(r50["kept_genes"]))
    y = ref.preprocess_transform_lsi(counts, sf, r50, gidx)
    for (a, b) in ((0, 0), (1, 0), (0, 1)):
        assert y[a, b] == pytest.approx(r50["LSI"][a, b], abs=1e-6)


def test_scale_without_center_is_root_mean_square_dense_check():
    """sparse_prcomp_irlba(center=FALSE, scale.=TRUE) scales by sqrt(colSums(x^2) / (nrow - 1)) (R/utils.R:395-397):
    the restatement against a dense SVD of the explicitly scaled matrix."""
    rng = np.random.default_rng(3)
    import scipy.sparse as sp
    X = sp.random(300, 40, density=0.2, format="csc", random_state=rng)
    X.data = np.round(X.data * 5) + 1
    o = ref.sparse_prcomp_irlba(X, 5, center=False, scale=True)
    D = X.toarray()
    rms = np.sqrt((D ** 2).sum(axis=0) / (D.shape[0] - 1))
    np.testing.assert_allclose(o["svd_scale"], rms, rtol=1e-14)
    s = np.linalg.svd(D / rms[None, :], compute_uv=False)[:5]
    np.testing.assert_allclose(o["d"], s, rtol=1e-6)
    assert o["center"] is None
