"""align_cds residual regression on the device against the oracle
(oracle.align_residuals, restated lm.fit; pinned by the reference's own value beta[[1]] = 2.071,
tests/testthat/test-alignment.R:100, in tests/test_oracle_align.py)."""
import numpy as np
import pytest

from oracle import monocle3_ref as ref

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def m3():
    import monocle3_b200 as m
    from monocle3_b200 import engine
    m.set_context(engine.Context(device=0, small_svd="device"))
    return m


def test_align_residuals_vs_oracle(m3, worm_l2):
    counts, extra = worm_l2
    cds = m3.new_cell_data_set(counts)
    cds = m3.preprocess_cds(cds, num_dim=20)
    rng = np.random.default_rng(3)
    n = counts.shape[1]
    cds.col_data["n.umi"] = np.asarray(counts.sum(axis=0)).ravel()
    cds.col_data["batch"] = np.array(["b1", "b2", "b3"])[rng.integers(0, 3, n)]
    cds.col_data["dup"] = 2.0 * cds.col_data["n.umi"]          # collinear with n.umi: aliased -> NA -> 0
    X = cds.reduced_dims["PCA"].copy()
    cds = m3.align_cds(cds, residual_model_formula_str="~ n.umi + batch + dup")
    M, names = ref.model_matrix({k: cds.col_data[k] for k in ("n.umi", "batch", "dup")}, ["n.umi", "batch", "dup"])
    Xo, beta_o = ref.align_residuals(X, M)
    model = cds.reduce_dim_aux["Aligned"]["model"]
    assert model["beta_colnames"] == names[1:]
    assert np.all(model["beta"][:, -1] == 0.0) and np.all(beta_o[:, -1] == 0.0)   # the aliased column
    np.testing.assert_allclose(model["beta"], beta_o, rtol=1e-8, atol=1e-12)
    np.testing.assert_allclose(cds.reduced_dims["Aligned"], Xo, rtol=0, atol=1e-9 * np.abs(X).max())
    # align_beta_transform re-applies the saved beta (R/projection.R:468-478)
    cds2 = m3.align_beta_transform(cds)
    np.testing.assert_allclose(cds2.reduced_dims["Aligned"], Xo, rtol=0, atol=1e-9 * np.abs(X).max())


def test_align_cds_argument_checks(m3, a549):
    counts, extra = a549
    cds = m3.new_cell_data_set(counts)
    with pytest.raises(ValueError, match="does not exist"):
        m3.align_cds(cds, residual_model_formula_str="~ x")
    cds = m3.preprocess_cds(cds, num_dim=5)
    with pytest.raises(ValueError, match="not found"):
        m3.align_cds(cds, residual_model_formula_str="~ nope")
    with pytest.raises(NotImplementedError):
        m3.align_cds(cds, alignment_group="batch")
    cds = m3.align_cds(cds)                    # neither argument: Aligned = PCA (R/alignment.R:131)
    assert np.array_equal(cds.reduced_dims["Aligned"], cds.reduced_dims["PCA"])


def test_transform_model_round_trip(m3, a549, tmp_path):
    counts, extra = a549
    cds = m3.preprocess_cds(m3.new_cell_data_set(counts, cell_names=extra["cells"], gene_names=extra["genes"]), num_dim=10)
    m3.save_transform_models(cds, str(tmp_path))
    q = m3.new_cell_data_set(counts, cell_names=extra["cells"], gene_names=extra["genes"])
    q = m3.load_transform_models(q, str(tmp_path))
    q = m3.preprocess_transform(q, "PCA")
    np.testing.assert_allclose(q.reduced_dims["PCA"], cds.reduced_dims["PCA"], rtol=0, atol=1e-9)
