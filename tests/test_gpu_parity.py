"""GPU parity tests: the CUDA path (through the C ABI / Python host mirror) against the
oracle on the committed fixtures.  Bars (BASELINE.json north_star): size factors and nnz
bookkeeping exact; singular values 1e-6 relative; sign-aligned loadings / embeddings
subspace angle < 1e-4 rad; iter/mprod equal to the oracle's."""
import os

import numpy as np
import pytest
import scipy.sparse as sp

from oracle import monocle3_ref as ref
from tests.parity import column_angles, subspace_angle, sv_rel_err

pytestmark = pytest.mark.gpu

SV_TOL = 1e-6      # relative, north_star
ANGLE_TOL = 1e-4   # rad, north_star


@pytest.fixture(scope="module")
def m3():
    import monocle3_b200 as m
    from monocle3_b200 import engine
    m.set_context(engine.Context(device=0, small_svd="device"))
    return m


def _cds(m3, counts, extra=None):
    genes = extra["genes"] if extra and "genes" in extra else None
    cells = extra["cells"] if extra and "cells" in extra else None
    return m3.new_cell_data_set(counts, cell_names=cells, gene_names=genes)


@pytest.mark.parametrize("name", ["a549", "worm_l2", "worm_embryo"])
def test_size_factors_exact(m3, request, name):
    counts, extra = request.getfixturevalue(name)
    cds = _cds(m3, counts, extra)
    sf_o, tot_o = ref.estimate_sf_sparse(counts)
    dev = cds.device_matrix()
    assert np.array_equal(dev.cell_totals(True), tot_o)  # exact integer sums
    assert np.array_equal(m3.size_factors(cds), sf_o)     # host finish in long double like R
    for method in ("mean-geometric-mean-log-total",):
        cds2 = m3.estimate_size_factors(_cds(m3, counts, extra), method=method)
        np.testing.assert_allclose(m3.size_factors(cds2), ref.estimate_sf_sparse(counts, method=method)[0], rtol=1e-15)


def test_size_factors_rounding_and_zero_cell(m3, a549):
    counts, _ = a549
    c = counts.copy().astype(np.float64)
    c.data = c.data + 0.5  # exercises round-half-even
    cds = _cds(m3, c)
    assert np.array_equal(cds.device_matrix().cell_totals(True), ref.estimate_sf_sparse(c)[1])
    lil = counts.tolil()
    lil[:, 3] = 0
    with pytest.warns(UserWarning, match="zero reads"):
        cds0 = _cds(m3, lil.tocsc())
    assert m3.size_factors(cds0) is None  # cds returned unchanged (R/utils.R:32-39)


@pytest.mark.parametrize("nm", ["log", "size_only", "binary"])
def test_normalized_counts(m3, a549, nm):
    counts, extra = a549
    cds = _cds(m3, counts, extra)
    got = m3.normalized_counts(cds, norm_method=nm)
    exp = ref.normalized_counts(counts, m3.size_factors(cds), norm_method=nm)
    assert np.array_equal(got.indptr, exp.indptr) and np.array_equal(got.indices, exp.indices)
    np.testing.assert_allclose(got.data, exp.data, rtol=4e-16, atol=0)  # <= 2 ulp (CUDA log10 vs libm)
    with pytest.raises(ValueError, match="Pseudocount"):
        m3.normalized_counts(cds, pseudocount=2)


@pytest.mark.parametrize("name", ["a549", "worm_l2"])
def test_filter_nnz_moments(m3, request, name):
    from monocle3_b200 import _lib as L
    counts, extra = request.getfixturevalue(name)
    cds = _cds(m3, counts, extra)
    sf = m3.size_factors(cds)
    dev = cds.device_matrix()
    dev.normalize(sf, L.NORM_LOG2, 1.0)
    y = dev.export_values()
    FM = ref.normalize_expr_data(counts, sf, "log")
    np.testing.assert_allclose(y, FM.data, rtol=4e-16)
    keep = dev.select_genes(None)
    FMf, kept = ref.select_and_filter_genes(FM)
    assert np.array_equal(np.nonzero(keep)[0], kept)
    assert np.array_equal(dev.gene_nnz(), np.diff(counts.tocsr().indptr))  # exact bookkeeping
    c, s = dev.gene_moments()
    co, so = ref.gene_moments(FMf.T.tocsc())
    np.testing.assert_allclose(c, co, rtol=1e-13)
    np.testing.assert_allclose(s, so, rtol=1e-12)


def _check_pca(m3, counts, extra, num_dim, svd_mode="device", **kw):
    """svd_mode: 'device' (structured secular-equation SVD of B, Jacobi fallback), 'jacobi'
    (dense one-sided Jacobi kernel only), 'lapack' (host callback, the R shim's mode)."""
    from monocle3_b200 import engine, _lib as L
    if svd_mode == "jacobi":
        m3.set_context(engine.Context(device=0, flags=L.FLAG_JACOBI_ONLY, small_svd="device"))
    else:
        m3.set_context(engine.Context(device=0, small_svd=svd_mode))
    cds = _cds(m3, counts, extra)
    cds = m3.preprocess_cds(cds, num_dim=num_dim, **kw)
    okw = dict(kw)
    if "use_genes" in okw:
        lut = {g: k for k, g in enumerate(cds.gene_names)}
        okw["use_genes_idx"] = np.array([lut[g] for g in okw.pop("use_genes")])
    o = ref.preprocess_cds(counts, m3.size_factors(cds), num_dim=num_dim, **okw)
    model = cds.reduce_dim_aux["PCA"]["model"]
    ir = cds.reduce_dim_aux["PCA"]["irlba"]
    assert (ir["iter"], ir["mprod"]) == (o["iter"], o["mprod"])
    assert sv_rel_err(ir["d"], o["d"]) < SV_TOL
    assert subspace_angle(model["svd_v"], o["svd_v"]) < ANGLE_TOL
    assert subspace_angle(cds.reduced_dims["PCA"], o["PCA"]) < ANGLE_TOL
    assert column_angles(model["svd_v"], o["svd_v"]).max() < ANGLE_TOL
    assert column_angles(cds.reduced_dims["PCA"], o["PCA"]).max() < ANGLE_TOL
    np.testing.assert_allclose(model["svd_sdev"], o["svd_sdev"], rtol=SV_TOL)
    np.testing.assert_allclose(model["prop_var_expl"], o["prop_var_expl"], rtol=1e-5)
    if o["svd_center"] is not None:
        np.testing.assert_allclose(model["svd_center"], o["svd_center"], rtol=1e-13)
        np.testing.assert_allclose(model["svd_scale"], o["svd_scale"], rtol=1e-12)
    return cds, o


@pytest.mark.parametrize("svd_mode", ["device", "jacobi", "lapack"])
@pytest.mark.parametrize("num_dim", [20, 50])
def test_pca_a549(m3, a549, num_dim, svd_mode):
    counts, extra = a549
    cds, o = _check_pca(m3, counts, extra, num_dim, svd_mode)
    if svd_mode == "lapack":  # sign fidelity: the reference's own goldens, signs included
        assert cds.reduced_dims["PCA"][0, 0] == pytest.approx(2.4207391, abs=1e-5)  # test-preprocessing.R:11
        assert cds.reduced_dims["PCA"][1, 0] == pytest.approx(1.982232, abs=1e-5)  # :75
        # LAPACK's sign choice for the other columns depends on rounding-level details of B,
        # so the full matrix is compared sign-aligned (the north-star definition)
        from tests.parity import sign_align
        np.testing.assert_allclose(sign_align(cds.reduced_dims["PCA"], o["PCA"])[0], o["PCA"], atol=1e-7)


def test_pca_a549_model_slots(m3, a549):
    counts, extra = a549
    cds, _ = _check_pca(m3, counts, extra, 20, "lapack")
    model = cds.reduce_dim_aux["PCA"]["model"]
    # tests/testthat/test-preprocessing.R:45-51
    assert model["num_dim"] == 20 and model["norm_method"] == "log"
    assert model["svd_v"][0, 0] == pytest.approx(0.1362, abs=1e-3)
    assert model["svd_sdev"][0] == pytest.approx(2.67, abs=1e-2)
    assert model["svd_center"][0] == pytest.approx(2.22475, abs=1e-4)
    assert model["svd_scale"][0] == pytest.approx(0.9393, abs=1e-3)
    assert model["prop_var_expl"][0] == pytest.approx(0.1493, abs=1e-3)
    assert cds.reduced_dims["PCA"].shape == (500, 20)
    assert cds.reduced_dims_dimnames["PCA"][0][0] == "E11_A01_RT_467"  # test-reduce_dimension.R:53


@pytest.mark.parametrize("kw,golden,pos", [
    (dict(norm_method="size_only"), 2.222207, (0, 0)),                                   # test-preprocessing.R:22
    (dict(norm_method="none"), -2.42836, (0, 0)),                                        # :34
    (dict(norm_method="size_only", pseudo_count=1.4, scaling=False), -24.30544, (0, 0)),  # :69 (dense path)
])
def test_pca_a549_variants(m3, a549, kw, golden, pos):
    counts, extra = a549
    cds, _ = _check_pca(m3, counts, extra, 20, "lapack", **kw)
    assert cds.reduced_dims["PCA"][pos] == pytest.approx(golden, abs=1e-5)


def test_pca_a549_use_genes(m3, a549):
    counts, extra = a549
    cds, _ = _check_pca(m3, counts, extra, 20, "lapack", use_genes=list(extra["genes"][:100]))
    assert cds.reduced_dims["PCA"][1, 0] == pytest.approx(-0.5347819, abs=1e-5)  # test-preprocessing.R:82


def test_pca_log_pseudocount_dense_paths(m3, a549):
    counts, extra = a549
    # with a pseudo-count the all-zero genes survive the row-sum filter (constant rows) and
    # have sd 0, which breaks the reference itself; restrict to expressed genes when scaling
    nz = list(extra["genes"][np.asarray(counts.sum(axis=1)).ravel() > 0])
    _check_pca(m3, counts, extra, 10, "device", norm_method="log", pseudo_count=2.0, use_genes=nz)
    _check_pca(m3, counts, extra, 10, "device", norm_method="log", pseudo_count=0.5, scaling=False)
    _check_pca(m3, counts, extra, 10, "device", norm_method="log", scaling=False)


@pytest.mark.parametrize("svd_mode", ["device", "jacobi"])
@pytest.mark.parametrize("name,num_dim", [("worm_l2", 50), ("worm_embryo", 50), ("worm_l2", 100)])
def test_pca_worm(m3, request, name, num_dim, svd_mode):
    counts, extra = request.getfixturevalue(name)
    cds, _ = _check_pca(m3, counts, extra, num_dim, svd_mode)
    st = m3.get_context().stats()
    if svd_mode == "device":  # the structured path must actually carry the restarts
        assert st["svd_fallbacks"] <= max(2, cds.reduce_dim_aux["PCA"]["irlba"]["iter"] // 20)


def test_transform_equals_embedding(m3, a549):
    """tests/testthat/test-projection.R:14-30: preprocess_transform of the same cells
    reproduces preprocess_cds's embedding."""
    counts, extra = a549
    cds, o = _check_pca(m3, counts, extra, 50, "lapack")
    cds2 = _cds(m3, counts, extra)
    cds2.reduce_dim_aux = {"PCA": cds.reduce_dim_aux["PCA"]}  # load_transform_models
    cds2 = m3.preprocess_transform(cds2)
    a, b = cds.reduced_dims["PCA"], cds2.reduced_dims["PCA"]
    for (i, j) in ((0, 0), (1, 0), (0, 1)):
        assert a[i, j] == pytest.approx(b[i, j], abs=1e-6)
    assert np.abs(a - b).max() < 1e-9
    assert b[0, 0] == pytest.approx(2.420739, abs=1e-5)  # test-io.R:267


def test_transform_query_subset_and_gene_mismatch(m3, worm_l2):
    counts, extra = worm_l2
    ref_counts = counts[:, :3000]
    q_counts = counts[:, 3000:]
    cds = _cds(m3, ref_counts)
    cds = m3.preprocess_cds(cds, num_dim=30)
    model = cds.reduce_dim_aux["PCA"]["model"]
    q = _cds(m3, q_counts)
    q.reduce_dim_aux = {"PCA": cds.reduce_dim_aux["PCA"]}
    q = m3.preprocess_transform(q)
    # oracle projection with the SAME model (so only the projection arithmetic is compared)
    lut = {g: k for k, g in enumerate(model["svd_v_rownames"])}
    gidx = np.array([lut.get(g, -1) for g in q.gene_names])
    yo = ref.preprocess_transform(q_counts, m3.size_factors(q), model, gidx)
    np.testing.assert_allclose(q.reduced_dims["PCA"], yo, atol=1e-9, rtol=1e-9)


def test_errors(m3, a549):
    from monocle3_b200 import _lib as L
    counts, extra = a549
    cds = _cds(m3, counts, extra)
    with pytest.raises(ValueError, match="norm_method"):
        m3.preprocess_cds(cds, norm_method="bogus")
    with pytest.raises(ValueError, match="use_genes"):
        m3.preprocess_cds(cds, use_genes=["not-a-gene"])
    with pytest.raises(L.M3BError) as ei:  # nv >= min(dim): irlba's own error
        cds.device_matrix().normalize(m3.size_factors(cds), L.NORM_LOG2, 1.0)
        cds.device_matrix().select_genes(None)
        cds.device_matrix().pca_irlba(400, np.ones(179))
    assert ei.value.status == L.M3B_E_DIMS
    with pytest.raises(L.M3BError) as ei:  # call-order violation
        from monocle3_b200 import engine
        d = engine.DeviceMatrix.from_csc(m3.get_context(), counts)
        d.select_genes(None)
    assert ei.value.status == L.M3B_E_STATE


def test_ragged_and_empty_inputs(m3):
    """Empty cells are rejected by the zero-read guard; empty genes are filtered; chunked
    append with ragged chunks gives the same result as one chunk."""
    from monocle3_b200 import engine, _lib as L
    rng = np.random.default_rng(0)
    G, C = 300, 700
    dense = rng.poisson(0.3, size=(G, C)).astype(np.float64)
    dense[5, :] = 0
    dense[17, :] = 0
    dense[:, dense.sum(axis=0) == 0] = 1
    counts = sp.csc_matrix(dense)
    ctx = m3.get_context()
    one = engine.DeviceMatrix.from_csc(ctx, counts)
    rag = engine.DeviceMatrix(ctx, G, C, C, 0)  # nnz_hint 0 forces the grow path
    cuts = [0, 1, 2, 130, 131, 699, 700]
    for a, b in zip(cuts[:-1], cuts[1:]):
        p = counts.indptr[a:b + 1] - counts.indptr[a]
        rag.append_csc(p, counts.indices[counts.indptr[a]:counts.indptr[b]], counts.data[counts.indptr[a]:counts.indptr[b]])
    rag.end()
    assert np.array_equal(one.cell_totals(), rag.cell_totals())
    sf = ref.estimate_sf_sparse(counts)[0]
    for d in (one, rag):
        d.normalize(sf, L.NORM_LOG2, 1.0)
    k1, k2 = one.select_genes(None), rag.select_genes(None)
    assert np.array_equal(k1, k2) and not k1[5] and not k1[17] and k1.sum() == G - 2
    v0 = np.linspace(-1, 1, int(k1.sum()))
    r1 = one.pca_irlba(10, v0)
    r2 = rag.pca_irlba(10, v0)
    assert np.array_equal(r1["d"], r2["d"]) and np.array_equal(r1["x"], r2["x"])  # bit-reproducible
    o = ref.preprocess_cds(counts, sf, num_dim=10, v0=v0)
    assert sv_rel_err(r1["d"], o["d"]) < SV_TOL and subspace_angle(r1["x"], o["PCA"]) < ANGLE_TOL


@pytest.mark.parametrize("scaling", [True, False])
def test_pca_synthetic_vs_oracle(m3, scaling):
    """Synthetic matrix from the bench generator (planted cell types, Poisson counts), small enough
    for the oracle: full parity incl. iter/mprod."""
    from monocle3_b200.synth import synth_csc
    counts = synth_csc(2500, 3000, 0.08, 60, seed=5)
    _check_pca(m3, counts, None, 50, "device", scaling=scaling)


def test_pca_multi_tile_layouts(m3):
    """More genes than one shared-memory tile holds (G' > 28672) and more cells than one tile
    (C > 28672): both layouts are tiled, partials of a row are combined across tiles."""
    from monocle3_b200 import engine, _lib as L
    from monocle3_b200.synth import synth_csc
    counts = synth_csc(30000, 29500, 0.004, 30, seed=7)  # ~3.5M nnz
    sf = ref.estimate_sf_sparse(counts)[0]
    ctx = m3.get_context()
    dev = engine.DeviceMatrix.from_csc(ctx, counts)
    dev.normalize(sf, L.NORM_LOG2, 1.0)
    keep = dev.select_genes(None)
    v0 = np.cos(np.arange(int(keep.sum())) * 0.37) + 0.1
    r = dev.pca_irlba(12, v0, maxit=1000)
    FM = ref.normalize_expr_data(counts, sf, "log")
    FMf, kept = ref.select_and_filter_genes(FM)
    assert np.array_equal(np.nonzero(keep)[0], kept) and len(kept) > 28672
    o = ref.sparse_prcomp_irlba(FMf.T.tocsc(), 12, center=True, scale=True, v0=v0)
    assert (r["iter"], r["mprod"]) == (o["iter"], o["mprod"])
    assert sv_rel_err(r["d"], o["d"]) < SV_TOL
    assert subspace_angle(r["v"], o["rotation"]) < ANGLE_TOL
    assert subspace_angle(r["x"], o["x"]) < ANGLE_TOL


def _pca_direct(ctx, counts, nd, v0=None, method=None, pc=1.0):
    from monocle3_b200 import engine, _lib as L
    sf = ref.estimate_sf_sparse(counts)[0]
    dev = engine.DeviceMatrix.from_csc(ctx, counts)
    dev.normalize(sf, L.NORM_LOG2 if method is None else method, pc)
    keep = dev.select_genes(None)
    if v0 is None:
        v0 = np.cos(np.arange(int(keep.sum())) * 0.37) + 0.1
    r = dev.pca_irlba(nd, v0, maxit=1000)
    r["keep"] = keep
    r["moments"] = dev.gene_moments()
    return r, sf, v0


def test_unit_split_matches_unsplit(m3, worm_l2):
    """Count-1 entries in the index-only unit layouts vs everything in the general stream: same
    moments, same iteration, same singular triplets (the two differ in summation order only)."""
    from monocle3_b200 import engine, _lib as L
    counts, _ = worm_l2
    r1, _, v0 = _pca_direct(engine.Context(device=0, small_svd="device"), counts, 30)
    r0, _, _ = _pca_direct(engine.Context(device=0, flags=L.FLAG_NO_UNIT_SPLIT, small_svd="device"), counts, 30, v0)
    assert np.array_equal(r1["keep"], r0["keep"])
    np.testing.assert_allclose(r1["moments"][0], r0["moments"][0], rtol=1e-13)
    np.testing.assert_allclose(r1["moments"][1], r0["moments"][1], rtol=1e-12)
    assert (r1["iter"], r1["mprod"]) == (r0["iter"], r0["mprod"])
    assert sv_rel_err(r1["d"], r0["d"]) < 1e-11
    assert subspace_angle(r1["v"], r0["v"]) < 1e-8 and subspace_angle(r1["x"], r0["x"]) < 1e-8


@pytest.mark.parametrize("case", ["all_ones", "no_ones", "tiny_rows", "size_only"])
def test_unit_split_edge_cases(m3, case):
    """Edge cases of the unit / general split against the oracle: a binary matrix (general stream
    empty), no count-1 entry at all (unit stream empty), thousands of 1-3 entry rows (dozens of
    item starts per 32-lane round, slot windows re-based mid-round), and the size_only values."""
    from monocle3_b200 import _lib as L
    from monocle3_b200.synth import synth_csc
    method, nm, pc = None, "log", 1.0
    if case == "tiny_rows":
        counts = synth_csc(9000, 2200, 0.0015, 20, seed=3)
    else:
        counts = synth_csc(1800, 2100, 0.05, 20, seed=4)
    if case == "all_ones":
        counts.data[:] = 1.0
    elif case == "no_ones":
        counts.data[:] = counts.data + 1.0
    elif case == "size_only":
        method, nm, pc = L.NORM_SIZE_ONLY, "size_only", 0.0
    r, sf, v0 = _pca_direct(m3.get_context(), counts, 15, method=method, pc=pc)
    FM = ref.normalize_expr_data(counts, sf, nm, pseudo_count=pc)
    FMf, kept = ref.select_and_filter_genes(FM)
    assert np.array_equal(np.nonzero(r["keep"])[0], kept)
    o = ref.sparse_prcomp_irlba(FMf.T.tocsc(), 15, center=True, scale=True, v0=v0)
    assert (r["iter"], r["mprod"]) == (o["iter"], o["mprod"])
    assert sv_rel_err(r["d"], o["d"]) < SV_TOL
    assert subspace_angle(r["v"], o["rotation"]) < ANGLE_TOL
    assert subspace_angle(r["x"], o["x"]) < ANGLE_TOL


def test_full_size_properties(m3):
    """Size-independent properties at a size the oracle cannot iterate on in seconds
    (20k cells x 8k genes, ~10M nnz): orthonormal loadings, singular-triplet residuals
    ||A^ v_i - x_i|| <= 1e-8 sigma_1 checked with one host product per vector, embeddings
    orthogonal with norms sigma_i, run-to-run bit reproducibility."""
    from monocle3_b200 import engine, _lib as L
    from monocle3_b200.synth import synth_csc
    from monocle3_b200.r_rng import rnorm_after_seed
    counts = synth_csc(8000, 20000, 0.06, 100, seed=11)
    sf = ref.estimate_sf_sparse(counts)[0]
    ctx = m3.get_context()
    dev = engine.DeviceMatrix.from_csc(ctx, counts)
    assert np.array_equal(dev.cell_totals(True), np.asarray(counts.sum(axis=0)).ravel())
    dev.normalize(sf, L.NORM_LOG2, 1.0)
    keep = dev.select_genes(None)
    v0 = rnorm_after_seed(2016, int(keep.sum()))
    r = dev.pca_irlba(30, v0)
    r2 = dev.pca_irlba(30, v0)
    assert r["status"] == 0 and np.array_equal(r["x"], r2["x"]) and np.array_equal(r["v"], r2["v"])
    V, X, d = r["v"], r["x"], r["d"]
    assert np.abs(V.T @ V - np.eye(30)).max() < 1e-10
    G = X.T @ X
    assert np.abs(G - np.diag(d ** 2)).max() < 1e-8 * d[0] ** 2
    assert np.all(np.diff(d) <= 0)
    FM = ref.normalize_expr_data(counts, sf, "log")
    A = FM.tocsr()[keep, :].T.tocsr()                   # cells x kept genes
    c, s = r["center"], r["scale"]
    for i in (0, 7, 29):
        xs = V[:, i] / s
        w = A @ xs - np.dot(xs, c)
        assert np.linalg.norm(w - X[:, i]) < 1e-8 * d[0]
    # irlba's own stopping rule: residual of the accepted triplets below tol * sigma_1 (tol = 1e-5)
    for i in (0, 29):
        u = X[:, i] / d[i]
        f = (A.T @ u) / s - np.sum(u) * c / s
        assert np.linalg.norm(f - d[i] * V[:, i]) < 1e-4 * d[0]


def test_lsi_branch(m3, a549):
    """LSI branch (tfidf + irlba without centre/scale) against the oracle and the reference's
    goldens, tests/testthat/test-preprocessing.R:13-16,54-61, and the LSI transform equality of
    tests/testthat/test-projection.R:32-34."""
    from monocle3_b200 import engine
    counts, extra = a549
    m3.set_context(engine.Context(device=0, small_svd="lapack"))
    cds = _cds(m3, counts, extra)
    cds = m3.preprocess_cds(cds, method="LSI", num_dim=20)
    o = ref.preprocess_cds_lsi(counts, m3.size_factors(cds), num_dim=20)
    model, ir = cds.reduce_dim_aux["LSI"]["model"], cds.reduce_dim_aux["LSI"]["irlba"]
    assert (ir["iter"], ir["mprod"]) == (o["iter"], o["mprod"])
    assert sv_rel_err(ir["d"], o["d"]) < SV_TOL
    assert subspace_angle(model["svd_v"], o["svd_v"]) < ANGLE_TOL
    assert subspace_angle(cds.reduced_dims["LSI"], o["LSI"]) < ANGLE_TOL
    np.testing.assert_allclose(model["col_sums"], o["col_sums"], rtol=1e-13)
    assert np.array_equal(model["row_sums"], o["row_sums"])               # exact counts
    assert cds.reduced_dims["LSI"][0, 0] == pytest.approx(13.73796, abs=1e-5)   # test-preprocessing.R:16
    assert model["col_sums"][0] == pytest.approx(18.7, abs=1e-1)                 # :58
    assert model["row_sums"][0] == pytest.approx(467, abs=1)                     # :59
    assert model["svd_v"][0, 0] == pytest.approx(0.16318, abs=1e-3)              # :60
    assert model["svd_sdev"][0] == pytest.approx(32.21, abs=1e-2)                # :61
    for nm, gold in (("size_only", 13.49733), ("none", 13.49733)):                # :28, :40
        c2 = m3.preprocess_cds(_cds(m3, counts, extra), method="LSI", num_dim=20, norm_method=nm)
        assert c2.reduced_dims["LSI"][0, 0] == pytest.approx(gold, abs=1e-5)
    # transform of the same cells == embedding
    cds50 = m3.preprocess_cds(_cds(m3, counts, extra), method="LSI", num_dim=50)
    q = _cds(m3, counts, extra)
    q.reduce_dim_aux = {"LSI": cds50.reduce_dim_aux["LSI"]}
    q = m3.preprocess_transform(q, reduction_method="LSI")
    assert np.abs(q.reduced_dims["LSI"] - cds50.reduced_dims["LSI"]).max() < 1e-8


@pytest.mark.parametrize("min_expr", [0, 0.1, 2, -1])
def test_detect_genes(m3, a549, min_expr):
    """SURVEY 8f row 2 / R/utils.R:449-457: thresholded expression counts, exact (integers)."""
    counts, extra = a549
    cds = m3.detect_genes(_cds(m3, counts, extra), min_expr=min_expr)
    pg, pc = ref.detect_genes(counts, min_expr)
    assert np.array_equal(cds.row_data["num_cells_expressed"], pg)
    assert np.array_equal(cds.col_data["num_genes_expressed"], pc)
    if min_expr == 0:  # dense cross-check of the restatement itself
        D = counts.toarray()
        assert np.array_equal(pg, (D > 0).sum(axis=1)) and np.array_equal(pc, (D > 0).sum(axis=0))


@pytest.mark.parametrize("kw", [dict(), dict(norm_method="size_only", gene_agg_fun="mean", cell_agg_fun="sum"),
                                dict(norm_method="binary", scale_agg_values=False), dict(cells=False),
                                dict(genes=False, scale_agg_values=False)])
def test_aggregate_gene_expression(m3, worm_l2, kw):
    """SURVEY 8f row 2 / R/cluster_genes.R:447-536 against the restatement: gene groups (one with a
    single gene -> dropped, one gene in two groups, an unknown id), cell groups incl. 'NA'."""
    counts, extra = worm_l2
    kw = dict(kw)
    cds = _cds(m3, counts, extra)
    genes, cells = cds.gene_names.tolist(), cds.cell_names.tolist()
    rng = np.random.default_rng(3)
    gsel = rng.choice(len(genes), 400, replace=False)
    gdf = [(genes[g], "mod%d" % (k % 7)) for k, g in enumerate(gsel)]
    gdf += [(genes[gsel[0]], "mod3"), ("not_a_gene", "mod1"), (genes[gsel[1]], "solo")]
    labels = ["NA" if k % 11 == 0 else "part%d" % (k % 5) for k in range(len(cells))]
    cdf = [(c, l) for c, l in zip(cells, labels)][::-1]  # the order of the table must not matter
    use_g = gdf if kw.pop("genes", True) else None
    use_c = cdf if kw.pop("cells", True) else None
    A, rn, cn = m3.aggregate_gene_expression(cds, use_g, use_c, **kw)
    Ao, rno, cno = ref.aggregate_gene_expression(counts, m3.size_factors(cds), genes, cells, use_g, use_c, **kw)
    assert rn == rno and cn == cno and A.shape == Ao.shape
    if use_g is not None:
        assert "solo" not in rn
    np.testing.assert_allclose(A, Ao, rtol=1e-10, atol=1e-12)


def test_invariant_subspace_random_restart(m3):
    """irlb.c replaces a Lanczos vector whose norm falls below eps (2.2e-16, absolute) by rnorm()
    from R's stream.  A rank-4 matrix of norm ~1e-8 gets there after five steps; with the rnorm
    callback the library follows irlba step for step (same stream, same iter / mprod, same
    singular values), without it the call reports the linear dependency."""
    from monocle3_b200 import engine, _lib as L
    from monocle3_b200.r_rng import RStream
    from oracle import irlba_ref
    from oracle.r_rng import RRandom
    rng = np.random.default_rng(5)
    G, Cn, r = 40, 70, 4
    dense = (rng.integers(1, 4, (G, r)) @ rng.integers(0, 3, (r, Cn))).astype(np.float64) * 1e-9
    counts = sp.csc_matrix(dense)
    ctx = engine.Context(device=0, small_svd="device")
    dev = engine.DeviceMatrix.from_csc(ctx, counts)
    dev.normalize(None, L.NORM_NONE, 0.0)
    keep = dev.select_genes(None)
    n = int(keep.sum())
    stream = RStream(7)
    v0 = stream.rnorm(n)
    ctx.set_rnorm(stream.rnorm)
    res = dev.pca_irlba(6, v0, center=False, scale=False)
    ctx.set_rnorm(None)
    orng = RRandom(7)
    v0o = orng.rnorm(n)
    assert np.array_equal(v0, v0o)
    At = sp.csc_matrix(dense[keep, :].T)
    o = irlba_ref.irlba(At, 6, v0o, rng=orng)
    assert res["status"] == 0 and o.err == 0
    assert (res["iter"], res["mprod"]) == (o.iter, o.mprod)
    np.testing.assert_allclose(res["d"][:r], o.d[:r], rtol=1e-6)
    assert np.all(res["d"][r:] < 1e-6 * res["d"][0])
    assert subspace_angle(res["v"][:, :r], o.v[:, :r]) < ANGLE_TOL
    # without a random stream the same input is reported as a linear dependency (irlba's -5)
    with pytest.raises(Exception, match="linear dependency|invariant"):
        dev.pca_irlba(6, v0, center=False, scale=False)


@pytest.mark.parametrize("n,k", [(500, 20), (300, 33), (64, 64), (7, 1), (300, 100), (40, 1500)])
def test_jaccard_coeff_bit_exact(m3, n, k):
    """The reference's native kernel (src/clustering.cpp:13-58): bit-identical to the restatement,
    with repeated ids inside a row and asymmetric neighbourhoods."""
    rng = np.random.default_rng(n + k)
    idx = rng.integers(1, n + 1, (n, k))
    idx[::7, : max(1, k // 3)] = idx[::7, :1]       # duplicates inside some rows
    for weight in (True, False):
        got = m3.jaccard_coeff(idx.astype(np.float64), weight)
        assert np.array_equal(got, ref.jaccard_coeff(idx, weight))
    with pytest.raises(Exception, match="1..n"):
        m3.jaccard_coeff(np.array([[0.0, 2.0], [1.0, 2.0]]), True)


@pytest.mark.parametrize("n,k", [(500, 20), (64, 64), (2000, 15), (200, 130)])
def test_jaccard_coeff_vs_reference_native(m3, n, k):
    """m3b_jaccard_coeff against the REFERENCE'S OWN code: oracle/_ref/libm3ref.so is
    src/clustering.cpp compiled unmodified against a shim of the Rcpp containers (oracle/Makefile).
    Bit-identical (ids, row order, weights)."""
    from tests.test_oracle_ref_native import LIB, native_jaccard
    if not os.path.exists(LIB):
        pytest.skip("oracle/_ref/libm3ref.so not built")
    rng = np.random.default_rng(3 * n + k)
    idx = rng.integers(1, n + 1, (n, k))
    idx[::7, : max(1, k // 3)] = idx[::7, :1]
    for weight in (True, False):
        got = m3.jaccard_coeff(idx.astype(np.float64), weight)
        assert np.array_equal(got, native_jaccard(idx, weight))


@pytest.mark.parametrize("n,d,k", [(1000, 50, 21), (130, 7, 5), (64, 100, 64), (257, 3, 1), (400, 20, 150)])
def test_knn_self_exact(m3, n, d, k):
    """Exact kNN (what RANN::nn2 returns for cluster_cells, R/cluster_cells.R:286-287) against the
    restatement: identical neighbour ids, distances to 1e-14 relative."""
    from monocle3_b200 import engine
    rng = np.random.default_rng(n * 7 + d)
    X = rng.standard_normal((n, d)) * rng.uniform(0.5, 3.0, d)
    idx, dist = engine.knn_self(m3.get_context(), X, k)
    io, do = ref.nn2_self(X, k)
    assert np.array_equal(idx, io)
    np.testing.assert_allclose(dist, do, rtol=1e-13, atol=1e-300)
    assert np.array_equal(idx[:, 0], np.arange(1, n + 1)) and np.all(dist[:, 0] == 0.0)


def test_knn_jaccard_graph_on_pca(m3, worm_l2):
    """The step after the path: kNN graph + Jaccard weights on the PCA embedding
    (cluster_cells_make_graph, R/cluster_cells.R:255-336, nn2 method), device vs restatement."""
    counts, extra = worm_l2
    cds = m3.preprocess_cds(_cds(m3, counts[:, :1500], None), num_dim=20)
    X = cds.reduced_dims["PCA"]
    g = m3.cluster_cells_make_graph(X, True, k=20, nn_control={"method": "nn2"})
    links, distm = ref.cluster_cells_make_graph(X, True, 20)
    assert np.array_equal(g["relations"][0], links[:, 0]) and np.array_equal(g["relations"][1], links[:, 1])
    assert np.array_equal(g["relations"][2], links[:, 2])
    np.testing.assert_allclose(g["distMatrix"], distm, rtol=1e-13)


@pytest.mark.parametrize("shape", [(5, 300), (300, 5), (4, 4000)])
def test_irlba_tiny_problem_dense_branch(m3, shape):
    """min(cells, kept genes) < 6: irlba switches to a dense svd (irlba R/irlba.R);
    iter = mprod = 0."""
    G, C = shape
    rng = np.random.default_rng(7)
    dense = rng.poisson(1.5, size=(G, C)).astype(np.float64)
    dense[:, dense.sum(axis=0) == 0] = 1.0
    dense[dense.sum(axis=1) == 0, 0] = 1.0
    counts = sp.csc_matrix(dense)
    nd = min(G, C) - 2
    for scaling in (True, False):
        cds = m3.preprocess_cds(_cds(m3, counts), num_dim=nd, scaling=scaling)
        o = ref.preprocess_cds(counts, ref.estimate_size_factors(counts), num_dim=nd, scaling=scaling)
        ir = cds.reduce_dim_aux["PCA"]["irlba"]
        assert ir["iter"] == 0 and ir["mprod"] == 0 and o["iter"] == 0
        assert sv_rel_err(ir["d"], o["d"]) < SV_TOL
        assert subspace_angle(cds.reduce_dim_aux["PCA"]["model"]["svd_v"], o["svd_v"]) < ANGLE_TOL
        assert subspace_angle(cds.reduced_dims["PCA"], o["PCA"]) < ANGLE_TOL


def test_abi_scale_without_center_is_root_mean_square(m3, worm_l2):
    """m3b_pca_irlba(center=0, scale=1): sparse_prcomp_irlba scales by sqrt(colSums(x^2) / (n - 1)) when
    it does not centre (R/utils.R:395-397), not by the standard deviation.  preprocess_cds never asks
    for this combination (center == scale == scaling), the C ABI can."""
    from monocle3_b200 import engine, _lib as L
    from monocle3_b200.r_rng import rnorm_after_seed
    counts, _ = worm_l2
    sf = ref.estimate_size_factors(counts)
    FM = ref.normalize_expr_data(counts, sf, "log")
    FM, kept = ref.select_and_filter_genes(FM)
    At = FM.T.tocsc()
    v0 = rnorm_after_seed(2016, At.shape[1])
    o = ref.sparse_prcomp_irlba(At, 20, center=False, scale=True, v0=v0)
    ctx = m3.get_context()
    dev = engine.DeviceMatrix.from_csc(ctx, counts)
    dev.normalize(sf, L.NORM_LOG2, 1.0)
    dev.select_genes(None)
    r = dev.pca_irlba(20, v0, center=False, scale=True)
    np.testing.assert_allclose(r["scale"], o["svd_scale"], rtol=1e-12)
    assert (r["iter"], r["mprod"]) == (o["iter"], o["mprod"])
    assert sv_rel_err(r["d"], o["d"]) < SV_TOL
    assert subspace_angle(r["v"], o["rotation"]) < ANGLE_TOL and subspace_angle(r["x"], o["x"]) < ANGLE_TOL


def test_csc_blocks_equal_one_matrix(m3, worm_l2):
    """A matrix handed over as consecutive cell blocks (monocle3_b200.CSCBlocks: how a matrix beyond
    the dgCMatrix limit of 2^31 entries reaches the library) gives bit for bit the
    result of the same matrix as one CSC, through the public API."""
    counts, _ = worm_l2
    C = counts.shape[1]
    cuts = [0, 1000, 1001, 3500, C]
    blocks = m3.CSCBlocks([counts[:, a:b] for a, b in zip(cuts[:-1], cuts[1:])])
    assert blocks.shape == counts.shape and blocks.nnz == counts.nnz
    a = m3.preprocess_cds(m3.new_cell_data_set(counts), num_dim=20)
    b = m3.preprocess_cds(m3.new_cell_data_set(blocks), num_dim=20)
    assert np.array_equal(m3.size_factors(a), m3.size_factors(b))
    assert np.array_equal(a.reduced_dims["PCA"], b.reduced_dims["PCA"])
    assert np.array_equal(a.reduce_dim_aux["PCA"]["model"]["svd_v"], b.reduce_dim_aux["PCA"]["model"]["svd_v"])
    with pytest.raises(ValueError, match="per block"):
        m3.normalized_counts(b)
