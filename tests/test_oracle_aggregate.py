"""CPU checks of the restated detect_genes / aggregate_gene_expression against
straightforward dense computations.  The reference's tests hold no expected values for these two
functions (parity unpinned), so this only guards the restatement against slips."""
import numpy as np

from oracle import monocle3_ref as ref


def test_detect_genes_dense(a549):
    counts, _ = a549
    D = counts.toarray()
    for thr in (0, 1.5, -0.5):
        pg, pc = ref.detect_genes(counts, thr)
        assert np.array_equal(pg, (D > thr).sum(axis=1)) and np.array_equal(pc, (D > thr).sum(axis=0))


def test_aggregate_dense(a549):
    counts, extra = a549
    genes = [str(g) for g in extra["genes"]]
    cells = [str(c) for c in extra["cells"]]
    sf = ref.estimate_sf_sparse(counts)[0]
    gdf = [(genes[k], "m%d" % (k % 3)) for k in range(60)]
    cdf = [(cells[k], "g%d" % (k % 4)) for k in range(len(cells))]
    A, rn, cn = ref.aggregate_gene_expression(counts, sf, genes, cells, gdf, cdf, scale_agg_values=False)
    N = np.log10(counts.toarray() / sf[None, :] + 1)
    assert rn == ["m0", "m1", "m2"] and cn == ["g0", "g1", "g2", "g3"]
    for i, m in enumerate(rn):
        rows = [k for k in range(60) if k % 3 == i]
        for j, g in enumerate(cn):
            cols = [k for k in range(len(cells)) if k % 4 == j]
            np.testing.assert_allclose(A[i, j], N[np.ix_(rows, cols)].sum(axis=0).mean(), rtol=1e-12)
    Z, _, _ = ref.aggregate_gene_expression(counts, sf, genes, cells, gdf, cdf)
    assert np.abs(Z).max() <= 3 and np.allclose(Z.mean(axis=1), 0, atol=1e-12)
