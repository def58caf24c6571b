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


def test_jaccard_hand_worked():
    """src/clustering.cpp:13-46 on cases small enough to do by hand."""
    # three nodes, every pair of neighbour sets shares exactly one id: u = 1, v = 2*2 - 1 = 3
    w = ref.jaccard_coeff(np.array([[2, 3], [1, 3], [1, 2]]), True)
    assert np.array_equal(w[:, 0], [1, 1, 2, 2, 3, 3]) and np.array_equal(w[:, 1], [2, 3, 1, 3, 1, 2])
    assert np.array_equal(w[:, 2], np.ones(6))          # all 1/3, divided by the maximum 1/3
    # node 1 and 2 share both neighbours (u = 2 -> 2/2 = 1); node 3's list has a duplicate and shares
    # only id 4 with node 1 (unique values count once: u = 1 -> 1/3); no overlap keeps weight 1
    idx = np.array([[3, 4], [3, 4], [4, 4], [1, 2]])
    w = ref.jaccard_coeff(idx, True)
    exp = {(1, 3): 1 / 3, (1, 4): 1.0, (2, 3): 1 / 3, (2, 4): 1.0, (3, 4): 1.0, (4, 1): 1.0, (4, 2): 1.0}
    for r in range(8):
        assert w[r, 2] == exp[(int(w[r, 0]), int(w[r, 1]))]
    assert np.array_equal(ref.jaccard_coeff(idx, False)[:, 2], np.ones(8))
