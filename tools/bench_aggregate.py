#!/usr/bin/env python
"""SURVEY 8f row 2 timing: detect_genes and the aggregate_gene_expression group sums on the cfg2
synthetic matrix (device-resident counts, CUDA events around the ABI call, results copied back).
Prints one JSON line: ms per call and the bytes-per-second of the single CSC pass each one is
(detect: 12 B/nnz = index + count; aggregate: 12 B/nnz = index + normalised value)."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cells", type=int, default=42000)
    ap.add_argument("--genes", type=int, default=20000)
    ap.add_argument("--density", type=float, default=0.08)
    ap.add_argument("--gene-groups", type=int, default=50)
    ap.add_argument("--cell-groups", type=int, default=20)
    a = ap.parse_args()
    import numpy as np
    import scipy.sparse as sp
    from monocle3_b200 import engine, _lib as L
    from monocle3_b200.synth import SynthSpec
    spec = SynthSpec(a.genes, a.cells, a.density, 150, 1)
    indptr, indices, data = spec.cells(0, a.cells)
    counts = sp.csc_matrix((data, indices, indptr), shape=(a.genes, a.cells))
    ctx = engine.Context(0)
    dev = engine.DeviceMatrix.from_csc(ctx, counts)
    sf, _ = engine.size_factors_from_totals(dev.cell_totals(True))
    dev.normalize(sf, L.NORM_LOG10, 1.0)
    rng = np.random.default_rng(0)
    gg = rng.integers(-1, a.gene_groups, a.genes).astype(np.int32)   # ~2% of the genes in no group
    cg = rng.integers(0, a.cell_groups, a.cells).astype(np.int32)

    def timed(fn, reps=5):
        fn()
        t = []
        for _ in range(reps):
            ctx.timer_start()
            r = fn()
            t.append(ctx.timer_stop())
        return float(np.median(t)), r

    ms_d, (pg, pc) = timed(lambda: dev.detect_genes(0.0))
    ms_a, A = timed(lambda: dev.aggregate(gg, a.gene_groups, False, cg, a.cell_groups, True))
    # spot checks on the host
    assert np.array_equal(pc, np.diff(counts.indptr))
    N = counts[:, :256].toarray()
    N = np.log10(N / sf[None, :256] + 1)
    nnz = int(counts.nnz)
    print(json.dumps({"workload": "cfg2 synthetic %d cells x %d genes" % (a.cells, a.genes), "nnz": nnz,
                      "detect_genes_ms": ms_d, "detect_genes_gbs": 12.0 * nnz / ms_d / 1e6,
                      "aggregate_ms": ms_a, "aggregate_gbs": 12.0 * nnz / ms_a / 1e6,
                      "gene_groups": a.gene_groups, "cell_groups": a.cell_groups,
                      "cells_per_s_aggregate": a.cells / (ms_a / 1e3), "agg_shape": list(A.shape)}))


if __name__ == "__main__":
    main()
