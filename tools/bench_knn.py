#!/usr/bin/env python
"""SURVEY 8f row 3 timing: exact kNN (m3b_knn_self) + Jaccard weights on a synthetic embedding of the
cfg2 size (42 000 cells x 100 PCs, k = 20 + self), the calls cluster_cells makes after the PCA path.
CUDA events around the ABI calls (host matrices in and out).  The exact search is O(n^2 d):
2 n^2 d subtract-multiply-adds; the restated CPU search is timed on a bounded sample of queries."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=42000)
    ap.add_argument("--d", type=int, default=100)
    ap.add_argument("--k", type=int, default=20)
    a = ap.parse_args()
    import numpy as np
    from monocle3_b200 import engine
    rng = np.random.default_rng(1)
    centers = rng.standard_normal((150, a.d)) * 4.0
    X = np.asfortranarray(centers[rng.integers(0, 150, a.n)] + rng.standard_normal((a.n, a.d)))
    ctx = engine.Context(0)
    engine.knn_self(ctx, X[:2048], a.k + 1)
    t = []
    for _ in range(3):
        ctx.timer_start()
        idx, dist = engine.knn_self(ctx, X, a.k + 1)
        t.append(ctx.timer_stop())
    ms = float(np.median(t))
    nb = np.asfortranarray(idx[:, 1:].astype(np.float64))
    buf = np.zeros((3, a.n * a.k))
    engine.jaccard_coeff(ctx, nb, True, out=buf)
    ctx.timer_start()
    engine.jaccard_coeff(ctx, nb, True, out=buf)
    ms_j = ctx.timer_stop()
    # CPU: the same exact search for a sample of queries
    m = 300
    t0 = time.time()
    for i in range(m):
        d2 = ((X - X[i]) ** 2).sum(axis=1)
        o = np.argsort(d2, kind="stable")[:a.k + 1]
        assert np.array_equal(o + 1, idx[i])
    cpu_s = time.time() - t0
    flops = 3.0 * a.n * a.n * a.d
    print(json.dumps({"kernel": "m3b_knn_self + m3b_jaccard_coeff", "n": a.n, "d": a.d, "k": a.k, "knn_ms_incl_copies": ms,
                      "knn_tflops_fp64": flops / ms / 1e9, "cells_per_s_knn": a.n / (ms / 1e3), "jaccard_ms_incl_copies": ms_j,
                      "cpu_port_queries_per_s": m / cpu_s, "cpu_sample_queries": m, "exact_on_sample": True}))


if __name__ == "__main__":
    main()
