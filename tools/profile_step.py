#!/usr/bin/env python
"""Short driver for ncu captures: builds the bench workload (cfg2 by default) and runs the
hot path with a small restart budget, so that a `ncu -k regex:<kernel> -s N -c M` pass over
it finishes in seconds.  Not a benchmark: numbers printed under ncu are never bench values.

  python tools/profile_step.py [--cells 42000 --genes 20000 --density 0.08 --num-dim 100 --maxit 3]
"""
import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cells", type=int, default=42000)
    ap.add_argument("--genes", type=int, default=20000)
    ap.add_argument("--density", type=float, default=0.08)
    ap.add_argument("--num-dim", type=int, default=100)
    ap.add_argument("--maxit", type=int, default=3)
    a = ap.parse_args()
    import numpy as np
    import scipy.sparse as sp
    from monocle3_b200 import engine, _lib as L
    from monocle3_b200.synth import SynthSpec
    from monocle3_b200.r_rng import rnorm_after_seed
    spec = SynthSpec(a.genes, a.cells, a.density, 150, 1)
    indptr, indices, data = spec.cells(0, a.cells)
    counts = sp.csc_matrix((data, indices, indptr), shape=(a.genes, a.cells))
    ctx = engine.Context(0, small_svd="device")
    dev = engine.DeviceMatrix.from_csc(ctx, counts)
    tot = dev.cell_totals(True)
    sf, _ = engine.size_factors_from_totals(tot)
    t0 = time.time()
    dev.normalize(sf, L.NORM_LOG2, 1.0)
    dev.select_genes(None)
    v0 = rnorm_after_seed(2016, dev.n_kept)
    r = dev.pca_irlba(a.num_dim, v0, maxit=a.maxit, want_x=False)
    print("profile_step: nnz=%d kept=%d iter=%d mprod=%d status=%d %.2fs" % (
        counts.nnz, dev.n_kept, r["iter"], r["mprod"], r["status"], time.time() - t0))


if __name__ == "__main__":
    main()
