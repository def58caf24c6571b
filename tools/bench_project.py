#!/usr/bin/env python
"""K10 (preprocess_transform projection) timing: query cells x genes synthetic matrix projected on
nv loadings through m3b_project.  Prints one JSON line (cells/s incl. the D2H of Y, kernel GB/s)."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cells", type=int, default=100000)
    ap.add_argument("--genes", type=int, default=20000)
    ap.add_argument("--density", type=float, default=0.08)
    ap.add_argument("--nv", type=int, default=100)
    a = ap.parse_args()
    import numpy as np
    import scipy.sparse as sp
    from monocle3_b200 import engine, _lib as L
    from monocle3_b200.synth import SynthSpec
    spec = SynthSpec(a.genes, a.cells, a.density, 150, 2)
    indptr, indices, data = spec.cells(0, a.cells)
    counts = sp.csc_matrix((data, indices, indptr), shape=(a.genes, a.cells))
    ctx = engine.Context(0)
    dev = engine.DeviceMatrix.from_csc(ctx, counts)
    sf, _ = engine.size_factors_from_totals(dev.cell_totals(True))
    dev.normalize(sf, L.NORM_LOG2, 1.0)
    keep = dev.select_genes(None)
    n = int(keep.sum())
    rng = np.random.default_rng(0)
    V = np.linalg.qr(rng.standard_normal((n, a.nv)))[0]
    c, s = dev.gene_moments()
    mr = np.arange(n, dtype=np.int32)
    dev.project(mr, V, c, s)
    t = []
    for _ in range(3):
        ctx.timer_start()
        Y = dev.project(mr, V, c, s)
        t.append(ctx.timer_stop())
    ms = float(np.median(t))
    # check a few rows against a host product
    FM = counts.tocsr()[keep, :].T.tocsr()[:64]
    FM.data = np.log2(FM.data / np.repeat(sf[:64], np.diff(FM.indptr)) + 1)
    ref = (FM @ (V / s[:, None])) - (c / s) @ V
    err = float(np.abs(ref - Y[:64]).max())
    nnz = counts.nnz
    print(json.dumps({"kernel": "K10 m3b_project", "cells": a.cells, "genes_kept": n, "nnz": int(nnz), "nv": a.nv,
                      "ms": ms, "cells_per_s": a.cells / (ms / 1e3), "gflops": 2.0 * nnz * a.nv / ms / 1e6,
                      "max_abs_err_first_64_cells": err}))


if __name__ == "__main__":
    main()
