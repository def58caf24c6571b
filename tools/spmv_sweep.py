#!/usr/bin/env python
"""A/B of the product-kernel variants (flat.cu) on one synthetic matrix, in ONE process:

  python tools/spmv_sweep.py --cells 250000 --genes 30000 --density 0.06 --num-dim 50 \
      --tile-w 20480,16384 --variants 1,3,4 [--maxit 6]

For every tile width (a new context: M3B_TILE_W is read by m3b_create) the matrix is ingested and
normalised once; for every variant (M3B_FLAT_VARIANT, re-read per launch under M3B_FLAT_DYNAMIC)
the layouts are rebuilt and `maxit` restart cycles of IRLBA run with every product bracketed by
CUDA events (M3B_FLAG_TIME_SPMV).  Prints one JSON line per combination: us per launch of the two
products, physical stream GB/s, and the leading Ritz values after `maxit` cycles, which must agree
between variants to rounding (a variant that streams wrong data does not get past this check).
Not a benchmark line: the numbers feed DESIGN.md's kernel notes only."""
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
    ap.add_argument("--num-dim", type=int, default=100)
    ap.add_argument("--maxit", type=int, default=6)
    ap.add_argument("--tile-w", default="20480")
    ap.add_argument("--variants", default="1,3")
    ap.add_argument("--extra-env", default="", help="comma-separated NAME=VALUE pairs applied to every run")
    a = ap.parse_args()
    os.environ["M3B_FLAT_DYNAMIC"] = "1"
    for kv in [x for x in a.extra_env.split(",") if x]:
        k, v = kv.split("=", 1)
        os.environ[k] = v
    import numpy as np
    import scipy.sparse as sp
    from monocle3_b200 import engine, _lib as L
    from monocle3_b200.synth import SynthSpec
    from monocle3_b200.r_rng import rnorm_after_seed
    spec = SynthSpec(a.genes, a.cells, a.density, 150, 1)
    indptr, indices, data = spec.cells(0, a.cells)
    counts = sp.csc_matrix((data, indices, indptr), shape=(a.genes, a.cells))
    del spec
    ref_d = None
    ok_all = True
    for tw in [int(x) for x in a.tile_w.split(",")]:
        os.environ["M3B_TILE_W"] = str(tw)
        ctx = engine.Context(0, small_svd="device")
        dev = engine.DeviceMatrix.from_csc(ctx, counts)
        sf, _ = engine.size_factors_from_totals(dev.cell_totals(True))
        dev.normalize(sf, L.NORM_LOG2, 1.0)
        for var in [int(x) for x in a.variants.split(",")]:
            os.environ["M3B_FLAT_VARIANT"] = str(var)
            dev.select_genes(None)
            v0 = rnorm_after_seed(2016, dev.n_kept)
            nv = min(a.num_dim, min(dev.n_kept, a.cells) - 1)
            dev.pca_irlba(nv, v0, maxit=2, want_x=False)  # warm-up
            ctx.set_flags(L.FLAG_TIME_SPMV)
            ctx.reset_stats()
            r = dev.pca_irlba(nv, v0, maxit=a.maxit, want_x=False)
            st = ctx.stats()
            ctx.set_flags(0)
            nA, nB = st["spmv_cells_out"], st["spmv_genes_out"]
            usA = 1e3 * st["spmv_cells_out_ms"] / max(1, nA)
            usB = 1e3 * st["spmv_genes_out_ms"] / max(1, nB)
            d = np.asarray(r["d"][:5])
            if ref_d is None:
                ref_d = d
            rel = float(np.max(np.abs(d - ref_d) / np.abs(ref_d)))
            ok = rel < 1e-10
            ok_all &= ok
            print("SPMV_SWEEP " + json.dumps({
                "cells": a.cells, "genes": a.genes, "nnz": int(counts.nnz), "tile_w": tw, "variant": var,
                "us_cells_out": round(usA, 2), "us_genes_out": round(usB, 2),
                "phys_gbs_cells_out": round(st["spmv_stream_bytes_cells_out"] / max(usA, 1e-9) / 1e3, 1),
                "phys_gbs_genes_out": round(st["spmv_stream_bytes_genes_out"] / max(usB, 1e-9) / 1e3, 1),
                "launches": [int(nA), int(nB)], "mprod": r["mprod"], "d_rel_vs_first": rel, "ok": ok,
                "irlba_ms": round(st["irlba_ms"], 2)}), flush=True)
        dev.free()
        ctx.close()
    print("SPMV_SWEEP_OK %s" % ok_all)
    sys.exit(0 if ok_all else 1)


if __name__ == "__main__":
    main()
