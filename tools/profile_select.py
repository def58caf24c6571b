#!/usr/bin/env python
"""m3b_select_genes (gene filter + dual layout build + plans) alone on a bench workload's matrix:
prints the host wall time per phase (m3b_stats.select_ms) for a few repetitions.  Run it under
`ncu --metrics gpu__time_duration.sum` for the per-kernel launch list of the layout build."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="cfg3")
    ap.add_argument("--cells", type=int, default=0)
    ap.add_argument("--reps", type=int, default=3)
    a = ap.parse_args()
    import scipy.sparse as sp
    import bench
    import monocle3_b200 as m3
    from monocle3_b200 import engine, _lib as L
    from monocle3_b200.synth import SynthSpec
    wl = dict(bench.WORKLOADS[a.workload])
    if a.cells:
        wl["cells"] = a.cells
    spec = SynthSpec(wl["n_genes"], wl["cells"], wl["density"], wl["n_types"], wl["seed"])
    indptr, indices, data = spec.cells(0, wl["cells"])
    counts = sp.csc_matrix((data, indices, indptr.astype(np.int32)), shape=(wl["n_genes"], wl["cells"]))
    ctx = engine.Context(device=0, small_svd="device")
    m3.set_context(ctx)
    cds = m3.new_cell_data_set(counts)
    dev = cds.device_matrix()
    sf = m3.size_factors(cds)
    for r in range(a.reps):
        dev.normalize(sf, L.NORM_LOG2, 1.0)
        t0 = time.perf_counter()
        dev.select_genes(None)
        dt = 1e3 * (time.perf_counter() - t0)
        print("SELECT " + json.dumps({"rep": r, "cells": wl["cells"], "nnz": int(counts.nnz), "ms": round(dt, 2),
                                      "phases": [round(v, 2) for v in ctx.stats()["select_ms"]],
                                      "env": {k: v for k, v in os.environ.items() if k.startswith("M3B_")}}), flush=True)


if __name__ == "__main__":
    main()
