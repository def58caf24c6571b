#!/usr/bin/env python
"""BASELINE-sized parity: one GPU run of a named bench workload against the oracle ON THE SAME
MATRIX at full size, and the stored expectation the sharded runs are checked against.

  python tools/fullsize_parity.py --workload cfg2|cfg3 [--no-oracle] [--out gpurun_out]

1. generates the workload's synthetic matrix exactly as bench.py does (same generator / seed /
   device for the random streams),
2. runs estimate_size_factors + preprocess_cds through the public API on cuda:0 and writes
   <out>/expected_<workload>.json  {iter, mprod, d[]}  (copied to profiles/ once checked),
3. runs the oracle's preprocess_cds on the host cores on the whole matrix (minutes) and writes
   <out>/fullsize_parity_<workload>.json: size factors array_equal, iter / mprod equality,
   singular values (bar 1e-6 relative), subspace angles of loadings and embedding (bar 1e-4 rad).
Exit status 0 = all bars met."""
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
    ap.add_argument("--workload", default="cfg2")
    ap.add_argument("--no-oracle", action="store_true")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out"))
    args = ap.parse_args()
    import scipy.sparse as sp
    import bench
    import monocle3_b200 as m3
    from monocle3_b200 import engine
    from monocle3_b200.synth import SynthSpec
    from tests.parity import subspace_angle, sv_rel_err, column_angles
    wl = dict(bench.WORKLOADS[args.workload], key=args.workload)
    os.makedirs(args.out, exist_ok=True)
    spec = SynthSpec(wl["n_genes"], wl["cells"], wl["density"], wl["n_types"], wl["seed"])
    indptr, indices, data = spec.cells(0, wl["cells"])
    counts = sp.csc_matrix((data, indices, indptr.astype(np.int32)), shape=(wl["n_genes"], wl["cells"]))
    m3.set_context(engine.Context(device=0, small_svd="device"))
    t0 = time.perf_counter()
    cds = m3.new_cell_data_set(counts)
    cds = m3.preprocess_cds(cds, num_dim=wl["num_dim"])
    t_gpu = time.perf_counter() - t0
    ir = cds.reduce_dim_aux["PCA"]["irlba"]
    model = cds.reduce_dim_aux["PCA"]["model"]
    exp = {"workload": args.workload, "cells": wl["cells"], "genes": wl["n_genes"], "num_dim": wl["num_dim"], "n_gpus": 1,
           "nnz": int(counts.nnz), "iter": int(ir["iter"]), "mprod": int(ir["mprod"]), "d": [float(v) for v in ir["d"]]}
    json.dump(exp, open(os.path.join(args.out, "expected_%s.json" % args.workload), "w"))
    print("GPU run: %.2f s, iter %d, mprod %d" % (t_gpu, ir["iter"], ir["mprod"]), flush=True)
    if args.no_oracle:
        return 0
    from oracle import monocle3_ref as ref
    t0 = time.perf_counter()
    sf_o = ref.estimate_size_factors(counts)
    o = ref.preprocess_cds(counts, sf_o, num_dim=wl["num_dim"])
    t_cpu = time.perf_counter() - t0
    ang_v = column_angles(model["svd_v"], o["svd_v"])
    res = {"workload": args.workload, "cells": wl["cells"], "genes": wl["n_genes"], "nnz": int(counts.nnz), "num_dim": wl["num_dim"],
           "oracle_s": round(t_cpu, 1), "oracle_cores": os.cpu_count(), "gpu_e2e_s": round(t_gpu, 3),
           "sf_exact": bool(np.array_equal(m3.size_factors(cds), sf_o)),
           "kept_genes_equal": bool(np.array_equal(np.asarray(model["svd_v_rownames"]).astype(str),
                                                   np.arange(wl["n_genes"])[o["kept_genes"]].astype(str))),
           "iter": [int(ir["iter"]), int(o["iter"])], "mprod": [int(ir["mprod"]), int(o["mprod"])],
           "sv_rel": sv_rel_err(ir["d"], o["d"]),
           "angle_v_subspace": subspace_angle(model["svd_v"], o["svd_v"]),
           "angle_x_subspace": subspace_angle(cds.reduced_dims["PCA"], o["PCA"]),
           "angle_v_max_column": float(np.max(ang_v)),
           "center_max_abs": float(np.max(np.abs(model["svd_center"] - o["svd_center"]))),
           "scale_max_rel": float(np.max(np.abs(model["svd_scale"] - o["svd_scale"]) / o["svd_scale"]))}
    res["ok"] = bool(res["sf_exact"] and res["kept_genes_equal"] and res["iter"][0] == res["iter"][1]
                     and res["mprod"][0] == res["mprod"][1] and res["sv_rel"] < 1e-6
                     and res["angle_v_subspace"] < 1e-4 and res["angle_x_subspace"] < 1e-4)
    json.dump(res, open(os.path.join(args.out, "fullsize_parity_%s.json" % args.workload), "w"), indent=1)
    print("FULLSIZE_PARITY " + json.dumps(res), flush=True)
    return 0 if res["ok"] else 1


if __name__ == "__main__":
    sys.exit(main())
