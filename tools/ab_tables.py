#!/usr/bin/env python
"""A/B of the device-side table builder (default) against the host builders (M3B_HOST_TABLES=1):
the same matrix through m3b_select_genes + m3b_pca_irlba twice.  Items, slots and flat pieces are
the same by construction (tests/test_tables_host.py), so singular values, loadings and embedding
must be BIT-IDENTICAL; prints the wall time of m3b_select_genes for both."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import monocle3_b200 as m3
    from monocle3_b200 import engine, _lib as L
    from monocle3_b200.synth import synth_csc
    from monocle3_b200.r_rng import rnorm_after_seed
    cases = [("a: 60k x 30k 6% (3 cell tiles, 2 gene tiles)", 30000, 60000, 0.06, 40, 30),
             ("b: 42k x 20k 8% (cfg2 shape)", 20000, 42000, 0.08, 150, 50),
             ("c: 3k x 2k 5% with dropped genes", 2000, 3000, 0.05, 10, 10)]
    ctx = engine.Context(device=0, small_svd="device")
    m3.set_context(ctx)
    out = []
    for name, G, C, dens, K, nd in cases:
        counts = synth_csc(G, C, dens, K, seed=11)
        if name.startswith("c"):
            lil = counts.tolil()
            for g in (5, 6, 700, 1999):   # genes without any count: dropped by the zero-gene filter
                lil[g, :] = 0
            counts = lil.tocsc()
        cds = m3.new_cell_data_set(counts)
        dev = cds.device_matrix()
        sf = m3.size_factors(cds)
        res = {}
        for mode in ("device", "host", "oldscatter"):
            os.environ.pop("M3B_HOST_TABLES", None)
            os.environ.pop("M3B_OLD_SCATTER", None)
            if mode == "host":
                os.environ["M3B_HOST_TABLES"] = "1"
            if mode == "oldscatter":      # device tables, but the round-1 scatter / count kernels
                os.environ["M3B_OLD_SCATTER"] = "1"
            dev.normalize(sf, L.NORM_LOG2, 1.0)
            dev.select_genes(None)          # warm
            dev.normalize(sf, L.NORM_LOG2, 1.0)
            t0 = time.perf_counter()
            keep = dev.select_genes(None)
            t_sel = time.perf_counter() - t0
            v0 = rnorm_after_seed(2016, dev.n_kept)
            r = dev.pca_irlba(nd, v0, center=True, scale=True)
            res[mode] = (r, keep, t_sel, ctx.stats()["select_ms"])
        os.environ.pop("M3B_HOST_TABLES", None)
        os.environ.pop("M3B_OLD_SCATTER", None)
        a, b = res["device"][0], res["host"][0]
        c_old = res["oldscatter"][0]
        rec = dict(case=name, kept=int(dev.n_kept), iter=(a["iter"], b["iter"]), mprod=(a["mprod"], b["mprod"]),
                   d_identical=bool(np.array_equal(a["d"], b["d"])), v_identical=bool(np.array_equal(a["v"], b["v"])),
                   x_identical=bool(np.array_equal(a["x"], b["x"])), keep_identical=bool(np.array_equal(res["device"][1], res["host"][1])),
                   d_max_rel=float(np.max(np.abs(a["d"] - b["d"]) / b["d"])),
                   new_vs_old_scatter_identical=bool(np.array_equal(a["d"], c_old["d"]) and np.array_equal(a["x"], c_old["x"])),
                   new_vs_old_scatter_d_max_rel=float(np.max(np.abs(a["d"] - c_old["d"]) / c_old["d"])),
                   select_ms_oldscatter=round(1e3 * res["oldscatter"][2], 2),
                   select_ms_device=round(1e3 * res["device"][2], 2), select_ms_host=round(1e3 * res["host"][2], 2),
                   select_phases_device=[round(v, 2) for v in res["device"][3]], select_phases_host=[round(v, 2) for v in res["host"][3]])
        print("AB_TABLES " + json.dumps(rec), flush=True)
        out.append(rec)
        dev.free()
    # bit-identical wherever the two builders produce the same plans; the host builder makes no side-stream
    # plan for multi-tile layouts (case a), where the first product of a cycle is then summed in another order
    ok = all((r["d_identical"] and r["v_identical"] and r["x_identical"] or r["d_max_rel"] < 1e-13) and r["keep_identical"]
             and r["new_vs_old_scatter_d_max_rel"] < 1e-13 for r in out)
    print("AB_TABLES_OK %s" % ok)
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
