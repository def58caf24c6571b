"""Generate tests/golden/*.npz from the reference's bundled fixtures.

TEST INFRASTRUCTURE ONLY.  Run in the build container (where /root/reference is
mounted); the GPU box never reads /root/reference -- it uses the committed .npz.

  python -m oracle.make_golden

Inputs (read-only):  /root/reference/inst/extdata/small_a549_dex_{exprs,pdata,fdata}.rda
                     (what load_a549() reads, R/io.R:10-26),
                     worm_l2/worm_l2_expression_matrix.rds, worm_embryo/...rds.
Outputs: tests/golden/a549.npz (p,i,x,dim + gene/cell names),
         tests/golden/worm_l2.npz, tests/golden/worm_embryo.npz (p,i,x,dim; values
         are integers and stored as uint16 to keep the files small),
         tests/golden/a549_oracle.npz (oracle outputs on a549, used by GPU parity
         tests as precomputed vectors in addition to calling the oracle live).
"""
import os
import sys

import numpy as np
import scipy.sparse as sp

from . import rds, monocle3_ref as ref

REF = "/root/reference/inst/extdata"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def main():
    os.makedirs(OUT, exist_ok=True)
    p, i, x, dim, rn, cn = rds.dgcmatrix(rds.read_rds(f"{REF}/small_a549_dex_exprs.rda"))
    pd_rn = rds.dataframe_rownames(rds.read_rds(f"{REF}/small_a549_dex_pdata.rda"))
    fd_rn = rds.dataframe_rownames(rds.read_rds(f"{REF}/small_a549_dex_fdata.rda"))
    # load_a549 reorders columns by the pdata row names (R/io.R:20); identity here.
    assert pd_rn == cn and fd_rn == rn
    np.savez_compressed(f"{OUT}/a549.npz", p=p, i=i, x=x, dim=np.array(dim),
                        genes=np.array(rn), cells=np.array(cn))
    for name, path in (("worm_l2", "worm_l2/worm_l2_expression_matrix.rds"),
                       ("worm_embryo", "worm_embryo/worm_embryo_expression_matrix.rds")):
        p2, i2, x2, dim2, _, _ = rds.dgcmatrix(rds.read_rds(f"{REF}/{path}"))
        assert np.all(x2 == np.round(x2)) and x2.max() < 65536 and x2.min() >= 0
        np.savez_compressed(f"{OUT}/{name}.npz", p=p2, i=i2, x=x2.astype(np.uint16), dim=np.array(dim2))
    # oracle outputs on a549 (log, num_dim 20 and 50)
    counts = sp.csc_matrix((x, i, p), shape=dim)
    sf = ref.estimate_size_factors(counts)
    out = {"sf": sf}
    for nd in (20, 50):
        r = ref.preprocess_cds(counts, sf, num_dim=nd)
        out[f"d{nd}"] = r["d"]
        out[f"v{nd}"] = r["svd_v"]
        out[f"x{nd}"] = r["PCA"]
        out[f"iter{nd}"] = np.array([r["iter"], r["mprod"]])
        out[f"center{nd}"] = r["svd_center"]
        out[f"scale{nd}"] = r["svd_scale"]
        out[f"kept{nd}"] = r["kept_genes"]
    np.savez_compressed(f"{OUT}/a549_oracle.npz", **out)
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    sys.exit(main())
