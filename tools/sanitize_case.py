#!/usr/bin/env python
"""Small end-to-end case for compute-sanitizer (memcheck / racecheck): a549 through
size factors, normalize, both layouts, moments, IRLBA (structured + Jacobi SVD), projection, LSI."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import monocle3_b200 as m3
from monocle3_b200 import engine, _lib as L

for flags in (0, L.FLAG_JACOBI_ONLY):
    m3.set_context(engine.Context(0, flags=flags, small_svd="device"))
    cds = m3.load_a549()
    cds = m3.preprocess_cds(cds, num_dim=10)
    q = m3.load_a549()
    q.reduce_dim_aux = {"PCA": cds.reduce_dim_aux["PCA"]}
    q = m3.preprocess_transform(q)
    print("flags", flags, "iter", cds.reduce_dim_aux["PCA"]["irlba"]["iter"],
          "max |transform - embedding|", float(np.abs(q.reduced_dims["PCA"] - cds.reduced_dims["PCA"]).max()))
cds = m3.preprocess_cds(m3.load_a549(), method="LSI", num_dim=10)
print("LSI[1,1]", cds.reduced_dims["LSI"][0, 0])
nc = m3.normalized_counts(m3.load_a549())
print("normalized_counts nnz", nc.nnz)
