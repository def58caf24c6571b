import sys; sys.path.insert(0,'/root/repo')
import numpy as np, monocle3_b200 as m3
from monocle3_b200 import engine
m3.set_context(engine.Context(0, small_svd="device"))
cds=m3.load_a549(); cds=m3.preprocess_cds(cds,num_dim=20)
st=m3.get_context().stats(); print({k:st[k] for k in ("svd_fallbacks","svd_fallback_reasons","svd_sweeps")}, cds.reduce_dim_aux["PCA"]["irlba"]["iter"])
