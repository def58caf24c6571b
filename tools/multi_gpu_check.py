#!/usr/bin/env python
"""Cell-sharded parity check (run under torchrun, one rank per GPU):

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
      --master-port 29511 tools/multi_gpu_check.py [fixture|synth20k] [num_dim]

Every rank takes its contiguous cell range (cut at nnz quantiles) of a committed fixture or of
a seeded synthetic matrix (synth20k: 20 000 cells x 4 000 genes, ~6 % dense, 40 planted types),
runs estimate_size_factors + preprocess_cds through the public API with the cross-rank sums
inside libm3b.so, and rank 0 compares the gathered embedding / loadings with the oracle on the
WHOLE matrix: size factors array_equal, iter / mprod equal to the oracle's, singular values
<= 1e-6 relative, subspace angles < 1e-4 rad, replicated gene-space results bit-identical on
every rank.  tests/test_gpu_multi.py drives it; exit status 0 = all of the above hold."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import scipy.sparse as sp
    import torch
    import torch.distributed as td
    import monocle3_b200 as m3
    from monocle3_b200 import engine, dist
    from tests.conftest import load_fixture
    from tests.parity import subspace_angle, sv_rel_err
    name = sys.argv[1] if len(sys.argv) > 1 else "worm_l2"
    nd = int(sys.argv[2]) if len(sys.argv) > 2 else 50
    lr = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(lr)
    td.init_process_group("nccl", device_id=torch.device("cuda", lr))
    rank, world = td.get_rank(), td.get_world_size()
    if name == "invariant":
        return invariant_case(td, lr)
    if name == "synth20k":
        from monocle3_b200.synth import synth_csc
        # generated on the CPU so that every rank (and the oracle) sees the same matrix
        counts = synth_csc(4000, 20000, 0.06, 40, seed=5, device="cpu")
    else:
        counts, _ = load_fixture(name)
    ranges = dist.shard_ranges_by_nnz(counts.indptr, world)
    lo, hi = ranges[rank]
    ctx = engine.Context(device=lr, small_svd="device")
    m3.set_context(ctx)
    dist.init_comm(ctx)
    cds = m3.new_cell_data_set(counts[:, lo:hi], shard=dist.make_shard(counts.shape[1]))
    cds = m3.preprocess_cds(cds, num_dim=nd)
    model = cds.reduce_dim_aux["PCA"]["model"]
    x_local = np.ascontiguousarray(cds.reduced_dims["PCA"])
    ag = dist.make_allgather()
    # gather the embedding row blocks (column by column keeps the helper 1-D)
    X = np.stack([ag(x_local[:, j].copy()) for j in range(x_local.shape[1])], axis=1)
    sf_all = ag(m3.size_factors(cds))
    ok = True
    if rank == 0:
        from oracle import monocle3_ref as ref
        sf_o = ref.estimate_size_factors(counts)
        o = ref.preprocess_cds(counts, sf_o, num_dim=nd)
        ir = cds.reduce_dim_aux["PCA"]["irlba"]
        res = dict(world=world, fixture=name, num_dim=nd, cells=int(counts.shape[1]), sf_exact=bool(np.array_equal(sf_all, sf_o)),
                   iter=(ir["iter"], o["iter"]), mprod=(ir["mprod"], o["mprod"]),
                   sv_rel=sv_rel_err(ir["d"], o["d"]), angle_v=subspace_angle(model["svd_v"], o["svd_v"]),
                   angle_x=subspace_angle(X, o["PCA"]))
        ok = (res["sf_exact"] and res["sv_rel"] < 1e-6 and res["angle_v"] < 1e-4 and res["angle_x"] < 1e-4
              and ir["iter"] == o["iter"] and ir["mprod"] == o["mprod"])
        res["ok"] = bool(ok)
        print("MULTI_GPU_CHECK " + json.dumps(res), flush=True)
    # replicated gene-space state must be bit-identical on every rank
    v = np.ascontiguousarray(model["svd_v"]).ravel()
    vs = ag(np.array([float(np.sum(v)), float(np.sum(np.abs(v))), float(v[0]), float(v[-1])]))
    same = all(np.array_equal(vs[:4], vs[4 * r:4 * r + 4]) for r in range(world))
    if rank == 0:
        print("REPLICAS_IDENTICAL %s" % same, flush=True)
    td.destroy_process_group()
    sys.exit(0 if (ok and same) else 1)


def invariant_case(td, lr):
    """irlba's answer to an exact invariant subspace (a vanished Lanczos vector is replaced by rnorm()
    from R's stream), cell-sharded: every rank continues the same stream, cell-length draws are
    sliced by the library.  Same iter / mprod / singular values as the restatement on the whole
    matrix (the single-GPU twin is tests/test_gpu_parity.py::test_invariant_subspace_random_restart)."""
    import scipy.sparse as sp
    from monocle3_b200 import engine, dist, _lib as L
    from monocle3_b200.r_rng import RStream
    from tests.parity import subspace_angle
    rank, world = td.get_rank(), td.get_world_size()
    rng = np.random.default_rng(5)
    G, Cn, r = 40, 70, 4
    dense = (rng.integers(1, 4, (G, r)) @ rng.integers(0, 3, (r, Cn))).astype(np.float64) * 1e-9
    counts = sp.csc_matrix(dense)
    lo, hi = dist.shard_ranges_even(Cn, world)[rank]
    ctx = engine.Context(device=lr, small_svd="device")
    dist.init_comm(ctx)
    dev = engine.DeviceMatrix.from_csc(ctx, counts[:, lo:hi], n_cells_total=Cn)
    dev.normalize(None, L.NORM_NONE, 0.0)
    keep = dev.select_genes(None)
    n = int(keep.sum())
    stream = RStream(7)
    v0 = stream.rnorm(n)
    ctx.set_rnorm(stream.rnorm)
    res = dev.pca_irlba(6, v0, center=False, scale=False)
    ctx.set_rnorm(None)
    ok = True
    if rank == 0:
        from oracle import irlba_ref
        from oracle.r_rng import RRandom
        orng = RRandom(7)
        v0o = orng.rnorm(n)
        o = irlba_ref.irlba(sp.csc_matrix(dense[keep, :].T), 6, v0o, rng=orng)
        out = dict(world=world, fixture="invariant", status=int(res["status"]), iter=(res["iter"], o.iter), mprod=(res["mprod"], o.mprod),
                   sv_rel=float(np.max(np.abs(res["d"][:r] - o.d[:r]) / o.d[:r])),
                   angle_v=subspace_angle(res["v"][:, :r], o.v[:, :r]))
        ok = (res["status"] == 0 and o.err == 0 and (res["iter"], res["mprod"]) == (o.iter, o.mprod)
              and out["sv_rel"] < 1e-6 and out["angle_v"] < 1e-4)
        out["ok"] = bool(ok)
        print("MULTI_GPU_CHECK " + json.dumps(out), flush=True)
    ag = dist.make_allgather()
    v = np.ascontiguousarray(res["v"]).ravel()
    vs = ag(np.array([float(np.sum(v)), float(np.sum(np.abs(v))), float(v[0]), float(v[-1])]))
    same = all(np.array_equal(vs[:4], vs[4 * q:4 * q + 4]) for q in range(world))
    if rank == 0:
        print("REPLICAS_IDENTICAL %s" % same, flush=True)
    td.destroy_process_group()
    sys.exit(0 if (ok and same) else 1)


if __name__ == "__main__":
    main()
