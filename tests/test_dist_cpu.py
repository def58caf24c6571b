"""World-size-2 gloo tests (CPU) of the host-side logic of the cell-sharded path: shard
cuts, the variable-length all-gather, and the sharded size-factor finish being bit-identical
to the unsharded one.  (The device-side allreduces are covered on real GPUs by
tools/multi_gpu_check.py.)"""
import os
import socket

import numpy as np
import pytest

from oracle import monocle3_ref as ref


def test_shard_ranges_by_nnz(worm_l2):
    from monocle3_b200 import dist
    counts, _ = worm_l2
    for world in (1, 2, 3, 8):
        r = dist.shard_ranges_by_nnz(counts.indptr, world)
        assert r[0][0] == 0 and r[-1][1] == counts.shape[1]
        assert all(a[1] == b[0] for a, b in zip(r[:-1], r[1:]))
        nnz = [counts.indptr[hi] - counts.indptr[lo] for lo, hi in r]
        assert max(nnz) - min(nnz) <= 2 * np.diff(counts.indptr).max()  # balanced to within a cell or two
    assert dist.shard_ranges_even(10, 3) == [(0, 4), (4, 7), (7, 10)]


def test_synth_is_shard_invariant():
    from monocle3_b200.synth import synth_csc
    full = synth_csc(300, 2500, 0.07, 20, seed=3, device="cpu")
    part = synth_csc(300, 2500, 0.07, 20, seed=3, lo=900, hi=2300, device="cpu")
    assert abs(full[:, 900:2300] - part).sum() == 0
    assert 0.05 < full.nnz / (300 * 2500) < 0.09
    assert np.all(np.asarray(full.sum(axis=0)).ravel() > 0)  # no empty cells


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    import torch.distributed as td
    from monocle3_b200 import dist, engine
    from tests.conftest import load_fixture
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    td.init_process_group("gloo", rank=rank, world_size=world)
    counts, _ = load_fixture("worm_l2")
    lo, hi = dist.shard_ranges_by_nnz(counts.indptr, world)[rank]
    shard = dist.make_shard(counts.shape[1])
    assert shard[:3] == (rank, world, counts.shape[1])
    local = counts[:, lo:hi]
    # local totals (on the GPU these come from K1; any exact integer sum will do here)
    tot_local = np.asarray(np.round(local).sum(axis=0)).ravel()
    tot_all = shard[3](tot_local)                      # variable-length all-gather
    sf_all, any_zero = engine.size_factors_from_totals(tot_all)
    n_per_rank = shard[3](np.array([hi - lo], dtype=np.float64)).astype(int)
    off = int(n_per_rank[:rank].sum())
    np.save(os.path.join(out_dir, "sf_%d.npy" % rank), sf_all[off:off + hi - lo])
    np.save(os.path.join(out_dir, "tot_%d.npy" % rank), tot_all)
    td.destroy_process_group()


def test_sharded_size_factors_gloo(tmp_path, worm_l2):
    import torch.multiprocessing as mp
    counts, _ = worm_l2
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    sf = np.concatenate([np.load(tmp_path / ("sf_%d.npy" % r)) for r in range(world)])
    sf_o, tot_o = ref.estimate_sf_sparse(counts)
    assert np.array_equal(np.load(tmp_path / "tot_0.npy"), tot_o)
    assert np.array_equal(np.load(tmp_path / "tot_1.npy"), tot_o)
    assert np.array_equal(sf, sf_o)  # sharded == unsharded, bit for bit
