"""Cell sharding across GPUs: one process per GPU, torch.distributed for the plumbing.

The data path needs exactly three kinds of exchange:
  * host side, once: all-gather of the per-cell totals so every rank can finish the
    geometric mean of estimate_size_factors (R/utils.R:61-62) identically;
  * device side, per Lanczos step: sum-allreduces issued by libm3b.so itself over NCCL
    (comm.cpp) -- torch only transports the 128-byte NCCL unique id;
  * nothing else: embeddings stay sharded, gene-space results are replicated.
"""
import numpy as np


def shard_ranges_by_nnz(indptr, world):
    """Cut the cells into `world` contiguous ranges at nnz quantiles of the CSC column
    pointer (balances the SpMV stream, not the cell count).  Returns [(lo, hi)] * world."""
    indptr = np.asarray(indptr, dtype=np.int64)
    n = len(indptr) - 1
    total = int(indptr[-1])
    cuts = [0]
    for r in range(1, world):
        target = total * r / world
        c = int(np.searchsorted(indptr, target, side="left"))
        c = min(max(c, cuts[-1]), n)
        cuts.append(c)
    cuts.append(n)
    return [(cuts[r], cuts[r + 1]) for r in range(world)]


def shard_ranges_even(n_cells, world):
    base, rem = divmod(n_cells, world)
    out, lo = [], 0
    for r in range(world):
        hi = lo + base + (1 if r < rem else 0)
        out.append((lo, hi))
        lo = hi
    return out


def make_allgather(group=None):
    """allgather(np.ndarray float64, variable length) -> concatenation in rank order, via
    torch.distributed (gloo or nccl).  Works for world size 1 without torch."""
    try:
        import torch
        import torch.distributed as td
    except ImportError:  # pragma: no cover
        td = None
    if td is None or not td.is_available() or not td.is_initialized():
        return lambda a: np.asarray(a, dtype=np.float64)
    world = td.get_world_size(group)
    backend = td.get_backend(group)
    dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")

    def allgather(a):
        a = np.ascontiguousarray(a, dtype=np.float64)
        n = torch.tensor([a.size], dtype=torch.int64, device=dev)
        sizes = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
        td.all_gather(sizes, n, group=group)
        sizes = [int(s.item()) for s in sizes]
        mx = max(sizes + [1])
        buf = torch.zeros(mx, dtype=torch.float64, device=dev)
        buf[:a.size] = torch.from_numpy(a).to(dev)
        outs = [torch.zeros(mx, dtype=torch.float64, device=dev) for _ in range(world)]
        td.all_gather(outs, buf, group=group)
        return np.concatenate([o[:s].cpu().numpy() for o, s in zip(outs, sizes)])

    return allgather


def init_comm(ctx, group=None):
    """Join this rank's context to the NCCL communicator libm3b.so uses for its
    allreduces.  Rank 0 creates the unique id; torch.distributed broadcasts it."""
    import torch.distributed as td
    rank, world = td.get_rank(group), td.get_world_size(group)
    obj = [ctx.comm_unique_id() if rank == 0 else None]
    td.broadcast_object_list(obj, src=0, group=group)
    ctx.comm_init(rank, world, obj[0])
    # peer-memory exchange for the fused per-step reduction (one box; M3B_NO_P2P=1 keeps NCCL only)
    import os
    if world > 1 and not os.environ.get("M3B_NO_P2P"):
        handles = [None] * world
        td.all_gather_object(handles, ctx.comm_p2p_handle(), group=group)
        ctx.comm_p2p_open(handles)
    return rank, world


def make_shard(n_cells_total, group=None):
    """The `shard` tuple CellDataSet expects: (rank, world, n_cells_total, allgather)."""
    import torch.distributed as td
    return (td.get_rank(group), td.get_world_size(group), int(n_cells_total), make_allgather(group))
