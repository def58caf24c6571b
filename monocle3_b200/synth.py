"""Deterministic synthetic count matrices for bench.py and the parity tests.

Value model (calibrated to the reference's bundled fixtures):
  x_gc ~ Poisson(lambda_gc),  lambda_gc = alpha * b_g * t[type(c), g] * d_c
  b_g ~ LogNormal(0, 1.2)   gene abundance (skewed occupancy)
  d_c ~ LogNormal(0, 0.5)   cell depth
  t   : K planted cell types, each with a sparse multiplicative program
        (10% of the genes, log-fold ~ N(0,1))
  alpha solved so that the realised density E[x > 0] hits the target.
Cells are generated in fixed blocks of BLOCK cells, each block with its own seed, so
any rank can generate any cell range and the global matrix does not depend on how the
cells are sharded.  Uses torch only as a random-number engine (GPU when available).
"""
import numpy as np
import torch

BLOCK = 1024


class SynthSpec:
    def __init__(self, n_genes, n_cells, density, n_types, seed=1, device=None):
        self.G, self.C, self.density, self.K, self.seed = int(n_genes), int(n_cells), float(density), int(n_types), int(seed)
        self.device = torch.device(device) if device is not None else (
            torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else torch.device("cpu"))
        g = torch.Generator(device="cpu")
        g.manual_seed(self.seed * 7919 + 13)
        self.log_b = (torch.randn(self.G, generator=g, dtype=torch.float64) * 1.2)
        prog = torch.rand(self.K, self.G, generator=g) < 0.10
        lf = torch.randn(self.K, self.G, generator=g, dtype=torch.float64)
        self.log_t = torch.where(prog, lf, torch.zeros_like(lf))
        self.alpha = self._solve_alpha()
        self.log_b_dev = self.log_b.to(self.device)
        self.log_t_dev = self.log_t.to(self.device)

    def _cell_params(self, block):
        g = torch.Generator(device="cpu")
        g.manual_seed(self.seed * 1000003 + 97 * block + 1)
        types = torch.randint(0, self.K, (BLOCK,), generator=g)
        log_d = torch.randn(BLOCK, generator=g, dtype=torch.float64) * 0.5
        return types, log_d

    def _solve_alpha(self):
        # density as a function of alpha on a fixed sample of cells (bisection in log space)
        types, log_d = self._cell_params(0)
        n = min(BLOCK, 256)
        base = self.log_b[None, :] + self.log_t[types[:n]] + log_d[:n, None]
        lo, hi = -30.0, 10.0
        for _ in range(60):
            mid = 0.5 * (lo + hi)
            dens = float((1.0 - torch.exp(-torch.exp(base + mid))).mean())
            if dens < self.density:
                lo = mid
            else:
                hi = mid
        return 0.5 * (lo + hi)

    def block_csc(self, block):
        """(counts per cell [<=BLOCK], gene indices int32, values float64) of one cell block."""
        c0 = block * BLOCK
        nb = min(BLOCK, self.C - c0)
        types, log_d = self._cell_params(block)
        dev = self.device
        lam = torch.exp(self.log_b_dev[None, :] + self.log_t_dev[types[:nb].to(dev)] + (log_d[:nb].to(dev) + self.alpha)[:, None])
        g = torch.Generator(device=dev)
        g.manual_seed(self.seed * 1000003 + 97 * block + 2)
        x = torch.poisson(lam.to(torch.float32), generator=g)
        # no empty cells: a cell without reads gets one read of its most likely gene
        empty = x.sum(dim=1) == 0
        if bool(empty.any()):
            top = lam.argmax(dim=1)
            rows = torch.nonzero(empty).ravel()
            x[rows, top[rows]] = 1.0
        nz = torch.nonzero(x)  # sorted by (cell, gene) => CSC order
        vals = x[nz[:, 0], nz[:, 1]].to(torch.float64)
        per_cell = torch.bincount(nz[:, 0], minlength=nb)
        return per_cell.cpu().numpy().astype(np.int64), nz[:, 1].to(torch.int32).cpu().numpy(), vals.cpu().numpy()

    def cells(self, lo, hi):
        """CSC (indptr int64, indices int32, data float64) of the cell range [lo, hi)."""
        counts, idxs, vals = [], [], []
        for b in range(lo // BLOCK, (hi + BLOCK - 1) // BLOCK):
            pc, gi, v = self.block_csc(b)
            c0 = b * BLOCK
            a, z = max(lo, c0) - c0, min(hi, c0 + len(pc)) - c0
            off = np.concatenate([[0], np.cumsum(pc)])
            counts.append(pc[a:z])
            idxs.append(gi[off[a]:off[z]])
            vals.append(v[off[a]:off[z]])
        pc = np.concatenate(counts) if counts else np.zeros(0, dtype=np.int64)
        indptr = np.concatenate([[0], np.cumsum(pc)]).astype(np.int64)
        return indptr, (np.concatenate(idxs) if idxs else np.zeros(0, np.int32)), (np.concatenate(vals) if vals else np.zeros(0))


def synth_csc(n_genes, n_cells, density, n_types, seed=1, lo=0, hi=None, device=None):
    """scipy.sparse.csc_matrix (genes x cells[lo:hi]) of the synthetic matrix."""
    import scipy.sparse as sp
    spec = SynthSpec(n_genes, n_cells, density, n_types, seed, device)
    hi = n_cells if hi is None else hi
    indptr, indices, data = spec.cells(lo, hi)
    return sp.csc_matrix((data, indices, indptr), shape=(n_genes, hi - lo))


def _concat_csc(parts):
    """Concatenate cell ranges given as (indptr, indices, data) triples."""
    parts = [p for p in parts if len(p[0]) > 1]
    if not parts:
        return np.zeros(1, dtype=np.int64), np.zeros(0, np.int32), np.zeros(0)
    counts = np.concatenate([np.diff(p[0]) for p in parts])
    indptr = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
    return indptr, np.concatenate([p[1] for p in parts]), np.concatenate([p[2] for p in parts])


def balanced_shard(spec, rank, world, allgather=None):
    """This rank's contiguous cell range of the ONE matrix `spec` describes, cut at nnz quantiles
    of the global column pointer (dist.shard_ranges_by_nnz).  Every rank
    generates its equal-cell range, the per-cell counts are all-gathered (`allgather`: see
    dist.make_allgather), the cuts are computed identically everywhere, and the few cells that
    moved across a cut are generated on top.  Returns (lo, hi, indptr, indices, data)."""
    from .dist import shard_ranges_by_nnz, shard_ranges_even
    if world == 1:
        indptr, idx, val = spec.cells(0, spec.C)
        return 0, spec.C, indptr, idx, val
    lo0, hi0 = shard_ranges_even(spec.C, world)[rank]
    indptr, idx, val = spec.cells(lo0, hi0)
    all_counts = np.rint(allgather(np.diff(indptr).astype(np.float64))).astype(np.int64)
    gptr = np.concatenate([[0], np.cumsum(all_counts)])
    lo, hi = shard_ranges_by_nnz(gptr, world)[rank]
    parts = []
    if lo < lo0:
        parts.append(spec.cells(lo, min(lo0, hi)))
    a, z = max(lo, lo0), min(hi, hi0)
    if a < z:
        p0, p1 = int(indptr[a - lo0]), int(indptr[z - lo0])
        parts.append((indptr[a - lo0:z - lo0 + 1] - p0, idx[p0:p1], val[p0:p1]))
    if hi > hi0:
        parts.append(spec.cells(max(hi0, lo), hi))
    ip, ii, vv = _concat_csc(parts)
    return lo, hi, ip, ii, vv
