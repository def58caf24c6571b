"""Thin object wrappers over the C ABI (include/m3b.h): Context and DeviceMatrix.

No arithmetic happens here -- only marshaling of NumPy buffers to the extern "C"
entry points, exactly what the Rcpp layer does in the R package.
"""
import ctypes as C

import numpy as np

from . import _lib as L


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double)) if a is not None else None


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int32)) if a is not None else None


def lapack_small_svd():
    """small_svd callback backed by NumPy's LAPACK gesdd -- the stand-in for R's
    own LAPACK that the R shim passes for sign fidelity with stock monocle3."""
    def cb(n, Bp, Up, sp, Vtp, _user):
        try:
            B = np.ctypeslib.as_array(Bp, shape=(n * n,)).reshape(n, n).T  # column-major -> B
            U, s, Vt = np.linalg.svd(B)
            np.ctypeslib.as_array(Up, shape=(n * n,))[:] = np.asfortranarray(U).ravel(order="F")
            np.ctypeslib.as_array(sp, shape=(n,))[:] = s
            np.ctypeslib.as_array(Vtp, shape=(n * n,))[:] = np.asfortranarray(Vt).ravel(order="F")
            return 0
        except Exception:  # never let an exception cross the C boundary
            return 1
    return L.SMALL_SVD_FN(cb)


class Context:
    """One GPU (m3b_ctx). `small_svd`: 'device' (Jacobi kernel) or 'lapack' (host callback)."""

    def __init__(self, device=0, flags=0, small_svd="device"):
        self.lib = L.load()
        h = C.c_void_p()
        st = self.lib.m3b_create(int(device), int(flags), C.byref(h))
        if st != L.M3B_OK:
            raise L.M3BError(st, self.lib.m3b_status_string(st).decode())
        self.h = h
        self.device = device
        self._cb = None
        self.rank, self.world = 0, 1
        self.set_small_svd(small_svd)

    def set_rnorm(self, fn):
        """fn(n) -> n standard normals from the caller's stream (used only when the Lanczos process
        hits an invariant subspace); None unregisters."""
        if fn is None:
            self._rnorm_cb = None
            self.check(self.lib.m3b_set_rnorm(self.h, C.cast(None, L.RNORM_FN), None))
            return

        def cb(n, out, _user):
            try:
                v = np.asarray(fn(int(n)), dtype=np.float64)
                C.memmove(out, v.ctypes.data, int(n) * 8)
                return 0
            except Exception:
                return 1
        self._rnorm_cb = L.RNORM_FN(cb)
        self.check(self.lib.m3b_set_rnorm(self.h, self._rnorm_cb, None))

    def set_small_svd(self, mode):
        if mode == "lapack":
            self._cb = lapack_small_svd()
            self.check(self.lib.m3b_set_small_svd(self.h, self._cb, None))
        elif mode == "device":
            self._cb = None
            self.check(self.lib.m3b_set_small_svd(self.h, C.cast(None, L.SMALL_SVD_FN), None))
        else:
            raise ValueError("small_svd must be 'device' or 'lapack'")
        self.small_svd = mode

    def check(self, st, allow=()):
        if st != L.M3B_OK and st not in allow:
            raise L.M3BError(st, self.lib.m3b_last_error(self.h).decode() or self.lib.m3b_status_string(st).decode())
        return st

    def comm_unique_id(self):
        buf = C.create_string_buffer(L.COMM_ID_BYTES)
        self.check(self.lib.m3b_comm_unique_id(self.h, buf))
        return buf.raw

    def comm_init(self, rank, world, uid):
        self.check(self.lib.m3b_comm_init(self.h, int(rank), int(world), uid))
        self.rank, self.world = rank, world

    def comm_p2p_handle(self):
        buf = C.create_string_buffer(64)
        self.check(self.lib.m3b_comm_p2p_handle(self.h, buf))
        return buf.raw

    def comm_p2p_open(self, handles):
        self.check(self.lib.m3b_comm_p2p_open(self.h, b"".join(handles)))

    def stats(self):
        s = L.Stats()
        self.check(self.lib.m3b_get_stats(self.h, C.byref(s)))
        out = {k: getattr(s, k) for k, _ in L.Stats._fields_}
        out["phase_ms"] = list(out["phase_ms"])
        out["host_ms"] = list(out["host_ms"])
        out["select_ms"] = list(out["select_ms"])
        out["xchg_diag"] = list(out["xchg_diag"])
        return out

    def measure_fp64_peak(self):
        """Measured fp64 FMA peak of this GPU in TFLOP/s (the roof of the projection kernel)."""
        v = C.c_double()
        self.check(self.lib.m3b_measure_fp64_peak(self.h, C.byref(v)))
        return v.value

    def set_flags(self, flags):
        self.check(self.lib.m3b_set_flags(self.h, int(flags)))

    def timer_start(self):
        self.check(self.lib.m3b_timer_start(self.h))

    def timer_stop(self):
        ms = C.c_double()
        self.check(self.lib.m3b_timer_stop(self.h, C.byref(ms)))
        return ms.value

    def reset_stats(self):
        self.check(self.lib.m3b_reset_stats(self.h))

    def close(self):
        if getattr(self, "h", None):
            self.lib.m3b_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def knn_self(ctx, X, k):
    """Exact kNN of the rows of X among themselves: (nn_idx n x k, 1-based, nn_dists n x k), the
    result layout of RANN::nn2(X, X, k)."""
    a = np.asarray(X, dtype=np.float64)
    if not a.flags.f_contiguous:
        a = np.asfortranarray(a)
    n, d = a.shape
    idx = np.empty((k, n), dtype=np.int32)   # column-major n x k
    dist = np.empty((k, n))
    ctx.check(ctx.lib.m3b_knn_self(ctx.h, a.ctypes.data_as(C.POINTER(C.c_double)), int(n), int(d), int(k),
                                   idx.ctypes.data_as(C.POINTER(C.c_int32)), _dp(dist)))
    return idx.T, dist.T


def jaccard_coeff(ctx, idx, weight=True, out=None):
    """(n*k) x 3 matrix (from, to, weight) of the kNN graph `idx` (n x k, 1-based ids).  `out`:
    optional (3, n*k) C-ordered buffer (= the column-major result); then the transposed VIEW of it
    is returned and nothing is allocated or copied on the host."""
    a = np.asarray(idx, dtype=np.float64)
    if not a.flags.f_contiguous:
        a = np.asfortranarray(a)  # an R NumericMatrix: doubles, column-major
    n, k = a.shape
    res = np.empty((3, n * k)) if out is None else out
    ctx.check(ctx.lib.m3b_jaccard_coeff(ctx.h, a.ctypes.data_as(C.POINTER(C.c_double)), int(n), int(k),
                                        int(bool(weight)), _dp(res)))
    return res.T.copy() if out is None else res.T


def align_residuals(ctx, X, M, beta=None):
    """align_cds residual regression on the device: X (cells x nv) and the design M (cells x p,
    intercept first) -> (X_aligned, beta nv x (p-1)).  With `beta` given, only subtracts its effects
    (align_beta_transform)."""
    Xf = np.array(X, dtype=np.float64, order="F", copy=True)
    Mf = np.asfortranarray(M, dtype=np.float64)
    n, nv = Xf.shape
    p = Mf.shape[1]
    if beta is None:
        b = np.zeros((nv, max(p - 1, 0)), order="F")
        bb = b if b.size else np.zeros(1)
        ctx.check(ctx.lib.m3b_align_residuals(ctx.h, _dp(Xf), n, nv, max(n, 1), _dp(Mf), p, _dp(bb)))
    else:
        b = np.asfortranarray(beta, dtype=np.float64)
        bb = b if b.size else np.zeros(1)
        ctx.check(ctx.lib.m3b_align_apply_beta(ctx.h, _dp(Xf), n, nv, max(n, 1), _dp(Mf), p, _dp(bb)))
    return Xf, b


class CSCBlocks:
    """A genes x cells count matrix given as consecutive CELL BLOCKS, each a scipy csc_matrix (a
    dgCMatrix: < 2^31 stored entries).  This is how a matrix beyond the dgCMatrix limit reaches the
    library (the 2M-cell target has 3.2e9 entries): the R shim appends the
    blocks in order through m3b_matrix_append_csc, exactly like the chunks of one dgCMatrix."""

    def __init__(self, blocks):
        self.blocks = list(blocks)
        if not self.blocks:
            raise ValueError("CSCBlocks needs at least one block")
        G = self.blocks[0].shape[0]
        if any(b.shape[0] != G for b in self.blocks):
            raise ValueError("all blocks must have the same number of genes")
        self.shape = (G, int(sum(b.shape[1] for b in self.blocks)))
        self.nnz = int(sum(int(b.nnz) for b in self.blocks))


class DeviceMatrix:
    """The counts matrix (local cell shard) resident on the GPU (m3b_mat)."""

    def __init__(self, ctx, n_genes, n_cells_local, n_cells_total=None, nnz_hint=0):
        self.ctx = ctx
        self.lib = ctx.lib
        self.n_genes = int(n_genes)
        self.n_cells = int(n_cells_local)
        self.n_cells_total = int(n_cells_total if n_cells_total is not None else n_cells_local)
        h = C.c_void_p()
        ctx.check(self.lib.m3b_matrix_begin(ctx.h, self.n_genes, self.n_cells, self.n_cells_total, int(nnz_hint), C.byref(h)))
        self.h = h
        self.nnz = 0
        self.n_sel = self.n_kept = None

    @classmethod
    def from_csc(cls, ctx, counts, n_cells_total=None, chunk_cells=None):
        """counts: scipy.sparse.csc_matrix genes x (local) cells.  Appended in cell-range
        chunks whose nnz fits int32 (the dgCMatrix limit)."""
        G, Cn = counts.shape
        if isinstance(counts, CSCBlocks):
            m = cls(ctx, G, Cn, n_cells_total, counts.nnz)
            for blk in counts.blocks:
                if not blk.has_canonical_format:
                    blk = blk.copy()
                    blk.sum_duplicates()
                cls._append_ranges(m, blk, chunk_cells)
            m.end()
            return m
        if not counts.has_canonical_format:  # a dgCMatrix is sorted and duplicate-free; scipy need not be
            counts = counts.copy()
            counts.sum_duplicates()
        m = cls(ctx, G, Cn, n_cells_total, counts.nnz)
        cls._append_ranges(m, counts, chunk_cells)
        m.end()
        return m

    @staticmethod
    def _append_ranges(m, counts, chunk_cells=None):
        """Append one CSC block in cell ranges whose nnz fits int32."""
        Cn = counts.shape[1]
        indptr = np.asarray(counts.indptr)
        indices = np.ascontiguousarray(counts.indices, dtype=np.int32)
        data = np.ascontiguousarray(counts.data, dtype=np.float64)
        lo = 0
        limit = 2 ** 31 - 1
        while lo < Cn:
            hi = Cn if chunk_cells is None else min(Cn, lo + chunk_cells)
            while int(indptr[hi]) - int(indptr[lo]) > limit:
                hi = lo + max(1, (hi - lo) // 2)
            m.append_csc(indptr[lo:hi + 1] - indptr[lo], indices[indptr[lo]:indptr[hi]], data[indptr[lo]:indptr[hi]])
            lo = hi

    def append_csc(self, p, i, x):
        p = np.ascontiguousarray(p, dtype=np.int32)
        i = np.ascontiguousarray(i, dtype=np.int32)
        x = np.ascontiguousarray(x, dtype=np.float64)
        self.ctx.check(self.lib.m3b_matrix_append_csc(self.h, len(p) - 1, _ip(p), _ip(i), _dp(x)))
        self.nnz += len(x)

    def end(self):
        self.ctx.check(self.lib.m3b_matrix_end(self.h))

    def cell_totals(self, round_exprs=True):
        out = np.empty(max(self.n_cells, 1), dtype=np.float64)
        self.ctx.check(self.lib.m3b_cell_totals(self.h, int(bool(round_exprs)), _dp(out)))
        return out[:self.n_cells]

    def normalize(self, sf, norm_method, pseudo_count):
        sf_a = None if sf is None else np.ascontiguousarray(sf, dtype=np.float64)
        self.ctx.check(self.lib.m3b_normalize(self.h, _dp(sf_a), int(norm_method), float(pseudo_count)))

    def export_values(self):
        y = np.empty(max(self.nnz, 1), dtype=np.float64)
        self.ctx.check(self.lib.m3b_export_values(self.h, _dp(y)))
        return y[:self.nnz]

    def select_genes(self, gene_order=None, for_transform=False):
        """for_transform: the light variant (m3b_select_genes_for_transform), only project() may follow."""
        if gene_order is None:
            n_sel, go = self.n_genes, None
        else:
            go = np.ascontiguousarray(gene_order, dtype=np.int32)
            n_sel = len(go)
        keep = np.zeros(max(n_sel, 1), dtype=np.uint8)
        nk = C.c_int32()
        fn = self.lib.m3b_select_genes_for_transform if for_transform else self.lib.m3b_select_genes
        self.ctx.check(fn(self.h, _ip(go), n_sel, keep.ctypes.data_as(C.POINTER(C.c_uint8)), C.byref(nk)))
        self.n_sel, self.n_kept = n_sel, nk.value
        return keep[:n_sel].astype(bool)

    def gene_nnz(self):
        out = np.zeros(max(self.n_sel, 1), dtype=np.int64)
        self.ctx.check(self.lib.m3b_gene_nnz(self.h, out.ctypes.data_as(C.POINTER(C.c_int64))))
        return out[:self.n_sel]

    def detect_genes(self, min_expr=0.0):
        """(num_cells_expressed[G] over all ranks, num_genes_expressed[local cells])."""
        pg = np.zeros(max(self.n_genes, 1), dtype=np.int64)
        pc = np.zeros(max(self.n_cells, 1), dtype=np.int64)
        self.ctx.check(self.lib.m3b_detect_genes(self.h, float(min_expr), pg.ctypes.data_as(C.POINTER(C.c_int64)),
                                                 pc.ctypes.data_as(C.POINTER(C.c_int64))))
        return pg[:self.n_genes], pc[:self.n_cells]

    def aggregate(self, gene_group, n_gene_groups, gene_mean=False, cell_group=None, n_cell_groups=0, cell_mean=True):
        """Group sums of the normalised values: (n_gene_groups x n_cell_groups) or (n_gene_groups x local cells)."""
        gg = np.ascontiguousarray(gene_group, dtype=np.int32)
        ncol = n_cell_groups if cell_group is not None else self.n_cells
        out = np.zeros((max(ncol, 1), max(n_gene_groups, 1)))  # column-major n_gene_groups x ncol
        cg = None if cell_group is None else np.ascontiguousarray(cell_group, dtype=np.int32)
        ip = C.POINTER(C.c_int32)
        self.ctx.check(self.lib.m3b_aggregate_expression(
            self.h, gg.ctypes.data_as(ip), int(n_gene_groups), int(bool(gene_mean)),
            None if cg is None else cg.ctypes.data_as(ip), int(n_cell_groups), int(bool(cell_mean)), _dp(out),
            int(max(n_gene_groups, 1))))
        return out[:ncol, :n_gene_groups].T.copy()

    def gene_moments(self):
        c = np.empty(max(self.n_kept, 1))
        s = np.empty(max(self.n_kept, 1))
        self.ctx.check(self.lib.m3b_gene_moments(self.h, _dp(c), _dp(s)))
        return c[:self.n_kept], s[:self.n_kept]

    def tfidf(self, frequencies=True, log_scale_tf=True, scale_factor=100000.0, row_sums=None, num_cols=0):
        """In-place TF-IDF of the prepared matrix (LSI branch). Returns (col_sums local, row_sums)."""
        cs = np.empty(max(self.n_cells, 1))
        rs = np.empty(max(self.n_kept, 1))
        rin = None if row_sums is None else np.ascontiguousarray(row_sums, dtype=np.float64)
        self.ctx.check(self.lib.m3b_tfidf(self.h, int(frequencies), int(log_scale_tf), float(scale_factor), _dp(rin),
                                          float(num_cols), _dp(cs), _dp(rs)))
        return cs[:self.n_cells], rs[:self.n_kept]

    def pca_irlba(self, nv, v0, center=True, scale=True, work=0, maxit=1000, tol=1e-5, svtol=None, want_x=True):
        """Returns dict(d, v, x, center, scale, iter, mprod, status)."""
        n = self.n_kept
        v0 = np.ascontiguousarray(v0, dtype=np.float64)
        if len(v0) != n:
            raise ValueError("start vector length %d != kept genes %d" % (len(v0), n))
        d = np.empty(nv)
        it, mp = C.c_int(), C.c_int()
        svt = tol if svtol is None else svtol
        if want_x:
            V = np.empty((n, nv), order="F")
            X = np.empty((self.n_cells, nv), order="F")
            cen = np.empty(n) if center else None
            sc = np.empty(n) if scale else None
            st = self.lib.m3b_pca_irlba(self.h, nv, work, maxit, tol, svt, _dp(v0), int(center), int(scale), _dp(d), _dp(V),
                                        _dp(X), max(self.n_cells, 1), _dp(cen), _dp(sc), C.byref(it), C.byref(mp))
        else:
            V = X = cen = sc = None
            st = self.lib.m3b_pca_irlba_device(self.h, nv, work, maxit, tol, svt, _dp(v0), int(center), int(scale), _dp(d),
                                               C.byref(it), C.byref(mp))
        self.ctx.check(st, allow=(L.M3B_W_NOT_CONVERGED,))
        return dict(d=d, v=V, x=X, center=cen, scale=sc, iter=it.value, mprod=mp.value, status=st)

    def project(self, model_row, V, center, scale):
        V = np.asfortranarray(V, dtype=np.float64)
        n_model, nv = V.shape
        mr = np.ascontiguousarray(model_row, dtype=np.int32)
        Y = np.empty((self.n_cells, nv), order="F")
        c = None if center is None else np.ascontiguousarray(center, dtype=np.float64)
        s = None if scale is None else np.ascontiguousarray(scale, dtype=np.float64)
        self.ctx.check(self.lib.m3b_project(self.h, _ip(mr), n_model, _dp(V), _dp(c), _dp(s), nv, _dp(Y), max(self.n_cells, 1)))
        return Y

    def free(self):
        if getattr(self, "h", None):
            self.lib.m3b_matrix_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def size_factors_from_totals(cell_total, method=L.SF_TOTAL):
    """Host-side finish of estimate_sf_sparse (R/utils.R:61-68) through the ABI."""
    lib = L.load()
    tot = np.ascontiguousarray(cell_total, dtype=np.float64)
    sf = np.empty(max(len(tot), 1))
    az = C.c_int()
    st = lib.m3b_size_factors_from_totals(_dp(tot), len(tot), int(method), _dp(sf), C.byref(az))
    if st != L.M3B_OK:
        raise L.M3BError(st, lib.m3b_status_string(st).decode())
    return sf[:len(tot)], bool(az.value)
