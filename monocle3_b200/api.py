"""Host-side mirror of the reference's R interface for the preprocess_cds PCA path.

R is not available in this image, so the part of the path that stays in R in the
real drop-in (argument checks, set.seed(2016)+rnorm, name->index translation,
dimnames, model-slot assembly) is written here in Python with the same function
names, argument meaning and error behaviour as the reference:

  estimate_size_factors   R/utils.R:26-50
  size_factors            R/methods-cell_data_set.R:36-41
  normalized_counts       R/utils.R:477-509
  preprocess_cds          R/preprocess_cds.R:62-248   (PCA and LSI branches)
  preprocess_transform    R/projection.R:72-261       (PCA and LSI branches)
  new_cell_data_set       R/cell_data_set.R:67-137
  load_a549               R/io.R:10-26 (reads the committed fixture)

All arithmetic goes through the C ABI (engine.py -> libm3b.so); nothing here
computes on the CPU beyond O(genes) bookkeeping.
"""
import os
import warnings

import numpy as np
import scipy.sparse as sp

from . import _lib as L
from . import engine
from .r_rng import RStream, rnorm_after_seed  # noqa: F401

_default_ctx = None


def get_context(device=None):
    """Process-wide default context (one process drives one GPU)."""
    global _default_ctx
    if _default_ctx is None:
        if device is None:
            device = int(os.environ.get("LOCAL_RANK", "0"))
        _default_ctx = engine.Context(device=device, small_svd=os.environ.get("M3B_SMALL_SVD", "lapack"))
    return _default_ctx


def set_context(ctx):
    global _default_ctx
    _default_ctx = ctx


_DEFAULT_NAMES = {}


def _default_names(n):
    """Default dimnames "0".."n-1" (read-only, shared between data sets of the same size: building
    tens of thousands of Python strings per call showed up in the end-to-end time)."""
    a = _DEFAULT_NAMES.get(n)
    if a is None:
        a = np.arange(n).astype(str)
        a.setflags(write=False)
        if len(_DEFAULT_NAMES) > 8:
            _DEFAULT_NAMES.clear()
        _DEFAULT_NAMES[n] = a
    return a


class CellDataSet:
    """Minimal stand-in for the S4 `cell_data_set` (R/cell_data_set.R:28-34): the slots
    this path reads and writes.  `counts` is genes x cells CSC (the dgCMatrix).

    For a cell-sharded run `counts` holds this rank's contiguous cell range and
    `shard = (rank, world, n_cells_total, allgather)`, where allgather(np.ndarray)
    returns the concatenation over ranks in rank order (see dist.py)."""

    def __init__(self, counts, gene_names=None, cell_names=None, shard=None):
        if not isinstance(counts, engine.CSCBlocks) and not sp.isspmatrix_csc(counts):
            counts = sp.csc_matrix(counts)  # methods::as(expression_data, "dgCMatrix")
        self.counts = counts
        G, C = counts.shape
        self.gene_names = np.asarray(gene_names) if gene_names is not None else _default_names(G)
        self.cell_names = np.asarray(cell_names) if cell_names is not None else _default_names(C)
        if len(self.gene_names) != G or len(self.cell_names) != C:
            raise ValueError("dimnames do not match the expression matrix")
        self.col_data = {}           # colData(cds): 'Size_Factor', 'num_genes_expressed'
        self.row_data = {}           # rowData(cds): 'num_cells_expressed'
        self.reduced_dims = {}       # reducedDims(cds)
        self.reduce_dim_aux = {}     # cds@reduce_dim_aux
        self.shard = shard
        self._dev = None             # DeviceMatrix cache (counts resident in HBM)

    @property
    def n_cells_total(self):
        return self.shard[2] if self.shard else self.counts.shape[1]

    def device_matrix(self, ctx=None):
        ctx = ctx or get_context()
        if self._dev is None or self._dev.ctx is not ctx or self._dev.h is None:
            self._dev = engine.DeviceMatrix.from_csc(ctx, self.counts, self.n_cells_total)
        return self._dev


def size_factors(cds):
    """R/methods-cell_data_set.R:36-41."""
    return cds.col_data.get("Size_Factor")


def set_size_factors(cds, value):
    cds.col_data["Size_Factor"] = None if value is None else np.asarray(value, dtype=np.float64)
    return cds


def new_cell_data_set(expression_data, cell_names=None, gene_names=None, shard=None):
    """R/cell_data_set.R:67-137: coerce to dgCMatrix, then estimate_size_factors (:135)."""
    cds = CellDataSet(expression_data, gene_names, cell_names, shard)
    return estimate_size_factors(cds)


def load_a549():
    """R/io.R:10-26, from the committed copy of inst/extdata/small_a549_dex_*.rda."""
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "a549.npz")
    z = np.load(path, allow_pickle=False)
    dim = tuple(int(v) for v in z["dim"])
    m = sp.csc_matrix((z["x"], z["i"], z["p"]), shape=dim)
    return new_cell_data_set(m, cell_names=z["cells"], gene_names=z["genes"])


_SF_METHODS = {"mean-geometric-mean-total": L.SF_TOTAL, "mean-geometric-mean-log-total": L.SF_LOG_TOTAL}


def estimate_size_factors(cds, round_exprs=True, method="mean-geometric-mean-total"):
    """R/utils.R:26-50.  Totals come from the device (K1); the geometric mean is finished
    on the host in long double like R's mean()."""
    if method not in _SF_METHODS:
        raise ValueError("'arg' should be one of 'mean-geometric-mean-total', 'mean-geometric-mean-log-total'")
    dev = cds.device_matrix()
    raw = dev.cell_totals(round_exprs=False)
    if cds.shard:
        raw = cds.shard[3](raw)
    if np.any(raw == 0):  # R/utils.R:32-39
        warnings.warn("Your CDS object contains cells with zero reads. This causes size factor calculation to "
                      "fail. Please remove the zero read cells using cds <- cds[,Matrix::colSums(exprs(cds)) != 0] "
                      "and then run cds <- estimate_size_factors(cds)")
        return cds
    tot = dev.cell_totals(round_exprs=True) if round_exprs else None
    if tot is None:
        tot_all = raw
    else:
        tot_all = cds.shard[3](tot) if cds.shard else tot
    sf_all, _ = engine.size_factors_from_totals(tot_all, _SF_METHODS[method])
    if cds.shard:
        rank = cds.shard[0]
        counts_per_rank = cds.shard[3](np.array([cds.counts.shape[1]], dtype=np.float64)).astype(np.int64)
        off = int(counts_per_rank[:rank].sum())
        sf = sf_all[off:off + cds.counts.shape[1]]
    else:
        sf = sf_all
    return set_size_factors(cds, sf)


def normalized_counts(cds, norm_method="log", pseudocount=1):
    """R/utils.R:477-509 (sparse branch): returns a CSC with the input's pattern."""
    if norm_method not in ("log", "binary", "size_only"):
        raise ValueError("'arg' should be one of 'log', 'binary', 'size_only'")
    if isinstance(cds.counts, engine.CSCBlocks):
        raise ValueError("normalized_counts: the result would exceed a dgCMatrix; call it per block")
    dev = cds.device_matrix()
    if norm_method == "binary":
        dev.normalize(None, L.NORM_BINARY, 0.0)
    else:
        if norm_method == "log" and pseudocount != 1:
            raise ValueError("Pseudocount must equal 1 with sparse expression matrices")
        dev.normalize(size_factors(cds), L.NORM_LOG10 if norm_method == "log" else L.NORM_SIZE_ONLY,
                      1.0 if norm_method == "log" else 0.0)
    y = dev.export_values()
    out = sp.csc_matrix((y, cds.counts.indices.copy(), cds.counts.indptr.copy()), shape=cds.counts.shape)
    if norm_method == "binary":
        out.eliminate_zeros()
    return out


def jaccard_coeff(R_idx, R_weight):
    """R/RcppExports.R:4-6 -> src/clustering.cpp:49-54: Jaccard weights of a kNN graph."""
    return engine.jaccard_coeff(get_context(), R_idx, bool(R_weight))


def cluster_cells_make_graph(data, weight, cell_names=None, k=20, nn_control=None, verbose=False):
    """R/cluster_cells.R:255-336 up to the relations table, for nn_control method "nn2" (exact kNN;
    the approximate annoy / hnsw indices are out of scope): kNN of the rows of `data` (e.g.
    reducedDims(cds)[["PCA"]]) on the device, then the Jaccard weights on the device.  Returns
    dict(relations=(from, to, weight), distMatrix=...); the igraph object is left to the caller."""
    import warnings as _w
    data = np.asarray(data, dtype=np.float64)
    if data.ndim != 2:
        raise ValueError("Wrong input data, should be a data frame or matrix!")
    method = (nn_control or {}).get("method", "nn2")
    if method != "nn2":
        raise NotImplementedError("only nn_control method 'nn2' (exact search) is implemented on the device")
    if k < 1:
        raise ValueError("k must be a positive integer!")
    if k > data.shape[0] - 2:
        k = data.shape[0] - 2
        _w.warn("The nearest neighbors includes the point itself, k must be smaller than\nthe total number of points - 1 "
                "(all other points) - 1 (itself)! Total number of points is %d" % data.shape[0])
    ctx = get_context()
    idx, dist = engine.knn_self(ctx, data, k + 1)
    neighbor, distm = idx[:, 1:], dist[:, 1:]
    links = engine.jaccard_coeff(ctx, neighbor.astype(np.float64), bool(weight))
    links = links[links[:, 0] > 0]
    frm, to = links[:, 0].astype(np.int64), links[:, 1].astype(np.int64)
    if cell_names is not None:
        names = np.asarray(cell_names)
        frm, to = names[frm - 1], names[to - 1]
    return {"relations": (frm, to, links[:, 2]), "distMatrix": distm, "nn_idx": neighbor}


def detect_genes(cds, min_expr=0):
    """R/utils.R:449-457: rowData$num_cells_expressed = rowSums(counts > min_expr) (over all
    ranks' cells), colData$num_genes_expressed = colSums(counts > min_expr)."""
    if not isinstance(cds, CellDataSet):
        raise TypeError("cds is not a cell_data_set")
    if not isinstance(min_expr, (int, float, np.integer, np.floating)) or isinstance(min_expr, bool):
        raise TypeError("min_expr is not a numeric or integer vector")
    per_gene, per_cell = cds.device_matrix().detect_genes(float(min_expr))
    cds.row_data["num_cells_expressed"] = per_gene
    cds.col_data["num_genes_expressed"] = per_cell
    return cds


def _pairs(df):
    """A two-column table (sequence of (id, group) pairs, dict of two columns, 2-column array or
    pandas DataFrame) as a list of pairs."""
    if hasattr(df, "iloc"):
        return list(zip(df.iloc[:, 0].tolist(), df.iloc[:, 1].tolist()))
    if isinstance(df, dict):
        cols = list(df.values())
        return list(zip(list(cols[0]), list(cols[1])))
    return [(r[0], r[1]) for r in df]


def aggregate_gene_expression(cds, gene_group_df=None, cell_group_df=None, norm_method="log", pseudocount=1,
                              scale_agg_values=True, max_agg_value=3, min_agg_value=-3, exclude_na=True,
                              gene_agg_fun="sum", cell_agg_fun="mean"):
    """R/cluster_genes.R:447-536.  The group sums run on the device over the normalised values
    (normalized_counts, never exported); the row scaling / clamping of the small dense result is
    host work, as in the reference.  Gene ids are matched against the row names only (this
    stand-in carries no gene_short_name column).  Returns (matrix, row_names, col_names); not
    available on a cell-sharded cds without cell groups beyond the local columns."""
    if gene_group_df is None and cell_group_df is None:
        raise ValueError("one of either gene_group_df or cell_group_df must not be NULL.")
    if norm_method not in ("log", "binary", "size_only"):
        raise ValueError("'arg' should be one of 'log', 'binary', 'size_only'")
    if gene_agg_fun not in ("sum", "mean") or cell_agg_fun not in ("sum", "mean"):
        raise ValueError("aggregation functions must be 'sum' or 'mean'")
    if cds.shard and cell_group_df is None:
        # without cell groups the result has one column per cell: on a cell-sharded cds that would be
        # this rank's columns only, silently -- the caller gathers per-cell results itself
        raise ValueError("aggregate_gene_expression on a cell-sharded cds needs cell_group_df "
                         "(per-cell columns stay with their rank)")
    dev = cds.device_matrix()
    if norm_method == "binary":
        dev.normalize(None, L.NORM_BINARY, 0.0)
    else:
        if norm_method == "log" and pseudocount != 1:
            raise ValueError("Pseudocount must equal 1 with sparse expression matrices")
        dev.normalize(size_factors(cds), L.NORM_LOG10 if norm_method == "log" else L.NORM_SIZE_ONLY,
                      1.0 if norm_method == "log" else 0.0)
    G, C = cds.counts.shape
    if gene_group_df is not None:
        glut = {g: k for k, g in enumerate(cds.gene_names.tolist())}
        pairs = [(g, grp) for g, grp in _pairs(gene_group_df) if g in glut]
        groups, members = [], {}
        for g, grp in pairs:                      # unique() keeps the order of first appearance
            if grp not in members:
                members[grp] = []
                groups.append(grp)
            if g not in members[grp]:
                members[grp].append(g)
        groups = [grp for grp in groups if len(members[grp]) >= 2]  # a single row has no dim(): NA, dropped (:490-491)
        # The device pass takes ONE group id per gene.  A gene listed under several groups counts
        # in each of them in the reference, so the groups are dealt into layers inside which no
        # gene repeats (almost always a single layer) and every layer is one pass.
        layers = []  # [(gene_group array, [group positions])]
        for q, grp in enumerate(groups):
            idx = [glut[g] for g in members[grp]]
            for arr, pos in layers:
                if np.all(arr[idx] < 0):
                    arr[idx] = len(pos)
                    pos.append(q)
                    break
            else:
                arr = np.full(G, -1, dtype=np.int32)
                arr[idx] = 0
                layers.append((arr, [q]))
        row_names = list(groups)
    else:
        layers = []  # every gene is its own row: passes of at most 1024 rows (the device limit)
        for g0 in range(0, G, 1024):
            g1 = min(G, g0 + 1024)
            arr = np.full(G, -1, dtype=np.int32)
            arr[g0:g1] = np.arange(g1 - g0, dtype=np.int32)
            layers.append((arr, list(range(g0, g1))))
        groups = list(range(G))
        row_names = cds.gene_names.tolist()
    if len(groups) == 0:
        return np.zeros((0, 0)), [], []
    cell_group, levels = None, []
    if cell_group_df is not None:
        clut = {c: k for k, c in enumerate(cds.cell_names.tolist())}
        pairs = [(c, grp) for c, grp in _pairs(cell_group_df) if c in clut]
        levels = sorted(set(grp for _, grp in pairs))  # as.factor()
        lv = {l: q for q, l in enumerate(levels)}
        cell_group = np.full(C, -1, dtype=np.int32)
        for c, grp in pairs:
            cell_group[clut[c]] = lv[grp]
        col_names = list(levels)
    else:
        col_names = cds.cell_names.tolist()
    agg = np.zeros((len(groups), len(col_names)))
    for arr, pos in layers:
        part = dev.aggregate(arr, len(pos), gene_agg_fun == "mean", cell_group, len(levels), cell_agg_fun == "mean")
        agg[pos, :] = part
    if scale_agg_values:
        with np.errstate(invalid="ignore", divide="ignore"):
            mu = agg.mean(axis=1, keepdims=True)
            sd = agg.std(axis=1, ddof=1, keepdims=True) if agg.shape[1] > 1 else np.full((agg.shape[0], 1), np.nan)
            agg = (agg - mu) / sd
        agg[np.isnan(agg)] = 0
        agg = np.clip(agg, min_agg_value, max_agg_value)
    if exclude_na:
        rk = [k for k, n in enumerate(row_names) if n != "NA"]
        ck = [k for k, n in enumerate(col_names) if n != "NA"]
        agg = agg[np.ix_(rk, ck)]
        row_names = [row_names[k] for k in rk]
        col_names = [col_names[k] for k in ck]
    return agg, row_names, col_names


_NORM = {"log": L.NORM_LOG2, "size_only": L.NORM_SIZE_ONLY, "none": L.NORM_NONE}


def _normalize_on_device(cds, norm_method, pseudo_count):
    """normalize_expr_data (R/preprocess_cds.R:253-286) -- result stays on the device."""
    if pseudo_count is None:
        pseudo_count = 1 if norm_method == "log" else 0
    dev = cds.device_matrix()
    dev.normalize(size_factors(cds) if norm_method != "none" else None, _NORM[norm_method],
                  float(pseudo_count) if norm_method != "none" else 0.0)
    return dev


def preprocess_cds(cds, method="PCA", num_dim=50, norm_method="log", use_genes=None, pseudo_count=None,
                   scaling=True, verbose=False, build_nn_index=False, nn_control=None, _irlba_args=None):
    """R/preprocess_cds.R:62-248 (PCA and LSI branches)."""
    if method not in ("PCA", "LSI"):
        raise ValueError("method must be one of 'PCA' or 'LSI'")
    if norm_method not in _NORM:
        raise ValueError("norm_method must be one of 'log', 'size_only' or 'none'")
    if not (isinstance(num_dim, (int, np.integer)) and num_dim > 0):
        raise ValueError("num_dim is not a count (a single positive integer)")
    if build_nn_index:
        raise NotImplementedError("nearest-neighbour indices are outside this build's scope (SURVEY.md section 8f)")
    gene_order = None
    if use_genes is not None:
        use_genes = list(use_genes)
        lut = {g: k for k, g in enumerate(cds.gene_names)}
        if not all(isinstance(g, str) for g in use_genes) or not all(g in lut for g in use_genes):
            raise ValueError("use_genes must be NULL, or all must be present in the row.names of rowData(cds)")
        gene_order = np.array([lut[g] for g in use_genes], dtype=np.int32)
    if size_factors(cds) is None:
        raise ValueError("You must call estimate_size_factors before calling preprocess_cds.")
    if np.any(np.isnan(size_factors(cds))):
        raise ValueError("One or more cells has a size factor of NA.")

    dev = _normalize_on_device(cds, norm_method, pseudo_count)  # :111
    if cds.counts.shape[0] == 0:
        raise ValueError("all rows have standard deviation zero")
    keep = dev.select_genes(gene_order)  # :117-122 + Matrix::t
    sel = np.arange(cds.counts.shape[0]) if gene_order is None else gene_order
    kept_idx = sel[keep]
    if method == "LSI":
        return _preprocess_lsi(cds, dev, kept_idx, num_dim, norm_method, use_genes, pseudo_count, _irlba_args)
    if verbose:
        print("Remove noise by PCA ...")
    n = min(num_dim, min(len(kept_idx), cds.n_cells_total) - 1)  # :140
    # set.seed(2016) (:110); irlba draws rnorm(ncol(A)) for its start vector, and further vectors
    # from the same stream if the Lanczos process hits an invariant subspace
    stream = RStream(2016)
    v0 = stream.rnorm(len(kept_idx))
    # (cell-sharded runs: every rank continues the same stream; the library slices cell-length draws)
    dev.ctx.set_rnorm(stream.rnorm)
    try:
        res = dev.pca_irlba(n, v0, center=scaling, scale=scaling, **(_irlba_args or {}))
    finally:
        dev.ctx.set_rnorm(None)
    if res["status"] == L.M3B_W_NOT_CONVERGED:
        warnings.warn("did not converge--results might be invalid!; try increasing work or maxit")
    C_tot = cds.n_cells_total
    sdev = res["d"] / np.sqrt(max(1, C_tot - 1))  # R/utils.R:418
    cds.reduced_dims["PCA"] = res["x"]            # rows = colnames(cds), cols = PC1..
    cds.reduced_dims_dimnames = {"PCA": (cds.cell_names, ["PC%d" % (k + 1) for k in range(n)])}
    cds.reduce_dim_aux["PCA"] = {"model": {
        "num_dim": num_dim, "norm_method": norm_method, "use_genes": use_genes, "pseudo_count": pseudo_count,
        "svd_v": res["v"], "svd_v_rownames": cds.gene_names[kept_idx],
        "svd_sdev": sdev, "svd_center": res["center"], "svd_scale": res["scale"],
        "prop_var_expl": sdev ** 2 / np.sum(sdev ** 2),  # :158-160
    }, "irlba": {"iter": res["iter"], "mprod": res["mprod"], "d": res["d"]}}
    return cds


def _preprocess_lsi(cds, dev, kept_idx, num_dim, norm_method, use_genes, pseudo_count, irlba_args):
    """LSI branch, R/preprocess_cds.R:185-241: tfidf(FM) then irlba without centre / scale."""
    col_sums, row_sums = dev.tfidf(True, True, 100000.0)        # tfidf() defaults, :289-290
    n = min(num_dim, min(len(kept_idx), cds.n_cells_total) - 1)  # :194
    v0 = rnorm_after_seed(2016, len(kept_idx))
    res = dev.pca_irlba(n, v0, center=False, scale=False, **(irlba_args or {}))
    if res["status"] == L.M3B_W_NOT_CONVERGED:
        warnings.warn("did not converge--results might be invalid!; try increasing work or maxit")
    num_col = cds.n_cells_total
    cds.reduced_dims["LSI"] = res["x"]                          # irlba_res$u %*% diag(irlba_res$d)  :196
    cds.reduce_dim_aux["LSI"] = {"model": {
        "num_dim": num_dim, "norm_method": norm_method, "use_genes": use_genes, "pseudo_count": pseudo_count,
        "log_scale_tf": True, "frequencies": True, "scale_factor": 100000.0,
        "col_sums": col_sums, "row_sums": row_sums, "num_cols": num_col,
        "svd_v": res["v"], "svd_v_rownames": cds.gene_names[kept_idx],
        "svd_sdev": res["d"] / np.sqrt(max(1, num_col - 1)),    # :213
    }, "irlba": {"iter": res["iter"], "mprod": res["mprod"], "d": res["d"]}}
    return cds


def preprocess_transform(cds, reduction_method="PCA", block_size=None, cores=1):
    """R/projection.R:72-155, PCA branch.  block_size / cores are accepted for signature
    compatibility; the projection is a single sparse product on the device (K10)."""
    if reduction_method not in ("PCA", "LSI"):
        raise ValueError("reduction_method must be 'PCA' or 'LSI'")
    if size_factors(cds) is None:
        raise ValueError("You must call estimate_size_factors before calling preprocess_cds.")
    if np.any(np.isnan(size_factors(cds))):
        raise ValueError("One or more cells has a size factor of NA.")
    if cds.reduce_dim_aux.get(reduction_method) is None:
        raise ValueError("Reduction method '%s' is not in the model object." % reduction_method)
    model = cds.reduce_dim_aux[reduction_method]["model"]
    dev = _normalize_on_device(cds, model["norm_method"], model["pseudo_count"])  # :109
    if cds.counts.shape[0] == 0:
        raise ValueError("all rows have standard deviation zero")
    light = reduction_method == "PCA"  # the LSI branch rewrites the product layouts (TF-IDF): it needs them
    keep = dev.select_genes(None, for_transform=light)  # :117-118
    kept_names = cds.gene_names[keep]
    # intersect(rownames(rotation), rownames(FM)) is in the MODEL's order (:121-122); every kept
    # query gene must be in the model, otherwise FM[intersect_genes,] would silently drop rows --
    # R stops when match() yields NA.
    model_names = np.asarray(model["svd_v_rownames"])
    if len(model_names) == len(kept_names) and np.array_equal(model_names, kept_names):
        model_row = np.arange(len(kept_names), dtype=np.int32)  # the usual case: same gene set, same order
    else:  # match(kept_names, model_names) with sorted look-ups (no Python-level loop over 30k names)
        order = np.argsort(model_names, kind="stable")
        pos = np.searchsorted(model_names, kept_names, sorter=order)
        pos = np.clip(pos, 0, max(len(model_names) - 1, 0))
        cand = order[pos] if len(model_names) else np.zeros(len(kept_names), dtype=np.int64)
        hit = (model_names[cand] == kept_names) if len(model_names) else np.zeros(len(kept_names), dtype=bool)
        model_row = np.where(hit, cand, -1).astype(np.int32)
    n_common = int(np.sum(model_row >= 0))
    if len(kept_names) and n_common / len(kept_names) < 0.5:  # :128-130
        warnings.warn("fewer than half the genes in the subject matrix are also in the reference matrix: "
                      "are the matrices prepared using the same gene set?")
    if np.any(model_row < 0):
        # genes of the query that the model does not know are dropped by the intersect (:121,133)
        qidx = np.nonzero(keep)[0][model_row >= 0]
        keep2 = dev.select_genes(qidx.astype(np.int32), for_transform=light)
        if not np.all(keep2):
            raise RuntimeError("internal: filter changed between passes")
        model_row = model_row[model_row >= 0]
    if reduction_method == "LSI":
        # R/projection.R:190-216: the query's own column sums, the MODEL's row_sums / num_cols
        dev.tfidf(model["frequencies"], model["log_scale_tf"], model["scale_factor"],
                  row_sums=np.asarray(model["row_sums"])[model_row], num_cols=model["num_cols"])
        Y = dev.project(model_row, model["svd_v"], None, None)
    else:
        Y = dev.project(model_row, model["svd_v"], model["svd_center"], model["svd_scale"])
    cds.reduced_dims[reduction_method] = Y
    return cds


# ------------------------------------------------------------------------------------------------
# align_cds (residual regression part) and the transform-model hand-off
# ------------------------------------------------------------------------------------------------
def _parse_formula(formula):
    """'~ a + b' -> ['a', 'b'] (main effects only: what align_cds is used with in the reference's
    tests and vignette; interactions / functions are left to R's model.matrix)."""
    f = formula.strip()
    if not f.startswith("~"):
        raise ValueError("residual_model_formula_str must be a one-sided formula such as '~ batch'")
    terms = [t.strip() for t in f[1:].split("+") if t.strip() not in ("", "1")]
    for t in terms:
        if any(ch in t for ch in "*:()^|/ -"):
            raise NotImplementedError("only main effects are supported in the formula (term '%s')" % t)
    return terms


def model_matrix(cds, formula):
    """Matrix::sparse.model.matrix(as.formula(formula), colData(cds), drop.unused.levels = TRUE)
    for main-effect formulas: intercept, numeric columns as they are, factors with treatment
    contrasts (levels sorted like as.factor, first level dropped).  Returns (M, column names).
    On a cell-sharded cds the factor levels are agreed on across the ranks."""
    terms = _parse_formula(formula)
    n = cds.counts.shape[1]
    cols, names = [np.ones(n)], ["(Intercept)"]
    for t in terms:
        if t not in cds.col_data or cds.col_data[t] is None:
            raise ValueError("object '%s' not found in colData(cds)" % t)
        v = np.asarray(cds.col_data[t])
        if v.dtype.kind in "fiub":
            cols.append(v.astype(np.float64))
            names.append(t)
            continue
        levels = sorted(set(v.tolist()))
        if cds.shard:  # every rank must build the same columns
            import torch.distributed as td
            allv = [None] * cds.shard[1]
            td.all_gather_object(allv, levels)
            levels = sorted(set(x for l in allv for x in l))
        for lv in levels[1:]:
            cols.append((v == lv).astype(np.float64))
            names.append("%s%s" % (t, lv))
    return np.stack(cols, axis=1), names


def align_cds(cds, preprocess_method="PCA", alignment_group=None, alignment_k=20, residual_model_formula_str=None,
              verbose=False, build_nn_index=False, nn_control=None):
    """R/alignment.R:66-160.  The residual regression (limma::lmFit on the PCA / LSI coordinates,
    :106-117) runs on the device (m3b_align_residuals); the mutual-nearest-neighbour correction
    (alignment_group, batchelor::reducedMNN) is outside this build's scope."""
    if preprocess_method not in ("PCA", "LSI"):
        raise ValueError("preprocess_method must be one of 'PCA' or 'LSI'")
    X = cds.reduced_dims.get(preprocess_method)
    if X is None:
        raise ValueError("Preprocessing for '%s' does not exist. Please make sure you have run preprocess_cds with"
                         "preprocess_method = '%s' before calling align_cds." % (preprocess_method, preprocess_method))
    if build_nn_index:
        raise NotImplementedError("nearest-neighbour indices are outside this build's scope (SURVEY.md section 8f)")
    if alignment_group is not None:
        raise NotImplementedError("alignment_group (batchelor::reducedMNN) is outside this build's scope; use "
                                  "residual_model_formula_str for linear effects")
    model = {"alignment_group": alignment_group, "alignment_k": alignment_k,
             "residual_model_formula_str": residual_model_formula_str, "beta": None}
    if residual_model_formula_str is not None:
        if verbose:
            print("Removing residual effects")
        M, names = model_matrix(cds, residual_model_formula_str)
        X, beta = engine.align_residuals(get_context(), X, M)
        model["beta"] = beta
        model["beta_colnames"] = names[1:]
    cds.reduced_dims["Aligned"] = np.asarray(X)
    cds.reduce_dim_aux["Aligned"] = {"model": model}
    return cds


def align_beta_transform(cds, preprocess_method="PCA"):
    """R/projection.R:468-494: subtract the effects of the SAVED beta from the projected coordinates."""
    X = cds.reduced_dims.get(preprocess_method)
    model = (cds.reduce_dim_aux.get("Aligned") or {}).get("model")
    if X is None or model is None or model.get("beta") is None:
        raise ValueError("align_beta_transform needs reducedDims '%s' and an 'Aligned' model with beta" % preprocess_method)
    M, _ = model_matrix(cds, model["residual_model_formula_str"])
    if M.shape[1] - 1 != np.asarray(model["beta"]).shape[1]:
        raise ValueError("the design has %d columns, the saved beta %d" % (M.shape[1] - 1, np.asarray(model["beta"]).shape[1]))
    Xa, _ = engine.align_residuals(get_context(), X, M, beta=model["beta"])
    cds.reduced_dims["Aligned"] = Xa
    return cds


_MODEL_ARRAYS = ("svd_v", "svd_v_rownames", "svd_sdev", "svd_center", "svd_scale", "prop_var_expl", "col_sums", "row_sums",
                 "beta")


def save_transform_models(cds, directory_path):
    """The model hand-off of R/io.R:819-996 for the reductions this build produces (PCA, LSI,
    Aligned): everything preprocess_transform / align_beta_transform read, one .npz per method plus
    a small index.  (The reference writes RDS files and the nn indices; RDS stays in R.)"""
    import json
    os.makedirs(directory_path, exist_ok=True)
    index = {}
    for method, aux in cds.reduce_dim_aux.items():
        model = (aux or {}).get("model")
        if model is None:
            continue
        arrays = {k: np.asarray(v) for k, v in model.items() if k in _MODEL_ARRAYS and v is not None}
        scalars = {k: v for k, v in model.items() if k not in _MODEL_ARRAYS}
        np.savez(os.path.join(directory_path, "rdd_%s_transform_model.npz" % method.lower()), **arrays)
        index[method] = {"scalars": scalars, "arrays": sorted(arrays)}
    with open(os.path.join(directory_path, "file_index.json"), "w") as f:
        json.dump(index, f, default=lambda o: o.tolist() if hasattr(o, "tolist") else str(o))
    return cds


def load_transform_models(cds, directory_path):
    """R/io.R:1025-1146: put the saved models back into cds@reduce_dim_aux."""
    import json
    with open(os.path.join(directory_path, "file_index.json")) as f:
        index = json.load(f)
    for method, rec in index.items():
        z = np.load(os.path.join(directory_path, "rdd_%s_transform_model.npz" % method.lower()), allow_pickle=False)
        model = dict(rec["scalars"])
        for k in rec["arrays"]:
            model[k] = z[k]
        for k in _MODEL_ARRAYS:
            model.setdefault(k, None)
        cds.reduce_dim_aux[method] = {"model": model}
    return cds
