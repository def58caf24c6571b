"""CPU restatement (NumPy/SciPy) of monocle3's preprocess_cds PCA hot path.

TEST INFRASTRUCTURE ONLY (oracle/): the parity checker for tests/ and the CPU
baseline for bench.py.  The product package (monocle3_b200/) never imports it.

Each function cites the reference lines it follows (paths under /root/reference).
Third-party arithmetic (irlba, Matrix, DelayedMatrixStats, base-R RNG) is
restated in irlba_ref.py / r_rng.py.  Parity is pinned by the reference's own
golden values -- tests/test_oracle_golden.py.

detect_genes / aggregate_gene_expression are PARITY UNPINNED: the
reference's tests hold no expected values for them (test-cluster_genes.R only runs the call),
so they are restated from the R source alone.

Conventions: `counts` is a scipy.sparse.csc_matrix, genes x cells (the
`dgCMatrix` of counts(cds), R/cell_data_set.R:117,123).
"""
import math

import numpy as np
import scipy.sparse as sp

from . import irlba_ref
from .r_rng import RRandom


def r_mean(x):
    """R's mean() for doubles (summary.c): long-double sum, then one refinement pass."""
    x = np.asarray(x, dtype=np.float64)
    n = x.size
    xl = x.astype(np.longdouble)
    s = np.cumsum(xl)[-1] / n  # sequential accumulation in long double
    t = np.cumsum(xl - s)[-1] / n
    return np.float64(s + t)


def _libm_log(x):
    """Element-wise natural log through the C library, like R's log().  np.log is a SIMD routine of its
    own that is not always correctly rounded: it differs from libm in about 1 of 200 000 arguments
    (and log(log(x)) in 1 of 2 000), which would show up as a last-bit difference in a size factor."""
    import math
    x = np.asarray(x, dtype=np.float64)
    out = np.empty_like(x)
    flat_in, flat_out = x.ravel(), out.ravel()
    for k in range(flat_in.size):
        v = flat_in[k]
        flat_out[k] = math.log(v) if v > 0 else (-np.inf if v == 0 else np.nan)
    return out


def estimate_sf_sparse(counts, round_exprs=True, method="mean-geometric-mean-total"):
    """R/utils.R:54-70.  Returns (size_factors, cell_total)."""
    x = counts.data
    if round_exprs:
        x = np.round(x)  # R's round() is half-to-even on .5 for doubles, like np.round
    c = sp.csc_matrix((x, counts.indices, counts.indptr), shape=counts.shape)
    cell_total = np.asarray(c.sum(axis=0)).ravel()
    with np.errstate(divide="ignore", invalid="ignore"):
        if method == "mean-geometric-mean-total":
            sfs = cell_total / math.exp(float(r_mean(_libm_log(cell_total))))
        elif method == "mean-geometric-mean-log-total":
            lt = _libm_log(cell_total)
            sfs = lt / math.exp(float(r_mean(_libm_log(lt))))
        else:
            raise ValueError("method")
    sfs = np.where(np.isnan(sfs), 1.0, sfs)  # R/utils.R:68
    return sfs, cell_total


def estimate_size_factors(counts, round_exprs=True, method="mean-geometric-mean-total"):
    """R/utils.R:26-50: returns None (cds unchanged + warning) if any cell has zero reads."""
    raw_tot = np.asarray(counts.sum(axis=0)).ravel()
    if np.any(raw_tot == 0):
        return None
    return estimate_sf_sparse(counts, round_exprs, method)[0]


def normalize_expr_data(counts, sf, norm_method="log", pseudo_count=None):
    """R/preprocess_cds.R:253-286.  Returns FM as csc (sparse path) or ndarray (dense path)."""
    if pseudo_count is None:
        pseudo_count = 1 if norm_method == "log" else 0
    FM = counts.copy().astype(np.float64)
    if norm_method == "log":
        FM.data = FM.data / np.repeat(sf, np.diff(FM.indptr))  # t(t(FM)/sf)  :272
        if pseudo_count != 1:
            D = FM.toarray() + pseudo_count  # :274-276 (densifies)
            return np.log2(D)
        FM.data = np.log2(FM.data + 1)  # :278
        return FM
    if norm_method == "size_only":
        FM.data = FM.data / np.repeat(sf, np.diff(FM.indptr))  # :282
        if pseudo_count != 0:
            return FM.toarray() + pseudo_count  # :283 (densifies)
        return FM
    if norm_method == "none":
        return FM
    raise ValueError("norm_method must be one of 'log', 'size_only' or 'none'")


def normalized_counts(counts, sf, norm_method="log", pseudocount=1):
    """R/utils.R:477-509 (sparse branch). NOTE log10, not log2."""
    nm = counts.copy().astype(np.float64)
    if norm_method == "binary":
        nm.data = (nm.data > 0).astype(np.float64)
        nm.eliminate_zeros()
        return nm
    nm.data = nm.data / np.repeat(sf, np.diff(nm.indptr))  # :493
    if norm_method == "log":
        if pseudocount != 1:
            raise ValueError("Pseudocount must equal 1 with sparse expression matrices")
        nm.data = np.log10(nm.data + pseudocount)  # :496
    elif norm_method != "size_only":
        raise ValueError("norm_method")
    return nm


def select_and_filter_genes(FM, use_genes_idx=None):
    """R/preprocess_cds.R:117-122. Returns (FM', kept_original_gene_indices)."""
    idx = np.arange(FM.shape[0]) if use_genes_idx is None else np.asarray(use_genes_idx)
    if use_genes_idx is not None:
        FM = FM[idx, :] if not sp.issparse(FM) else FM.tocsr()[idx, :].tocsc()
    rs = np.asarray(FM.sum(axis=1)).ravel()
    keep = np.isfinite(rs) & (rs != 0)
    FM = FM[keep, :] if not sp.issparse(FM) else FM.tocsr()[keep, :].tocsc()
    return FM, idx[keep]


def gene_moments(At):
    """R/utils.R:383,389: colMeans2 / sqrt(colVars) of the cells x genes matrix (n-1 denominator)."""
    C = At.shape[0]
    if sp.issparse(At):
        A = At.tocsc()
        mean = np.asarray(A.sum(axis=0)).ravel() / C
        nnz = np.diff(A.indptr)
        dev = A.data - np.repeat(mean, nnz)
        ss = np.zeros(A.shape[1])
        np.add.at(ss, np.repeat(np.arange(A.shape[1]), nnz), dev * dev)
        ss += (C - nnz) * mean * mean
    else:
        mean = At.mean(axis=0)
        ss = ((At - mean) ** 2).sum(axis=0)
    return mean, np.sqrt(ss / (C - 1))


def sparse_prcomp_irlba(At, n, center=True, scale=False, v0=None, **irlba_args):
    """R/utils.R:365-430 with logical center/scale. (monocle3 always passes center == scale.)"""
    C, G = At.shape
    cen = sc = None
    if center:
        cen, sd = gene_moments(At)
        if scale:
            sc = sd
    elif scale:
        # R/utils.R:396-400: no centering -> root-mean-square scaling
        sq = At.multiply(At) if sp.issparse(At) else At * At
        sc = np.sqrt(np.asarray(sq.sum(axis=0)).ravel() / max(1, C - 1))
    if v0 is None:
        rng = RRandom(2016)  # set.seed(2016) at R/preprocess_cds.R:110, rnorm(n) inside irlba
        v0 = rng.rnorm(G)
        irlba_args.setdefault("rng", rng)  # the same stream serves irlba's invariant-subspace restarts
    s = irlba_ref.irlba(At, n, v0, center=cen, scale=sc, **irlba_args)
    if s.err not in (0, -2):
        raise RuntimeError("irlba error %d" % s.err)
    ans = dict(sdev=s.d / np.sqrt(max(1, C - 1)),  # :418
               rotation=s.v,  # :419
               center=cen, svd_scale=sc,
               x=s.u * s.d[None, :],  # sweep(u,2,d,'*')  :425
               d=s.d, u=s.u, iter=s.iter, mprod=s.mprod, err=s.err)
    return ans


def preprocess_cds(counts, sf, num_dim=50, norm_method="log", use_genes_idx=None,
                   pseudo_count=None, scaling=True, v0=None, **irlba_args):
    """PCA branch of R/preprocess_cds.R:62-248.  Returns the model dict + embedding.

    use_genes_idx: integer row indices (R passes names; translation is host glue).
    """
    FM = normalize_expr_data(counts, sf, norm_method, pseudo_count)  # :111
    if FM.shape[0] == 0:
        raise ValueError("all rows have standard deviation zero")
    FM, kept = select_and_filter_genes(FM, use_genes_idx)  # :117-122
    At = FM.T
    if sp.issparse(At):
        At = At.tocsc()
    n = min(num_dim, min(FM.shape) - 1)  # :140
    res = sparse_prcomp_irlba(At, n, center=scaling, scale=scaling, v0=v0, **irlba_args)
    sdev = res["sdev"]
    return dict(PCA=res["x"], svd_v=res["rotation"], svd_sdev=sdev,
                svd_center=res["center"], svd_scale=res["svd_scale"],
                prop_var_expl=sdev ** 2 / np.sum(sdev ** 2),  # :158-160
                kept_genes=kept, num_dim=num_dim, norm_method=norm_method,
                pseudo_count=pseudo_count, d=res["d"], u=res["u"],
                iter=res["iter"], mprod=res["mprod"], err=res["err"])


def tfidf(FM, frequencies=True, log_scale_tf=True, scale_factor=100000.0):
    """R/preprocess_cds.R:289-338 ("Andrew's tfidf") on a genes x cells csc matrix."""
    FM = FM.tocsc().astype(np.float64)
    num_cols = FM.shape[1]
    if frequencies:
        col_sums = np.asarray(FM.sum(axis=0)).ravel()
        tf = FM.copy()
        tf.data = tf.data / np.repeat(col_sums, np.diff(tf.indptr))  # t(t(count_matrix) / col_sums)
    else:
        col_sums = None
        tf = FM.copy()
    if log_scale_tf:
        tf.data = np.log1p(tf.data * (scale_factor if frequencies else 1.0))
    row_sums = np.asarray((FM > 0).sum(axis=1)).ravel().astype(np.float64)  # Matrix::rowSums(count_matrix > 0)
    with np.errstate(divide="ignore"):
        idf = np.log(1.0 + num_cols / row_sums)
    out = sp.diags(idf) @ tf  # tf * idf recycles idf down the rows
    return dict(tf_idf_counts=out.tocsc(), frequencies=frequencies, log_scale_tf=log_scale_tf,
                scale_factor=scale_factor, col_sums=col_sums, row_sums=row_sums, num_cols=num_cols)


def preprocess_cds_lsi(counts, sf, num_dim=50, norm_method="log", use_genes_idx=None, pseudo_count=None,
                       v0=None, **irlba_args):
    """LSI branch of R/preprocess_cds.R:185-241: TF-IDF then irlba without centre/scale."""
    FM = normalize_expr_data(counts, sf, norm_method, pseudo_count)
    if not sp.issparse(FM):
        raise NotImplementedError("dense pseudo-count path with LSI")
    FM, kept = select_and_filter_genes(FM, use_genes_idx)
    t = tfidf(FM)
    At = t["tf_idf_counts"].T.tocsc()
    num_col = At.shape[1]  # R: num_col <- ncol(preproc_res) = number of CELLS of the genes x cells matrix
    num_col = t["tf_idf_counts"].shape[1]
    n = min(num_dim, min(FM.shape) - 1)
    if v0 is None:
        v0 = RRandom(2016).rnorm(At.shape[1])
    s = irlba_ref.irlba(At, n, v0, **irlba_args)
    if s.err not in (0, -2):
        raise RuntimeError("irlba error %d" % s.err)
    return dict(LSI=s.u * s.d[None, :], svd_v=s.v, svd_sdev=s.d / np.sqrt(max(1, num_col - 1)),
                col_sums=t["col_sums"], row_sums=t["row_sums"], num_cols=t["num_cols"], kept_genes=kept,
                frequencies=True, log_scale_tf=True, scale_factor=100000.0, norm_method=norm_method,
                pseudo_count=pseudo_count, d=s.d, iter=s.iter, mprod=s.mprod)


def preprocess_transform_lsi(counts, sf, model, gene_index_in_model):
    """LSI branch of R/projection.R:156-241: the query's own column sums, the MODEL's idf."""
    FM = normalize_expr_data(counts, sf, model["norm_method"], model.get("pseudo_count"))
    rs = np.asarray(FM.sum(axis=1)).ravel()
    keep = np.isfinite(rs) & (rs != 0)
    gidx = np.asarray(gene_index_in_model)
    q_rows = np.nonzero(keep & (gidx >= 0))[0]
    order = np.argsort(gidx[q_rows], kind="stable")
    q_rows = q_rows[order]
    m_rows = gidx[q_rows]
    F = FM.tocsr()[q_rows, :].tocsc()
    cs = np.asarray(F.sum(axis=0)).ravel()
    tf = F.copy()
    tf.data = np.log1p(tf.data / np.repeat(cs, np.diff(tf.indptr)) * model["scale_factor"])
    idf = np.log(1.0 + model["num_cols"] / model["row_sums"])[m_rows]
    X = (sp.diags(idf) @ tf).T
    return np.asarray(X @ model["svd_v"][m_rows, :])


def preprocess_transform(counts, sf, model, gene_index_in_model, keep_override=None):
    """PCA branch of R/projection.R:102-149.

    gene_index_in_model[g] = row of model['svd_v'] for query gene g, or -1 if the
    query gene is not in the model (R does this by rowname intersect, :121-122).
    keep_override: the zero-gene filter of a LARGER query this block of cells belongs to (the filter
    is a property of the whole query matrix; bench.py checks a block of cells of a big query).
    """
    FM = normalize_expr_data(counts, sf, model["norm_method"], model.get("pseudo_count"))
    rs = np.asarray(FM.sum(axis=1)).ravel()
    keep = np.isfinite(rs) & (rs != 0)  # :117-118
    if keep_override is not None:
        keep = np.asarray(keep_override, dtype=bool)
    gidx = np.asarray(gene_index_in_model)
    # R: intersect(rownames(rotation), rownames(FM)) keeps the MODEL's order (:121)
    q_rows = np.nonzero(keep & (gidx >= 0))[0]
    order = np.argsort(gidx[q_rows], kind="stable")
    q_rows = q_rows[order]
    m_rows = gidx[q_rows]
    X = FM.tocsr()[q_rows, :].T if sp.issparse(FM) else FM[q_rows, :].T  # cells x genes
    V = model["svd_v"][m_rows, :]
    if model["svd_center"] is not None:
        c = model["svd_center"][m_rows]
        s = model["svd_scale"][m_rows]
        Xd = (np.asarray(X.todense()) if sp.issparse(X) else X)
        return ((Xd - c[None, :]) / s[None, :]) @ V  # :141-148, densified like DelayedArray does
    return np.asarray(X @ V)


def detect_genes(counts, min_expr=0):
    """R/utils.R:449-457: (rowSums(counts > min_expr), colSums(counts > min_expr)).  PARITY UNPINNED."""
    C = counts.tocsc()
    G, N = C.shape
    on = C.data > min_expr
    per_gene = np.bincount(C.indices[on], minlength=G).astype(np.int64)
    per_cell = np.add.reduceat(np.concatenate([on.astype(np.int64), [0]]), C.indptr[:-1])
    per_cell = np.where(np.diff(C.indptr) > 0, per_cell, 0).astype(np.int64)
    if min_expr < 0:  # the implicit zeros qualify as well
        per_gene += N - np.bincount(C.indices, minlength=G)
        per_cell += G - np.diff(C.indptr)
    return per_gene, per_cell


def _factor_levels(labels):
    """as.factor(): sorted unique values (numbers numerically, strings lexicographically)."""
    return sorted(set(labels))


def aggregate_gene_expression(counts, sf, gene_names, cell_names, gene_group_df=None, cell_group_df=None,
                              norm_method="log", pseudocount=1, scale_agg_values=True, max_agg_value=3,
                              min_agg_value=-3, exclude_na=True, gene_agg_fun="sum", cell_agg_fun="mean"):
    """R/cluster_genes.R:447-536.  *_group_df: sequences of (id, group) pairs.  Returns
    (matrix, row_names, col_names).  PARITY UNPINNED (no expected values in the reference's tests)."""
    if gene_group_df is None and cell_group_df is None:
        raise ValueError("one of either gene_group_df or cell_group_df must not be NULL.")
    agg = normalized_counts(counts, sf, norm_method=norm_method, pseudocount=pseudocount).tocsr()
    row_names, col_names = list(gene_names), list(cell_names)
    if gene_group_df is not None:
        glut = {g: k for k, g in enumerate(gene_names)}
        pairs = [(g, grp) for g, grp in gene_group_df if g in glut]           # :466-469
        groups = []
        for _, grp in pairs:                                                  # unique(), first appearance
            if grp not in groups:
                groups.append(grp)
        rows, names = [], []
        for grp in groups:
            members = []
            for g, gg in pairs:
                if gg == grp and g not in members:
                    members.append(g)
            if len(members) < 2:                                              # :490-491 a single row has no dim
                continue
            sub = agg[[glut[g] for g in members], :]
            res = np.asarray(sub.sum(axis=0)).ravel()
            if gene_agg_fun == "mean":
                res = res / len(members)
            rows.append(res)
            names.append(grp)
        agg = np.vstack(rows) if rows else np.zeros((0, agg.shape[1]))
        row_names = names
    else:
        agg = agg.toarray()
    if cell_group_df is not None:
        clut = {c: k for k, c in enumerate(cell_names)}
        pairs = [(c, grp) for c, grp in cell_group_df if c in clut]           # :517-518
        sub = agg[:, [clut[c] for c, _ in pairs]]                             # :519
        labels = [grp for _, grp in pairs]
        levels = _factor_levels(labels)                                       # as.factor
        out = np.zeros((agg.shape[0], len(levels)))
        for q, lv in enumerate(levels):
            idx = [k for k, l in enumerate(labels) if l == lv]
            out[:, q] = sub[:, idx].sum(axis=1)
            if cell_agg_fun == "mean":
                out[:, q] /= len(idx)
        agg = out
        col_names = levels
    if scale_agg_values:                                                      # :526-536
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


def jaccard_coeff(idx, weight=True):
    """src/clustering.cpp:13-46 (jaccard_coeff_cpp), loop for loop.  idx: n x k, 1-based neighbour
    ids.  Returns the (n*k) x 3 matrix (from, to, weight / max weight).  Integer arithmetic and two
    IEEE divisions: the device result must be bit-identical.  (The reference's tests hold no
    expected values for this kernel either; the restatement is checked on hand-worked cases in
    tests/test_oracle_aggregate.py.)"""
    idx = np.asarray(idx)
    n, k = idx.shape
    out = np.zeros((n * k, 3))
    sets = [set(int(v) for v in idx[i]) for i in range(n)]
    r = 0
    for i in range(n):
        for j in range(k):
            kk = int(idx[i, j]) - 1
            w = 1.0
            if weight:
                u = len(sets[i] & sets[kk])          # intersect(nodei, nodej).size(): unique common values
                v = 2 * k - u
                if u > 0:
                    w = float(u) / float(v)
            out[r] = (i + 1, kk + 1, w)
            r += 1
    out[:, 2] = out[:, 2] / out[:, 2].max()
    return out


def nn2_self(X, k):
    """RANN::nn2(X, X, k, searchtype = "standard") as cluster_cells calls it (R/cluster_cells.R:286-287):
    exact Euclidean k nearest neighbours of every row among all rows (the row itself included),
    sorted by distance; returns (nn_idx 1-based, nn_dists).  RANN (a kd-tree over the same distances)
    is not in /root/reference; the result is defined by the distances alone, ties (not defined by
    RANN) go to the smaller index.  PARITY UNPINNED."""
    X = np.asarray(X, dtype=np.float64)
    n = X.shape[0]
    idx = np.zeros((n, k), dtype=np.int64)
    dist = np.zeros((n, k))
    for i in range(n):
        d2 = ((X - X[i]) ** 2).sum(axis=1)          # direct differences, like a kd-tree leaf scan
        o = np.argsort(d2, kind="stable")[:k]
        idx[i] = o + 1
        dist[i] = np.sqrt(d2[o])
    return idx, dist


def cluster_cells_make_graph(data, weight, k):
    """R/cluster_cells.R:255-336 with nn_control method "nn2", up to the relations table:
    (relations (n*k) x 3 = from, to, weight; distMatrix n x k).  PARITY UNPINNED."""
    data = np.asarray(data, dtype=np.float64)
    if k < 1:
        raise ValueError("k must be a positive integer!")
    if k > data.shape[0] - 2:                         # :269-275
        k = data.shape[0] - 2
    idx, dist = nn2_self(data, k + 1)                 # :287
    neighbor, distm = idx[:, 1:], dist[:, 1:]        # :303-304
    links = jaccard_coeff(neighbor, weight)           # :313
    links = links[links[:, 0] > 0]                    # :318
    return links, distm


# ------------------------------------------------------------------------------------------------
# align_cds residual regression (R/alignment.R:106-117) and align_beta_transform
# (R/projection.R:468-478).  limma::lmFit(t(X), design) without weights / blocking fits every row of
# t(X) (every PC) by ordinary least squares through lm.fit's pivoted QR (LINPACK dqrdc2, tolerance
# 1e-7: a design column whose residual norm after projecting out the columns before it falls below
# tol x its norm is aliased -> coefficient NA); align_cds zeroes the NAs, drops the intercept and
# subtracts the fitted effects.  limma is not in /root/reference (DESCRIPTION: Imports), so this is
# restated from the published lm.fit semantics; the reference pins one value, beta[[1]] = 2.071
# (tests/testthat/test-alignment.R:100,139), on data drawn with R's rmultinom, which cannot be
# reproduced without R -- tests/test_oracle_align.py checks it statistically.
# ------------------------------------------------------------------------------------------------
def lm_fit_coefficients(Y, M, tol=1e-7):
    """Coefficients (p x q, NaN for aliased design columns) of the least-squares fits of the columns
    of Y (n x q) on the design M (n x p), with lm.fit's limited column pivoting."""
    Y = np.asarray(Y, dtype=np.float64)
    M = np.asarray(M, dtype=np.float64)
    n, p = M.shape
    Q = np.zeros((n, 0))
    used = []
    for j in range(p):
        col = M[:, j]
        r = col - Q @ (Q.T @ col)
        r = r - Q @ (Q.T @ r)  # second pass: Householder-grade orthogonality
        nrm0 = np.linalg.norm(col)
        if nrm0 == 0 or np.linalg.norm(r) < tol * nrm0:
            continue  # aliased: moved to the end by dqrdc2, coefficient NA
        Q = np.hstack([Q, (r / np.linalg.norm(r))[:, None]])
        used.append(j)
    coef = np.full((p, Y.shape[1]), np.nan)
    if used:
        sol, *_ = np.linalg.lstsq(M[:, used], Y, rcond=None)
        coef[used, :] = sol
    return coef


def align_residuals(X, M):
    """R/alignment.R:112-116: beta = fit$coefficients[, -1] (PCs x terms, NA -> 0);
    X_aligned = t(t(X) - beta %*% t(M[, -1])).  X: cells x PCs, M: cells x p with the intercept first.
    Returns (X_aligned, beta)."""
    coef = lm_fit_coefficients(X, M)          # p x nv
    beta = coef[1:, :].T.copy()               # nv x (p - 1)
    beta[np.isnan(beta)] = 0.0
    return np.asarray(X) - np.asarray(M)[:, 1:] @ beta.T, beta


def model_matrix(columns, terms):
    """The subset of stats::model.matrix / Matrix::sparse.model.matrix this path needs: main effects
    of numeric columns and of factors with treatment contrasts (first level dropped, levels sorted
    like as.factor), intercept first.  columns: dict name -> sequence; terms: list of names.
    Returns (M, column_names)."""
    n = len(next(iter(columns.values()))) if columns else 0
    cols, names = [np.ones(n)], ["(Intercept)"]
    for t in terms:
        v = np.asarray(columns[t])
        if v.dtype.kind in "fiub":
            cols.append(v.astype(np.float64))
            names.append(t)
        else:
            levels = sorted(set(v.tolist()))
            for lv in levels[1:]:
                cols.append((v == lv).astype(np.float64))
                names.append("%s%s" % (t, lv))
    return np.stack(cols, axis=1), names
