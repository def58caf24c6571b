# R shim of the drop-in: same function names and signatures as monocle3; only the BODIES of the
# dgCMatrix branches change.  Everything else (argument checks, set.seed(2016), dimnames, model
# slots, identity bookkeeping, nn index) is monocle3's own code and stays as it is.
# NOT executed in this image (no R).  Line references are into the monocle3 source tree.

# replaces estimate_sf_sparse, R/utils.R:54-70
estimate_sf_sparse <- function(counts, round_exprs=TRUE, method="mean-geometric-mean-total") {
  if (methods::is(counts, "dgCMatrix"))
    return(m3b_estimate_sf_sparse(counts, round_exprs,
                                  match(method, c("mean-geometric-mean-total", "mean-geometric-mean-log-total")) - 1L))
  stop("m3b: only dgCMatrix counts are supported on the GPU path")
}

# replaces the PCA branch body of preprocess_cds, R/preprocess_cds.R:111-160
.m3b_pca <- function(cds, num_dim, norm_method, use_genes, pseudo_count, scaling) {
  if (is.null(pseudo_count)) pseudo_count <- if (norm_method == "log") 1 else 0        # :262-267
  gene_order <- if (is.null(use_genes)) NULL else match(use_genes, rownames(cds)) - 1L   # :117-119
  set.seed(2016)                                                                         # :110
  v0 <- rnorm(nrow(cds))            # irlba draws rnorm(G'); the first G' values of this stream are the same draw
  res <- m3b_preprocess_pca(SingleCellExperiment::counts(cds), size_factors(cds),
                            match(norm_method, c("log", "size_only", "none")) - 1L, pseudo_count,
                            gene_order, num_dim, scaling, v0)
  kept <- (if (is.null(use_genes)) rownames(cds) else use_genes)[as.logical(res$keep)]
  x <- res$x; rownames(x) <- colnames(cds); colnames(x) <- paste0("PC", seq_len(ncol(x)))    # :142-144, R/utils.R:426
  v <- res$v; rownames(v) <- kept;          colnames(v) <- colnames(x)                       # :147
  sdev <- res$d / sqrt(max(1, ncol(cds) - 1))                                                # R/utils.R:418
  list(x = x, rotation = v, sdev = sdev,
       center = if (scaling) res$center else NULL, svd_scale = if (scaling) res$scale else NULL)
}

# replaces the multiply of preprocess_transform, R/projection.R:109-149
.m3b_transform <- function(cds, model) {
  row_of_gene <- match(rownames(cds), rownames(model$svd_v)) - 1L
  row_of_gene[is.na(row_of_gene)] <- -1L
  pc <- model$pseudo_count; if (is.null(pc)) pc <- if (model$norm_method == "log") 1 else 0
  y <- m3b_preprocess_transform(SingleCellExperiment::counts(cds), size_factors(cds),
                                match(model$norm_method, c("log", "size_only", "none")) - 1L, pc,
                                row_of_gene, model$svd_v, model$svd_center, model$svd_scale,
                                model$row_sums, if (is.null(model$num_cols)) 0 else model$num_cols)  # LSI: the MODEL's idf
  rownames(y) <- colnames(cds); y
}

# replaces the LSI branch body of preprocess_cds, R/preprocess_cds.R:185-241 (tfidf() + irlba without centre / scale)
.m3b_lsi <- function(cds, num_dim, norm_method, use_genes, pseudo_count) {
  if (is.null(pseudo_count)) pseudo_count <- if (norm_method == "log") 1 else 0
  gene_order <- if (is.null(use_genes)) NULL else match(use_genes, rownames(cds)) - 1L
  set.seed(2016)
  v0 <- rnorm(nrow(cds))
  res <- m3b_preprocess_lsi(SingleCellExperiment::counts(cds), size_factors(cds),
                            match(norm_method, c("log", "size_only", "none")) - 1L, pseudo_count, gene_order, num_dim, v0)
  kept <- (if (is.null(use_genes)) rownames(cds) else use_genes)[as.logical(res$keep)]
  x <- res$x; rownames(x) <- colnames(cds)                                                   # :196-197
  v <- res$v; rownames(v) <- kept                                                            # :199-200
  list(x = x, v = v, d = res$d, col_sums = res$col_sums, row_sums = res$row_sums, num_cols = res$num_cols,
       sdev = res$d / sqrt(max(1, res$num_cols - 1)))                                        # :213
}

# replaces the sparse branch of normalized_counts, R/utils.R:488-503
.m3b_normalized_counts <- function(cds, norm_method = c("log", "binary", "size_only"), pseudocount = 1) {
  norm_method <- match.arg(norm_method)
  norm_mat <- SingleCellExperiment::counts(cds)
  if (norm_method == "log" && pseudocount != 1) stop("Pseudocount must equal 1 with sparse expression matrices")
  code <- c(log = 3L, size_only = 1L, binary = 4L)[[norm_method]]        # M3B_NORM_LOG10 / SIZE_ONLY / BINARY
  norm_mat@x <- m3b_normalized_counts_x(norm_mat, size_factors(cds), code, if (norm_method == "log") 1 else 0)
  if (norm_method == "binary") norm_mat <- Matrix::drop0(norm_mat)
  norm_mat
}

# replaces lmFit + the subtraction in align_cds, R/alignment.R:112-117
.m3b_align_residuals <- function(preproc_res, X.model_mat) {
  res <- m3b_align_residuals_(as.matrix(preproc_res), as.matrix(X.model_mat))
  beta <- res$beta
  rownames(beta) <- colnames(preproc_res); colnames(beta) <- colnames(X.model_mat)[-1]
  aligned <- res$aligned; dimnames(aligned) <- dimnames(preproc_res)
  list(preproc_res = aligned, beta = beta)
}

# replaces the subtraction in align_beta_transform, R/projection.R:477-478
.m3b_align_beta_transform <- function(preproc_res, X.model_mat, beta) {
  out <- m3b_align_apply_beta_(as.matrix(preproc_res), as.matrix(X.model_mat), as.matrix(beta))
  dimnames(out) <- dimnames(preproc_res)
  out
}
