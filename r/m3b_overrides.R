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
                                row_of_gene, model$svd_v, model$svd_center, model$svd_scale)
  rownames(y) <- colnames(cds); y
}
