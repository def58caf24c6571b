/*
 * m3b.h -- C ABI of the B200-native preprocess_cds PCA path (monocle3 drop-in).
 *
 * This is the whole boundary between the reference's R code and the CUDA
 * (sm_100a) implementation.  Plain C types only: no R, Rcpp or torch types.
 * The reference has no native entry point for this path; what each function
 * replaces is the R-level call into a third-party package that the reference
 * makes today (paths relative to the monocle3 source tree):
 *
 *   m3b_cell_totals / m3b_size_factors ... estimate_sf_sparse: round() + Matrix::colSums
 *                                          R/utils.R:54-70 (called from R/utils.R:26-50)
 *   m3b_normalize / m3b_export_values .... normalize_expr_data R/preprocess_cds.R:253-286,
 *                                          normalized_counts R/utils.R:477-509
 *   m3b_select_genes ..................... FM[use_genes,], Matrix::rowSums zero-gene filter,
 *                                          Matrix::t(FM)  R/preprocess_cds.R:117-122,139
 *   m3b_gene_moments ..................... DelayedMatrixStats::colMeans2 / colVars
 *                                          R/utils.R:383,389
 *   m3b_pca_irlba ........................ irlba::irlba(A=t(FM), nv, center, scale) and the
 *                                          post-processing R/utils.R:417-428
 *   m3b_project .......................... preprocess_transform's delayed (x-c)/s %*% V,
 *                                          R/projection.R:135-148, R/utils.R:823-849
 *
 * The registration pattern a maintainer mirrors is src/RcppExports.cpp:15-47
 * (see INTEGRATION.md for the Rcpp stubs).
 *
 * Conventions
 *   - G = genes (rows of counts(cds)), C = cells (columns).  The input is the
 *     dgCMatrix of counts(cds): CSC, cell-major, 0-based int32 row indices,
 *     f64 values (R/cell_data_set.R:117,123).  Because a dgCMatrix cannot hold
 *     more than 2^31-1 non-zeros, matrices are appended as cell-range chunks.
 *   - Dense matrices are column-major (R layout) with an explicit leading
 *     dimension where noted.
 *   - All host pointers are caller-owned and borrowed only for the duration of
 *     the call.  Device memory is owned by the context / matrix handles.
 *   - Every function returns an int status: 0 on success, a negative M3B_E_*
 *     code otherwise; m3b_last_error() gives the message.  Nothing throws,
 *     exits or longjmps across this boundary.  M3B_W_NOT_CONVERGED (-2) mirrors
 *     irlba's "did not converge" warning: results are still written.
 *   - One context drives one GPU.  Multi-GPU runs use one context per process
 *     (or thread), each owning a contiguous cell range, joined by
 *     m3b_comm_init(); gene-space results are replicated, cell-space results
 *     stay sharded.
 *   - There is no CPU fallback: if no CUDA device is usable, m3b_create fails.
 */
#ifndef M3B_H
#define M3B_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define M3B_ABI_VERSION 3  /* 3: m3b_stats gained xchg_diag[8] at its end; m3b_jaccard_coeff accepts k <= 2048 */

/* status codes (negative = error; the first five mirror irlba's C path) */
#define M3B_OK                0
#define M3B_E_DIMS           -1  /* invalid dimensions / arguments                  */
#define M3B_W_NOT_CONVERGED  -2  /* maxit reached; outputs written (irlba warning)  */
#define M3B_E_NOMEM          -3  /* host or device allocation failed                */
#define M3B_E_NULLSPACE      -4  /* start vector in the null space of A             */
#define M3B_E_LINDEP         -5  /* linear dependency / invariant subspace hit      */
#define M3B_E_CUDA           -10 /* CUDA runtime failure (message has the details)  */
#define M3B_E_NCCL           -11 /* NCCL failure                                    */
#define M3B_E_STATE          -12 /* call sequence violated (e.g. normalize missing) */
#define M3B_E_NODEVICE       -13 /* no usable sm_100 device: there is no fallback   */
#define M3B_E_ABORTED        -14 /* should_abort callback asked to stop             */

/* (m3b_tfidf is documented next to m3b_gene_moments below) */
/* norm_method codes for m3b_normalize */
#define M3B_NORM_LOG2       0 /* log2(x/sf + pc)   preprocess_cds norm_method="log"        */
#define M3B_NORM_SIZE_ONLY  1 /* x/sf + pc         norm_method="size_only"                 */
#define M3B_NORM_NONE       2 /* x                 norm_method="none"                      */
#define M3B_NORM_LOG10      3 /* log10(x/sf + 1)   normalized_counts(norm_method="log")    */
#define M3B_NORM_BINARY     4 /* (x > 0) + 0       normalized_counts(norm_method="binary") */

/* size-factor methods (R/utils.R:60-66) */
#define M3B_SF_MEAN_GEOMETRIC_MEAN_TOTAL      0
#define M3B_SF_MEAN_GEOMETRIC_MEAN_LOG_TOTAL  1

typedef struct m3b_ctx m3b_ctx;
typedef struct m3b_mat m3b_mat;

/* Optional callbacks.  small_svd lets the R shim supply R's own LAPACK dgesdd
 * for the (work x work) bidiagonal matrix so signs match the reference bit for
 * bit; the built-in default is a one-sided Jacobi SVD.  All matrices are
 * column-major n x n: on entry B; on exit U (n x n), s (n, descending),
 * Vt (n x n, rows are right singular vectors).  Return 0 on success. */
typedef int (*m3b_small_svd_fn)(int n, const double* B, double* U, double* s, double* Vt, void* user);
/* Polled once per restart (irlba polls R_CheckUserInterrupt after each product). */
typedef int (*m3b_should_abort_fn)(void* user);
/* Standard normals from the CALLER's random stream (the R shim: rnorm(n)).  irlba replaces a
 * Lanczos vector by a random one when the process hits an invariant subspace (irlb.c: R_F < eps or
 * S < eps); with this callback set the library does the same (single rank), without it that case
 * returns M3B_E_LINDEP.  Return 0 on success. */
typedef int (*m3b_rnorm_fn)(int64_t n, double* out, void* user);

/* ---- library / context ------------------------------------------------ */
int         m3b_abi_version(void);
const char* m3b_status_string(int status);
int         m3b_device_count(void);

int         m3b_create(int device, int flags, m3b_ctx** out);
void        m3b_destroy(m3b_ctx* ctx);
const char* m3b_last_error(const m3b_ctx* ctx);
int         m3b_set_small_svd(m3b_ctx* ctx, m3b_small_svd_fn fn, void* user);
int         m3b_set_should_abort(m3b_ctx* ctx, m3b_should_abort_fn fn, void* user);
int         m3b_set_rnorm(m3b_ctx* ctx, m3b_rnorm_fn fn, void* user);

/* ---- multi-GPU plumbing (one context per rank) ------------------------ */
#define M3B_COMM_ID_BYTES 128
int m3b_comm_unique_id(m3b_ctx* ctx, char id[M3B_COMM_ID_BYTES]);              /* rank 0 */
int m3b_comm_init(m3b_ctx* ctx, int rank, int world, const char id[M3B_COMM_ID_BYTES]);
int m3b_comm_world(const m3b_ctx* ctx);                                        /* 1 if unsharded */
/* Optional, one box only: peer-memory exchange over NVLink for the per-step cross-rank sums of
 * the Lanczos loop (the gene-length partial of A^T w, the Gram-Schmidt coefficients of the cell
 * side and {||w||^2, sum w}): they then run INSIDE the fused vector kernel, which reads the peers'
 * contributions directly; NCCL stays the fallback and serves the preparation passes.  After
 * m3b_comm_init every rank exports a 64-byte CUDA-IPC handle; the host gathers the world's
 * handles (rank order) and hands them back to every rank. */
#define M3B_P2P_HANDLE_BYTES 64
int m3b_comm_p2p_handle(m3b_ctx* ctx, char handle[M3B_P2P_HANDLE_BYTES]);
int m3b_comm_p2p_open(m3b_ctx* ctx, const char* handles /* world x 64 bytes */);

/* ---- sparse ingest (K0) ------------------------------------------------ */
/* n_cells_local cells will be appended to this context; n_cells_total is the
 * cell count over all ranks (== n_cells_local when unsharded). */
int  m3b_matrix_begin(m3b_ctx* ctx, int32_t n_genes, int64_t n_cells_local,
                      int64_t n_cells_total, int64_t nnz_hint, m3b_mat** out);
/* p: chunk-local column pointers (n_cells_chunk+1, p[0]==0, non-decreasing); i: row indices in
 * [0, n_genes), STRICTLY INCREASING inside every column (sorted and unique, as in a dgCMatrix);
 * x: values.  The chunk is validated on the device; a violation returns M3B_E_DIMS and the chunk
 * is not appended. */
int  m3b_matrix_append_csc(m3b_mat* mat, int64_t n_cells_chunk, const int32_t* p,
                           const int32_t* i, const double* x);
int  m3b_matrix_end(m3b_mat* mat);
void m3b_matrix_free(m3b_mat* mat);
int  m3b_matrix_dims(const m3b_mat* mat, int32_t* n_genes, int64_t* n_cells_local,
                     int64_t* n_cells_total, int64_t* nnz_local);

/* ---- size factors (K1) -------------------------------------------------- */
/* Per-cell sums of (optionally rounded) counts for the local cells. Exact. */
int m3b_cell_totals(m3b_mat* mat, int round_exprs, double* cell_total /* n_cells_local */);
/* Host-only finish, given the totals of ALL cells in cell order (R/utils.R:61-68):
 * long-double two-pass mean like R's mean(); NA -> 1.  any_zero reports the
 * zero-read-cell guard of R/utils.R:32-39 (sf is still written). */
int m3b_size_factors_from_totals(const double* cell_total, int64_t n_cells, int method,
                                 double* sf /* n_cells */, int* any_zero);
/* Convenience for the unsharded case: totals + finish. */
int m3b_size_factors(m3b_mat* mat, int round_exprs, int method, double* cell_total,
                     double* sf, int* any_zero);

/* ---- normalisation (K2) ------------------------------------------------- */
/* sf: size factors of the local cells (ignored for NONE/BINARY).  pseudo_count
 * follows normalize_expr_data: LOG2 with pc != 1 and SIZE_ONLY with pc != 0 are
 * the reference's dense paths; here the matrix stays sparse and the constant
 * f(0) is carried implicitly through every later step. */
int m3b_normalize(m3b_mat* mat, const double* sf, int norm_method, double pseudo_count);
/* Normalised stored values in the input's CSC order (normalized_counts()@x). */
int m3b_export_values(m3b_mat* mat, double* y /* nnz_local */);

/* ---- gene selection, zero-gene filter, dual-layout build (K0) ---------- */
/* gene_order: optional list of n_sel gene indices (use_genes, in that order);
 * NULL = all genes.  keep[n_sel] receives the zero/non-finite row-sum filter
 * (1 = kept), n_kept the number of kept genes G'.  Collective across ranks. */
int m3b_select_genes(m3b_mat* mat, const int32_t* gene_order, int32_t n_sel,
                     uint8_t* keep, int32_t* n_kept);
/* The same selection and zero-gene filter for the PROJECTION path (preprocess_transform,
 * R/projection.R:117-133), without the two product layouts IRLBA needs: only m3b_project may
 * follow.  Gene orders that are not increasing, negative values and the dense pseudo-count paths
 * are handed to m3b_select_genes internally.  Collective across ranks. */
int m3b_select_genes_for_transform(m3b_mat* mat, const int32_t* gene_order, int32_t n_sel,
                                   uint8_t* keep, int32_t* n_kept);
/* Exact per-selected-gene bookkeeping after m3b_select_genes (local shard). */
int m3b_gene_nnz(m3b_mat* mat, int64_t* nnz_per_sel_gene /* n_sel */);

/* ---- gene moments (K3) -------------------------------------------------- */
int m3b_gene_moments(m3b_mat* mat, double* center /* G' */, double* scale /* G' */);

/* ---- TF-IDF (LSI branch: tfidf(), R/preprocess_cds.R:289-338; R/projection.R:190-216) ----- */
/* Call after m3b_normalize + m3b_select_genes: rewrites the prepared matrix in place as
 * log1p(FM / col_sums * scale_factor) * log(1 + num_cols / row_sums).  row_sums_in / num_cols:
 * NULL / 0 when building a model (computed here: rowSums(FM > 0) over all ranks, total cells);
 * the MODEL's values (per kept query gene) when projecting query cells.  col_sums_out
 * (n_cells_local) and row_sums_out (G') are optional.  Follow with m3b_pca_irlba(center=0,
 * scale=0) or m3b_project(center=NULL, scale=NULL). */
int m3b_tfidf(m3b_mat* mat, int frequencies, int log_scale_tf, double scale_factor,
              const double* row_sums_in, double num_cols, double* col_sums_out, double* row_sums_out);

/* ---- detect_genes / aggregate_gene_expression -------- */
/* detect_genes (R/utils.R:449-457): num_cells_expressed[G] = rowSums(counts > min_expr) over ALL
 * ranks' cells, num_genes_expressed[n_cells_local] = colSums(counts > min_expr).  Works on the raw
 * counts (any time after m3b_matrix_end).  Collective across ranks. */
int m3b_detect_genes(m3b_mat* mat, double min_expr, int64_t* num_cells_expressed, int64_t* num_genes_expressed);
/* The group sums of aggregate_gene_expression (R/cluster_genes.R:447-536) on the values of the
 * last m3b_normalize (normalized_counts): gene_group[G] = group id in [0, n_gene_groups) or -1
 * (gene not in any group); gene_agg_mean: 0 = colSums, 1 = colMeans over the member genes.
 * cell_group == NULL: out is n_gene_groups x n_cells_local (one column per local cell).
 * Otherwise cell_group[n_cells_local] = id in [0, n_cell_groups) or -1 and out is
 * n_gene_groups x n_cell_groups, summed (cell_agg_mean = 0) or averaged (1) over each group's
 * cells of ALL ranks (collective).  out is column-major with leading dimension ld_out. */
int m3b_aggregate_expression(m3b_mat* mat, const int32_t* gene_group, int32_t n_gene_groups, int gene_agg_mean,
                             const int32_t* cell_group, int32_t n_cell_groups, int cell_agg_mean,
                             double* out, int64_t ld_out);

/* ---- Jaccard weights of a kNN graph ------------------ */
/* Replaces the reference's native jaccard_coeff (src/clustering.cpp:13-58; registered as
 * _monocle3_jaccard_coeff, src/RcppExports.cpp; called at R/cluster_cells.R:313 and
 * R/graph_test.R:429,506), same argument and result layout: idx is an n x k matrix of 1-based
 * neighbour ids as doubles, column-major (an R NumericMatrix); out is (n*k) x 3, column-major:
 * (from, to, weight), row r = i*k + j, weights divided by their maximum.  k <= 2048 (the reference accepts any k; cluster_cells uses 20). */
int m3b_jaccard_coeff(m3b_ctx* ctx, const double* idx, int64_t n, int32_t k, int weight, double* out);
/* Exact k nearest neighbours (Euclidean) of every row of X among the rows of X: what
 * cluster_cells obtains from RANN::nn2(data, data, k, searchtype = "standard") with
 * nn_control method "nn2" (R/cluster_cells.R:286-287; the caller passes k + 1 because the point
 * itself is its own first neighbour).  X: n x d, column-major doubles (an R matrix, e.g.
 * reducedDims(cds)[["PCA"]]).  nn_idx: n x k, column-major, 1-based ids sorted by distance (ties:
 * smaller id first); nn_dists: n x k distances (not squared).  k <= 1024 and 512 d + 768 k + 49 664 <= 225 280 bytes of
 * shared memory per query block (d = 100: k <= 162; d = 50: k <= 195), else M3B_E_DIMS. */
int m3b_knn_self(m3b_ctx* ctx, const double* X, int64_t n, int32_t d, int32_t k, int32_t* nn_idx, double* nn_dists);

/* ---- align_cds residual regression ------------------- */
/* Replaces limma::lmFit(t(preproc_res), X.model_mat) and the subtraction that follows
 * (R/alignment.R:106-117): ordinary least squares of every column of X (cells x nv, the PCA or
 * LSI coordinates, column-major with leading dimension ldx, local cells) on the design M (cells x p,
 * column-major, leading dimension n, INTERCEPT FIRST -- what sparse.model.matrix returns for the
 * formula), with lm.fit's aliasing rule (collinear column -> NA -> 0).  beta receives
 * fit$coefficients[, -1] (nv x (p-1), column-major) and X is overwritten with
 * X - M[, -1] %*% t(beta).  Collective across ranks (the cells are sharded, beta is replicated).
 * p <= 64, p + nv <= 216. */
int m3b_align_residuals(m3b_ctx* ctx, double* X, int64_t n, int nv, int64_t ldx, const double* M, int p, double* beta);
/* align_beta_transform (R/projection.R:468-478): subtract the effects of a SAVED beta. */
int m3b_align_apply_beta(m3b_ctx* ctx, double* X, int64_t n, int nv, int64_t ldx, const double* M, int p, const double* beta);

/* ---- IRLBA PCA (K4-K9) -------------------------------------------------- */
/* nv: number of PCs; work: Lanczos subspace (0 => nv+7); v0: start vector (G').
 * center/scale: 0/1 flags (preprocess_cds passes scaling for both).
 * Outputs: d[nv]; V (G' x nv, ld G'); X = U*diag(d) for the local cells
 * (n_cells_local x nv, leading dimension ldx >= n_cells_local); center_out /
 * scale_out (G', nullable).  Collective across ranks. */
int m3b_pca_irlba(m3b_mat* mat, int nv, int work, int maxit, double tol, double svtol,
                  const double* v0, int center, int scale,
                  double* d, double* V, double* X, int64_t ldx,
                  double* center_out, double* scale_out, int* iter, int* mprod);
/* As above but leaves X on the device and returns timings only (bench). */
int m3b_pca_irlba_device(m3b_mat* mat, int nv, int work, int maxit, double tol, double svtol,
                         const double* v0, int center, int scale,
                         double* d, int* iter, int* mprod);

/* ---- projection of query cells (K10) ------------------------------------ */
/* Call after m3b_normalize + m3b_select_genes on the QUERY matrix.
 * model_row[G'_q]: for each kept query gene, its row in the model's V (>= 0).
 * V: n_model x nv (ld n_model); center/scale: n_model or NULL (scaling=FALSE).
 * Y: n_cells_local x nv, leading dimension ldy. */
int m3b_project(m3b_mat* query, const int32_t* model_row, int32_t n_model,
                const double* V, const double* center, const double* scale, int nv,
                double* Y, int64_t ldy);

/* ---- instrumentation ----------------------------------------------------- */
typedef struct m3b_stats {
  int64_t kernel_launches;     /* CUDA kernels launched by this context so far     */
  int64_t spmv_cells_out;      /* K4 launches                                      */
  int64_t spmv_genes_out;      /* K5 launches                                      */
  double  spmv_cells_out_ms;   /* summed CUDA-event time of K4 (when timing is on) */
  double  spmv_genes_out_ms;   /* summed CUDA-event time of K5                     */
  double  irlba_ms;            /* device time of the last m3b_pca_irlba            */
  double  prepare_ms;          /* normalize+select+moments device time             */
  int64_t spmv_bytes_cells_out;/* algorithmic bytes of one K4 launch               */
  int64_t spmv_bytes_genes_out;/* algorithmic bytes of one K5 launch               */
  int64_t h2d_bytes;           /* host->device bytes copied so far                 */
  int64_t d2h_bytes;           /* device->host bytes copied so far                 */
  /* with M3B_FLAG_TIME_SPMV: summed CUDA-event time of the IRLBA phases of the last run:
   * [0] K4 SpMV cells-out  [1] K5 SpMV genes-out  [2] K6 orthogonalisation (dots+apply)
   * [3] K8 small SVD + restart logic  [4] K9 restart GEMMs  [5] vector epilogues / combines
   * [6] allreduces  [7] reserved */
  double  phase_ms[8];
  int64_t svd_sweeps;          /* Jacobi sweeps summed over the restarts of the last run */
  /* host wall clock of the last m3b_pca_irlba: [0] setup+allocation [1] Lanczos loop
   * [2] final GEMMs + result copies [3] release */
  double  host_ms[4];
  int64_t svd_fallbacks;       /* restarts where the structured SVD gave up and Jacobi ran */
  double  k1_ms;               /* CUDA-event time of the last K1 (cell totals) kernel   */
  double  k2_ms;               /* CUDA-event time of the last K2 (normalize) kernel     */
  int64_t svd_fallback_reasons;/* OR of the reasons: 1 size, 2 irregular poles/weights, 4 root not converged, 8 Loewner,
                                  16 right vectors B^T u / s not unit (tiny singular value) */
  /* EXACT bytes one K4 / K5 launch moves to or from DRAM, counted from the layouts the kernel
   * streams: 10 B per (padded) general entry + 2 B per (padded) unit entry + the work-plan
   * descriptors + the dense operand (and the per-cell unit values on layout B) once + one 8-byte
   * partial per item.  The physical counterpart of spmv_bytes_* (the SURVEY 8d figure). */
  int64_t spmv_stream_bytes_cells_out;
  int64_t spmv_stream_bytes_genes_out;
  /* host wall clock of the phases of the last m3b_select_genes (ms): [0] layout B build
   * [1] row sums + keep mask [2] re-map + plans of B [3] layout A build + plans */
  double  select_ms[4];
  double  k10_ms;              /* CUDA-event time of the last K10 (projection) kernel(s)  */
  /* cell-sharded runs, last m3b_pca_irlba: what the in-kernel cross-rank exchanges of the fused
   * vector kernel cost, measured with clock64 on CTA 0 and summed over the launches (us at the
   * nominal SM clock): [0] wait for the peers' gene-vector rows  [1] wait for the peers'
   * Gram-Schmidt coefficients  [2] wait for the world's norm partials  [3] whole kernel, gene side
   * [4] whole kernel, cell side  [5] launches, gene side  [6] launches, cell side  [7] clock (kHz) */
  double  xchg_diag[8];
} m3b_stats;
int m3b_get_stats(const m3b_ctx* ctx, m3b_stats* out);
/* Measured fp64 FMA throughput of the device (TFLOP/s): the roof K10 (projection) is reported
 * against; SURVEY 8(d) notes it is not among the driver's measured peaks. */
int m3b_measure_fp64_peak(m3b_ctx* ctx, double* tflops);
int m3b_reset_stats(m3b_ctx* ctx);
/* flags for m3b_create / m3b_set_flags */
#define M3B_FLAG_TIME_SPMV 1  /* bracket every SpMV launch with CUDA events (read back after the run) */
#define M3B_FLAG_JACOBI_ONLY 2 /* small SVD: always use the dense Jacobi kernel (no structured fast path) */
#define M3B_FLAG_NO_UNIT_SPLIT 4 /* m3b_select_genes: keep count-1 entries in the 10 B/entry stream instead of the
                                    2 B/entry index-only "unit" stream (A/B switch for measurements) */
int m3b_set_flags(m3b_ctx* ctx, int flags);
/* CUDA-event stopwatch on the context's stream (the stream every kernel of this
 * library is launched on): start records an event, stop records a second one,
 * synchronises and returns the elapsed device time in milliseconds. */
int m3b_timer_start(m3b_ctx* ctx);
int m3b_timer_stop(m3b_ctx* ctx, double* elapsed_ms);

#ifdef __cplusplus
}
#endif
#endif /* M3B_H */
