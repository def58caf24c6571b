// Rcpp layer of the drop-in: marshaling only, no arithmetic.  NOT compiled in this image (no R);
// it mirrors the export/registration pattern of monocle3's src/RcppExports.cpp:15-47 and binds
// the extern "C" entry points of include/m3b.h.  Add to monocle3/src/ with
//   PKG_LIBS = -L$(M3B_HOME)/monocle3_b200 -lm3b      (src/Makevars)
//
// Error handling: the device matrix lives in a guard object whose destructor frees it, so a
// Rcpp::stop() raised by m3b_check() (a C++ exception, turned into an R error by END_RCPP) never
// leaks device memory.  R's random stream and interrupt check reach the library through the two
// optional callbacks of the ABI (m3b_set_rnorm, m3b_set_should_abort).
#include <Rcpp.h>
#include <R_ext/Lapack.h>
#include <R_ext/Random.h>
#include <R_ext/Rdynload.h>
#include <R_ext/Utils.h>
extern "C" {
#include "m3b.h"
}
using namespace Rcpp;

static void m3b_check(m3b_ctx* ctx, int st) {
  if (st == M3B_OK) return;
  if (st == M3B_W_NOT_CONVERGED) {  // irlba: warning, results still returned
    Rcpp::warning("did not converge--results might be invalid!; try increasing work or maxit");
    return;
  }
  if (st == M3B_E_ABORTED) Rcpp::stop("interrupted");
  Rcpp::stop("m3b: %s (%s)", m3b_last_error(ctx), m3b_status_string(st));
}

// ---- callbacks -------------------------------------------------------------------------------
// R's own LAPACK for the small SVD => signs identical to stock monocle3/irlba
static int r_dgesdd(int n, const double* B, double* U, double* s, double* Vt, void*) {
  std::vector<double> A(B, B + (size_t)n * n), work(1);
  std::vector<int> iwork(8 * n);
  int lwork = -1, info = 0;
  F77_CALL(dgesdd)("A", &n, &n, A.data(), &n, s, U, &n, Vt, &n, work.data(), &lwork, iwork.data(), &info FCONE);
  lwork = (int)work[0];
  work.resize(lwork);
  F77_CALL(dgesdd)("A", &n, &n, A.data(), &n, s, U, &n, Vt, &n, work.data(), &lwork, iwork.data(), &info FCONE);
  return info;
}

// irlba replaces a Lanczos vector by rnorm() when it hits an invariant subspace (irlb.c): the draws
// must come from R's stream, which set.seed(2016) positioned (R/preprocess_cds.R:110)
static int r_rnorm(int64_t n, double* out, void*) {
  GetRNGstate();
  for (int64_t k = 0; k < n; ++k) out[k] = norm_rand();
  PutRNGstate();
  return 0;
}

// irlba polls R_CheckUserInterrupt after every product; the library polls this once per restart.
// R_CheckUserInterrupt longjmps on an interrupt, which must not cross the ABI: run it under
// R_ToplevelExec and report the outcome as a return value.
static void check_interrupt_fn(void*) { R_CheckUserInterrupt(); }
static int r_should_abort(void*) { return R_ToplevelExec(check_interrupt_fn, nullptr) == FALSE; }

static m3b_ctx* the_ctx() {
  static m3b_ctx* ctx = nullptr;
  if (!ctx) {
    int st = m3b_create(0, 0, &ctx);
    if (st != M3B_OK) Rcpp::stop("m3b: %s", m3b_status_string(st));  // no CPU fallback
    m3b_set_small_svd(ctx, r_dgesdd, nullptr);
    m3b_set_rnorm(ctx, r_rnorm, nullptr);
    m3b_set_should_abort(ctx, r_should_abort, nullptr);
  }
  return ctx;
}

// ---- the counts matrix on the device, freed on every exit path ---------------------------------
struct DeviceCounts {
  m3b_ctx* ctx;
  m3b_mat* m = nullptr;
  int n_genes = 0, n_cells = 0;
  explicit DeviceCounts(S4 counts) : ctx(the_ctx()) {  // dgCMatrix -> device
    IntegerVector p = counts.slot("p"), i = counts.slot("i"), dim = counts.slot("Dim");
    NumericVector x = counts.slot("x");
    n_genes = dim[0];
    n_cells = dim[1];
    m3b_check(ctx, m3b_matrix_begin(ctx, dim[0], dim[1], dim[1], x.size(), &m));
    m3b_check(ctx, m3b_matrix_append_csc(m, dim[1], p.begin(), i.begin(), x.begin()));
    m3b_check(ctx, m3b_matrix_end(m));
  }
  // A matrix beyond one dgCMatrix (2^31 stored entries): consecutive CELL BLOCKS, each a dgCMatrix with
  // the same genes (what cbind() of them would be, if it could exist).  Appended in order, exactly
  // like the chunks of one matrix; the Python mirror's monocle3_b200.CSCBlocks is the same thing.
  explicit DeviceCounts(List blocks) : ctx(the_ctx()) {
    if (blocks.size() == 0) Rcpp::stop("m3b: empty list of count blocks");
    double nnz = 0;
    long long cells = 0;
    for (int b = 0; b < blocks.size(); ++b) {
      S4 blk = blocks[b];
      IntegerVector dim = blk.slot("Dim");
      if (b == 0) n_genes = dim[0];
      if (dim[0] != n_genes) Rcpp::stop("m3b: count blocks differ in their number of genes");
      cells += dim[1];
      nnz += (double)NumericVector(blk.slot("x")).size();
    }
    n_cells = (int)cells;  // (R-side bookkeeping only; the library counts cells in 64 bits)
    m3b_check(ctx, m3b_matrix_begin(ctx, n_genes, cells, cells, (int64_t)nnz, &m));
    for (int b = 0; b < blocks.size(); ++b) {
      S4 blk = blocks[b];
      IntegerVector p = blk.slot("p"), i = blk.slot("i"), dim = blk.slot("Dim");
      NumericVector x = blk.slot("x");
      const int st = m3b_matrix_append_csc(m, dim[1], p.begin(), i.begin(), x.begin());
      if (st != M3B_OK) {
        m3b_matrix_free(m);  // free before raising (the destructor does not run when the constructor throws)
        m = nullptr;
        m3b_check(ctx, st);
      }
    }
    m3b_check(ctx, m3b_matrix_end(m));
  }
  ~DeviceCounts() { m3b_matrix_free(m); }
  DeviceCounts(const DeviceCounts&) = delete;
  DeviceCounts& operator=(const DeviceCounts&) = delete;
};

// estimate_sf_sparse, R/utils.R:54-70
// [[Rcpp::export]]
NumericVector m3b_estimate_sf_sparse(S4 counts, bool round_exprs, int method) {
  DeviceCounts D(counts);
  NumericVector tot(D.n_cells), sf(D.n_cells);
  int any_zero = 0;
  m3b_check(D.ctx, m3b_size_factors(D.m, round_exprs, method, tot.begin(), sf.begin(), &any_zero));
  return sf;
}

// normalized_counts, R/utils.R:477-509 (sparse branch): the @x slot of the result; the caller
// re-uses @i / @p of the input.  norm_method: M3B_NORM_LOG10 / SIZE_ONLY / BINARY
// [[Rcpp::export]]
NumericVector m3b_normalized_counts_x(S4 counts, NumericVector sf, int norm_method, double pseudocount) {
  DeviceCounts D(counts);
  NumericVector x = counts.slot("x");
  NumericVector y(x.size());
  m3b_check(D.ctx, m3b_normalize(D.m, sf.begin(), norm_method, pseudocount));
  m3b_check(D.ctx, m3b_export_values(D.m, y.begin()));
  return y;
}

static int select(DeviceCounts& D, Nullable<IntegerVector> gene_order, RawVector& keep, bool for_transform = false) {
  int n_sel = D.n_genes, n_kept = 0;
  const int* go = nullptr;
  IntegerVector gov;
  if (gene_order.isNotNull()) {
    gov = IntegerVector(gene_order);
    go = gov.begin();
    n_sel = gov.size();
  }
  keep = RawVector(n_sel);
  m3b_check(D.ctx, (for_transform ? m3b_select_genes_for_transform : m3b_select_genes)(D.m, go, n_sel, (uint8_t*)keep.begin(), &n_kept));
  return n_kept;
}

// PCA branch of preprocess_cds, R/preprocess_cds.R:111-160
// [[Rcpp::export]]
List m3b_preprocess_pca(S4 counts, NumericVector sf, int norm_method, double pseudo_count,
                        Nullable<IntegerVector> gene_order, int num_dim, bool scaling, NumericVector v0_all) {
  DeviceCounts D(counts);
  m3b_check(D.ctx, m3b_normalize(D.m, sf.begin(), norm_method, pseudo_count));
  RawVector keep;
  const int n_kept = select(D, gene_order, keep);
  const int nv = std::min(num_dim, std::min(n_kept, D.n_cells) - 1);  // R/preprocess_cds.R:140
  NumericVector d(nv), cen(n_kept), sc(n_kept);
  NumericMatrix V(n_kept, nv), X(D.n_cells, nv);
  int iter = 0, mprod = 0;
  // v0_all: rnorm(G) drawn in R after set.seed(2016); irlba draws rnorm(G'), i.e. its first n_kept values
  m3b_check(D.ctx, m3b_pca_irlba(D.m, nv, 0, 1000, 1e-5, 1e-5, v0_all.begin(), scaling, scaling, d.begin(), V.begin(),
                                 X.begin(), D.n_cells, cen.begin(), sc.begin(), &iter, &mprod));
  return List::create(_["d"] = d, _["v"] = V, _["x"] = X, _["center"] = cen, _["scale"] = sc,
                      _["keep"] = keep, _["iter"] = iter, _["mprod"] = mprod);
}

// LSI branch of preprocess_cds, R/preprocess_cds.R:185-241: tfidf() then irlba without centre / scale
// [[Rcpp::export]]
List m3b_preprocess_lsi(S4 counts, NumericVector sf, int norm_method, double pseudo_count,
                        Nullable<IntegerVector> gene_order, int num_dim, NumericVector v0_all) {
  DeviceCounts D(counts);
  m3b_check(D.ctx, m3b_normalize(D.m, sf.begin(), norm_method, pseudo_count));
  RawVector keep;
  const int n_kept = select(D, gene_order, keep);
  NumericVector col_sums(D.n_cells), row_sums(n_kept);
  m3b_check(D.ctx, m3b_tfidf(D.m, /*frequencies*/ 1, /*log_scale_tf*/ 1, 100000.0, nullptr, 0.0, col_sums.begin(), row_sums.begin()));
  const int nv = std::min(num_dim, std::min(n_kept, D.n_cells) - 1);  // :194
  NumericVector d(nv);
  NumericMatrix V(n_kept, nv), X(D.n_cells, nv);
  int iter = 0, mprod = 0;
  m3b_check(D.ctx, m3b_pca_irlba(D.m, nv, 0, 1000, 1e-5, 1e-5, v0_all.begin(), 0, 0, d.begin(), V.begin(), X.begin(),
                                 D.n_cells, nullptr, nullptr, &iter, &mprod));
  return List::create(_["d"] = d, _["v"] = V, _["x"] = X, _["col_sums"] = col_sums, _["row_sums"] = row_sums,
                      _["num_cols"] = D.n_cells, _["keep"] = keep, _["iter"] = iter, _["mprod"] = mprod);
}

// preprocess_transform, R/projection.R:102-149 (PCA) and :156-241 (LSI: row_sums / num_cols of the MODEL)
// [[Rcpp::export]]
NumericMatrix m3b_preprocess_transform(S4 counts, NumericVector sf, int norm_method, double pseudo_count,
                                       IntegerVector model_row_of_gene /* per query gene, -1 = absent */,
                                       NumericMatrix V, Nullable<NumericVector> center, Nullable<NumericVector> scale,
                                       Nullable<NumericVector> lsi_row_sums, double lsi_num_cols) {
  DeviceCounts D(counts);
  const bool lsi = lsi_row_sums.isNotNull();
  m3b_check(D.ctx, m3b_normalize(D.m, sf.begin(), norm_method, pseudo_count));
  RawVector keep;
  int n_kept = select(D, R_NilValue, keep, /*for_transform=*/!lsi);
  // intersect(rownames(rotation), rownames(FM)) (R/projection.R:121): drop query genes the model lacks
  std::vector<int> sel, rows;
  for (int g = 0; g < D.n_genes; ++g)
    if (keep[g] && model_row_of_gene[g] >= 0) {
      sel.push_back(g);
      rows.push_back(model_row_of_gene[g]);
    }
  if ((int)sel.size() != n_kept)
    m3b_check(D.ctx, (lsi ? m3b_select_genes : m3b_select_genes_for_transform)(D.m, sel.data(), (int)sel.size(), nullptr, &n_kept));
  NumericMatrix Y(D.n_cells, V.ncol());
  NumericVector c, s;
  const double *cp = nullptr, *sp = nullptr;
  if (lsi) {
    NumericVector rs_model(lsi_row_sums), rs(rows.size());
    for (size_t k = 0; k < rows.size(); ++k) rs[k] = rs_model[rows[k]];
    m3b_check(D.ctx, m3b_tfidf(D.m, 1, 1, 100000.0, rs.begin(), lsi_num_cols, nullptr, nullptr));
  } else {
    if (center.isNotNull()) { c = NumericVector(center); cp = c.begin(); }
    if (scale.isNotNull()) { s = NumericVector(scale); sp = s.begin(); }
  }
  m3b_check(D.ctx, m3b_project(D.m, rows.data(), V.nrow(), V.begin(), cp, sp, V.ncol(), Y.begin(), D.n_cells));
  return Y;
}

// align_cds residual regression, R/alignment.R:106-117 (design = sparse.model.matrix(...), densified: cells x p)
// [[Rcpp::export]]
List m3b_align_residuals_(NumericMatrix preproc_res, NumericMatrix design) {
  NumericMatrix X = clone(preproc_res);
  NumericMatrix beta(X.ncol(), std::max(design.ncol() - 1, 0));
  std::vector<double> dummy(1);
  m3b_check(the_ctx(), m3b_align_residuals(the_ctx(), X.begin(), X.nrow(), X.ncol(), X.nrow(), design.begin(), design.ncol(),
                                           beta.size() ? beta.begin() : dummy.data()));
  return List::create(_["aligned"] = X, _["beta"] = beta);
}

// align_beta_transform, R/projection.R:468-478
// [[Rcpp::export]]
NumericMatrix m3b_align_apply_beta_(NumericMatrix preproc_res, NumericMatrix design, NumericMatrix beta) {
  NumericMatrix X = clone(preproc_res);
  m3b_check(the_ctx(), m3b_align_apply_beta(the_ctx(), X.begin(), X.nrow(), X.ncol(), X.nrow(), design.begin(), design.ncol(),
                                            beta.begin()));
  return X;
}

// detect_genes (R/utils.R:449-457): list(num_cells_expressed, num_genes_expressed)
// [[Rcpp::export]]
List m3b_detect_genes_(S4 counts, double min_expr) {
  DeviceCounts D(counts);
  std::vector<int64_t> per_gene(D.n_genes), per_cell(D.n_cells);
  m3b_check(D.ctx, m3b_detect_genes(D.m, min_expr, per_gene.data(), per_cell.data()));
  return List::create(_["num_cells_expressed"] = NumericVector(per_gene.begin(), per_gene.end()),
                      _["num_genes_expressed"] = NumericVector(per_cell.begin(), per_cell.end()));
}

// group sums behind aggregate_gene_expression (R/cluster_genes.R:486-522).  gene_group / cell_group are
// 0-based ids (-1 = not in a group; the R side builds them from the data frames: unique() order for gene
// groups, factor levels for cell groups); the scale()/clamp step (:526-536) stays in R.
// [[Rcpp::export]]
NumericMatrix m3b_aggregate_(S4 counts, NumericVector sf, int norm_method, IntegerVector gene_group, int n_gene_groups,
                             bool gene_mean, IntegerVector cell_group, int n_cell_groups, bool cell_mean) {
  DeviceCounts D(counts);
  m3b_check(D.ctx, m3b_normalize(D.m, sf.begin(), norm_method, norm_method == M3B_NORM_LOG10 ? 1.0 : 0.0));
  NumericMatrix out(n_gene_groups, n_cell_groups);
  m3b_check(D.ctx, m3b_aggregate_expression(D.m, gene_group.begin(), n_gene_groups, gene_mean, cell_group.begin(), n_cell_groups,
                                            cell_mean, out.begin(), n_gene_groups));
  return out;
}

// exact kNN for cluster_cells_make_graph's "nn2" method (R/cluster_cells.R:286-287): the result list of
// RANN::nn2(data, data, k, searchtype = "standard")
// [[Rcpp::export]]
List m3b_nn2_self(NumericMatrix data, int k) {
  IntegerMatrix idx(data.nrow(), k);
  NumericMatrix dists(data.nrow(), k);
  m3b_check(the_ctx(), m3b_knn_self(the_ctx(), data.begin(), data.nrow(), data.ncol(), k, idx.begin(), dists.begin()));
  return List::create(_["nn.idx"] = idx, _["nn.dists"] = dists);
}

// drop-in body for the package's native kernel (src/clustering.cpp:49-54)
// [[Rcpp::export]]
NumericMatrix jaccard_coeff(SEXP R_idx, SEXP R_weight) {
  NumericMatrix idx(R_idx);
  NumericMatrix weights(idx.nrow() * idx.ncol(), 3);
  m3b_check(the_ctx(), m3b_jaccard_coeff(the_ctx(), idx.begin(), idx.nrow(), idx.ncol(), as<bool>(R_weight), weights.begin()));
  return weights;
}

// ---- registration, as Rcpp::compileAttributes() generates it (src/RcppExports.cpp:15-47) -------
#define M3B_WRAP(NAME, ...)                                   \
  RcppExport SEXP _monocle3_##NAME(__VA_ARGS__)
#define M3B_BEGIN BEGIN_RCPP Rcpp::RObject rcpp_result_gen; Rcpp::RNGScope rcpp_rngScope_gen;
#define M3B_END return rcpp_result_gen; END_RCPP

M3B_WRAP(m3b_estimate_sf_sparse, SEXP countsSEXP, SEXP roundSEXP, SEXP methodSEXP) {
  M3B_BEGIN
  rcpp_result_gen = Rcpp::wrap(m3b_estimate_sf_sparse(as<S4>(countsSEXP), as<bool>(roundSEXP), as<int>(methodSEXP)));
  M3B_END
}
M3B_WRAP(m3b_normalized_counts_x, SEXP countsSEXP, SEXP sfSEXP, SEXP nmSEXP, SEXP pcSEXP) {
  M3B_BEGIN
  rcpp_result_gen = Rcpp::wrap(m3b_normalized_counts_x(as<S4>(countsSEXP), as<NumericVector>(sfSEXP), as<int>(nmSEXP), as<double>(pcSEXP)));
  M3B_END
}
M3B_WRAP(m3b_preprocess_pca, SEXP a, SEXP b, SEXP c, SEXP d, SEXP e, SEXP f, SEXP g, SEXP h) {
  M3B_BEGIN
  rcpp_result_gen = Rcpp::wrap(m3b_preprocess_pca(as<S4>(a), as<NumericVector>(b), as<int>(c), as<double>(d),
                                                  as<Nullable<IntegerVector>>(e), as<int>(f), as<bool>(g), as<NumericVector>(h)));
  M3B_END
}
M3B_WRAP(m3b_preprocess_lsi, SEXP a, SEXP b, SEXP c, SEXP d, SEXP e, SEXP f, SEXP g) {
  M3B_BEGIN
  rcpp_result_gen = Rcpp::wrap(m3b_preprocess_lsi(as<S4>(a), as<NumericVector>(b), as<int>(c), as<double>(d),
                                                  as<Nullable<IntegerVector>>(e), as<int>(f), as<NumericVector>(g)));
  M3B_END
}
M3B_WRAP(m3b_preprocess_transform, SEXP a, SEXP b, SEXP c, SEXP d, SEXP e, SEXP f, SEXP g, SEXP h, SEXP i, SEXP j) {
  M3B_BEGIN
  rcpp_result_gen = Rcpp::wrap(m3b_preprocess_transform(as<S4>(a), as<NumericVector>(b), as<int>(c), as<double>(d), as<IntegerVector>(e),
                                                        as<NumericMatrix>(f), as<Nullable<NumericVector>>(g),
                                                        as<Nullable<NumericVector>>(h), as<Nullable<NumericVector>>(i), as<double>(j)));
  M3B_END
}
M3B_WRAP(m3b_align_residuals_, SEXP a, SEXP b) {
  M3B_BEGIN
  rcpp_result_gen = Rcpp::wrap(m3b_align_residuals_(as<NumericMatrix>(a), as<NumericMatrix>(b)));
  M3B_END
}
M3B_WRAP(m3b_align_apply_beta_, SEXP a, SEXP b, SEXP c) {
  M3B_BEGIN
  rcpp_result_gen = Rcpp::wrap(m3b_align_apply_beta_(as<NumericMatrix>(a), as<NumericMatrix>(b), as<NumericMatrix>(c)));
  M3B_END
}
M3B_WRAP(m3b_detect_genes_, SEXP a, SEXP b) {
  M3B_BEGIN
  rcpp_result_gen = Rcpp::wrap(m3b_detect_genes_(as<S4>(a), as<double>(b)));
  M3B_END
}
M3B_WRAP(m3b_aggregate_, SEXP a, SEXP b, SEXP c, SEXP d, SEXP e, SEXP f, SEXP g, SEXP h, SEXP i) {
  M3B_BEGIN
  rcpp_result_gen = Rcpp::wrap(m3b_aggregate_(as<S4>(a), as<NumericVector>(b), as<int>(c), as<IntegerVector>(d), as<int>(e), as<bool>(f),
                                              as<IntegerVector>(g), as<int>(h), as<bool>(i)));
  M3B_END
}
M3B_WRAP(m3b_nn2_self, SEXP a, SEXP b) {
  M3B_BEGIN
  rcpp_result_gen = Rcpp::wrap(m3b_nn2_self(as<NumericMatrix>(a), as<int>(b)));
  M3B_END
}
M3B_WRAP(jaccard_coeff, SEXP R_idxSEXP, SEXP R_weightSEXP) {  // same symbol as src/RcppExports.cpp:15-23
  M3B_BEGIN
  rcpp_result_gen = Rcpp::wrap(jaccard_coeff(R_idxSEXP, R_weightSEXP));
  M3B_END
}

static const R_CallMethodDef CallEntries[] = {
    {"_monocle3_jaccard_coeff", (DL_FUNC)&_monocle3_jaccard_coeff, 2},
    {"_monocle3_m3b_estimate_sf_sparse", (DL_FUNC)&_monocle3_m3b_estimate_sf_sparse, 3},
    {"_monocle3_m3b_normalized_counts_x", (DL_FUNC)&_monocle3_m3b_normalized_counts_x, 4},
    {"_monocle3_m3b_preprocess_pca", (DL_FUNC)&_monocle3_m3b_preprocess_pca, 8},
    {"_monocle3_m3b_preprocess_lsi", (DL_FUNC)&_monocle3_m3b_preprocess_lsi, 7},
    {"_monocle3_m3b_preprocess_transform", (DL_FUNC)&_monocle3_m3b_preprocess_transform, 10},
    {"_monocle3_m3b_align_residuals_", (DL_FUNC)&_monocle3_m3b_align_residuals_, 2},
    {"_monocle3_m3b_align_apply_beta_", (DL_FUNC)&_monocle3_m3b_align_apply_beta_, 3},
    {"_monocle3_m3b_detect_genes_", (DL_FUNC)&_monocle3_m3b_detect_genes_, 2},
    {"_monocle3_m3b_aggregate_", (DL_FUNC)&_monocle3_m3b_aggregate_, 9},
    {"_monocle3_m3b_nn2_self", (DL_FUNC)&_monocle3_m3b_nn2_self, 2},
    {NULL, NULL, 0}};

// (the stock package registers jaccard_coeff and pnorm_over_mat here; the latter is untouched)
RcppExport void R_init_monocle3(DllInfo* dll) {
  R_registerRoutines(dll, NULL, CallEntries, NULL, NULL);
  R_useDynamicSymbols(dll, FALSE);
}
