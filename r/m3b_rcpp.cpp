// Rcpp layer of the drop-in: marshaling only, no arithmetic.  NOT compiled in this image (no R);
// it mirrors the export/registration pattern of monocle3's src/RcppExports.cpp:15-47 and binds
// the extern "C" entry points of include/m3b.h.  Add to monocle3/src/ with
//   PKG_LIBS = -L$(M3B_HOME)/monocle3_b200 -lm3b      (src/Makevars)
#include <Rcpp.h>
#include <R_ext/Lapack.h>
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
  Rcpp::stop("m3b: %s (%s)", m3b_last_error(ctx), m3b_status_string(st));
}

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

static m3b_ctx* the_ctx() {
  static m3b_ctx* ctx = nullptr;
  if (!ctx) {
    int st = m3b_create(0, 0, &ctx);
    if (st != M3B_OK) Rcpp::stop("m3b: %s", m3b_status_string(st));  // no CPU fallback
    m3b_set_small_svd(ctx, r_dgesdd, nullptr);
  }
  return ctx;
}

static m3b_mat* upload(S4 counts) {  // dgCMatrix -> device
  IntegerVector p = counts.slot("p"), i = counts.slot("i"), dim = counts.slot("Dim");
  NumericVector x = counts.slot("x");
  m3b_ctx* ctx = the_ctx();
  m3b_mat* m = nullptr;
  m3b_check(ctx, m3b_matrix_begin(ctx, dim[0], dim[1], dim[1], x.size(), &m));
  m3b_check(ctx, m3b_matrix_append_csc(m, dim[1], p.begin(), i.begin(), x.begin()));
  m3b_check(ctx, m3b_matrix_end(m));
  return m;
}

// [[Rcpp::export]]
NumericVector m3b_estimate_sf_sparse(S4 counts, bool round_exprs, int method) {
  m3b_mat* m = upload(counts);
  IntegerVector dim = counts.slot("Dim");
  NumericVector tot(dim[1]), sf(dim[1]);
  int any_zero = 0;
  int st = m3b_size_factors(m, round_exprs, method, tot.begin(), sf.begin(), &any_zero);
  m3b_matrix_free(m);
  m3b_check(the_ctx(), st);
  return sf;
}

// [[Rcpp::export]]
List m3b_preprocess_pca(S4 counts, NumericVector sf, int norm_method, double pseudo_count,
                        Nullable<IntegerVector> gene_order, int num_dim, bool scaling, NumericVector v0_all) {
  m3b_ctx* ctx = the_ctx();
  m3b_mat* m = upload(counts);
  IntegerVector dim = counts.slot("Dim");
  m3b_check(ctx, m3b_normalize(m, sf.begin(), norm_method, pseudo_count));
  int n_sel = dim[0];
  const int* go = nullptr;
  IntegerVector gov;
  if (gene_order.isNotNull()) { gov = IntegerVector(gene_order); go = gov.begin(); n_sel = gov.size(); }
  RawVector keep(n_sel);
  int n_kept = 0;
  m3b_check(ctx, m3b_select_genes(m, go, n_sel, (uint8_t*)keep.begin(), &n_kept));
  int nv = std::min(num_dim, std::min(n_kept, dim[1]) - 1);          // R/preprocess_cds.R:140
  NumericVector d(nv), cen(n_kept), sc(n_kept);
  NumericMatrix V(n_kept, nv), X(dim[1], nv);
  int iter = 0, mprod = 0;
  // v0_all: rnorm(G') drawn in R after set.seed(2016) -- the shim passes the first n_kept values
  int st = m3b_pca_irlba(m, nv, 0, 1000, 1e-5, 1e-5, v0_all.begin(), scaling, scaling, d.begin(), V.begin(),
                         X.begin(), dim[1], cen.begin(), sc.begin(), &iter, &mprod);
  m3b_matrix_free(m);
  m3b_check(ctx, st);
  return List::create(_["d"] = d, _["v"] = V, _["x"] = X, _["center"] = cen, _["scale"] = sc,
                      _["keep"] = keep, _["iter"] = iter, _["mprod"] = mprod);
}

// [[Rcpp::export]]
NumericMatrix m3b_preprocess_transform(S4 counts, NumericVector sf, int norm_method, double pseudo_count,
                                       IntegerVector model_row_of_gene /* per query gene, -1 = absent */,
                                       NumericMatrix V, Nullable<NumericVector> center, Nullable<NumericVector> scale) {
  m3b_ctx* ctx = the_ctx();
  m3b_mat* m = upload(counts);
  IntegerVector dim = counts.slot("Dim");
  m3b_check(ctx, m3b_normalize(m, sf.begin(), norm_method, pseudo_count));
  RawVector keep(dim[0]);
  int n_kept = 0;
  m3b_check(ctx, m3b_select_genes(m, nullptr, dim[0], (uint8_t*)keep.begin(), &n_kept));
  // intersect(rownames(rotation), rownames(FM)) (R/projection.R:121): drop query genes the model lacks
  std::vector<int> sel, rows;
  for (int g = 0; g < dim[0]; ++g)
    if (keep[g] && model_row_of_gene[g] >= 0) { sel.push_back(g); rows.push_back(model_row_of_gene[g]); }
  if ((int)sel.size() != n_kept)
    m3b_check(ctx, m3b_select_genes(m, sel.data(), (int)sel.size(), nullptr, &n_kept));
  NumericMatrix Y(dim[1], V.ncol());
  NumericVector c, s;
  const double *cp = nullptr, *sp = nullptr;
  if (center.isNotNull()) { c = NumericVector(center); cp = c.begin(); }
  if (scale.isNotNull()) { s = NumericVector(scale); sp = s.begin(); }
  int st = m3b_project(m, rows.data(), V.nrow(), V.begin(), cp, sp, V.ncol(), Y.begin(), dim[1]);
  m3b_matrix_free(m);
  m3b_check(ctx, st);
  return Y;
}

// detect_genes (R/utils.R:449-457): list(num_cells_expressed, num_genes_expressed)
// [[Rcpp::export]]
List m3b_detect_genes_(S4 counts, double min_expr) {
  m3b_mat* m = upload(counts);
  IntegerVector dim = counts.slot("Dim");
  std::vector<int64_t> per_gene(dim[0]), per_cell(dim[1]);
  int st = m3b_detect_genes(m, min_expr, per_gene.data(), per_cell.data());
  m3b_matrix_free(m);
  m3b_check(the_ctx(), st);
  return List::create(_["num_cells_expressed"] = NumericVector(per_gene.begin(), per_gene.end()),
                      _["num_genes_expressed"] = NumericVector(per_cell.begin(), per_cell.end()));
}

// group sums behind aggregate_gene_expression (R/cluster_genes.R:486-522).  gene_group / cell_group are
// 0-based ids (-1 = not in a group; the R side builds them from the data frames: unique() order for gene
// groups, factor levels for cell groups); the scale()/clamp step (:526-536) stays in R.
// [[Rcpp::export]]
NumericMatrix m3b_aggregate_(S4 counts, NumericVector sf, int norm_method, IntegerVector gene_group, int n_gene_groups,
                             bool gene_mean, IntegerVector cell_group, int n_cell_groups, bool cell_mean) {
  m3b_ctx* ctx = the_ctx();
  m3b_mat* m = upload(counts);
  m3b_check(ctx, m3b_normalize(m, sf.begin(), norm_method, norm_method == M3B_NORM_LOG10 ? 1.0 : 0.0));
  NumericMatrix out(n_gene_groups, n_cell_groups);
  int st = m3b_aggregate_expression(m, gene_group.begin(), n_gene_groups, gene_mean, cell_group.begin(), n_cell_groups,
                                    cell_mean, out.begin(), n_gene_groups);
  m3b_matrix_free(m);
  m3b_check(ctx, st);
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

// registration, as in src/RcppExports.cpp:38-47 (generated by Rcpp::compileAttributes() in practice)
