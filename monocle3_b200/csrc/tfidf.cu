// TF-IDF for the LSI branch of preprocess_cds.
//
// Replaces tfidf() (R/preprocess_cds.R:289-338) and its copy inside preprocess_transform
// (R/projection.R:190-216):
//   col_sums = colSums(FM);  tf = log1p(FM / col_sums * scale_factor);
//   row_sums = rowSums(FM > 0);  idf = log(1 + num_cols / row_sums);  tf_idf = tf * idf
// Both tiled layouts already hold the normalised, gene-filtered FM, so TF-IDF is an in-place
// element-wise rewrite of their value arrays: every item knows its row, the 16-bit index gives
// the other coordinate.  Column sums come from layout A (rows = cells), the positive counts per
// gene from layout B, with the same streaming kernel as the Lanczos products (OP_SUM /
// OP_COUNTPOS).  The idf vector (G' logs) is finished on the host with libm.
#include <cmath>
#include <vector>

#include "m3b_internal.h"

// ROWS_ARE_CELLS = true  (layout A): val = f(val / colsum[row]) * idf[tile_off + idx]
// ROWS_ARE_CELLS = false (layout B): val = f(val / colsum[tile_off + idx]) * idf[row]
template <bool ROWS_ARE_CELLS>
__global__ void __launch_bounds__(256) k_tfidf_items(unsigned short* __restrict__ idx, double* __restrict__ val,
                                                     const ItemDesc* __restrict__ items, long long n_items,
                                                     const int* __restrict__ item_tile, int tile_w,
                                                     const double* __restrict__ colsum, const double* __restrict__ idf,
                                                     int frequencies, int log_scale_tf, double scale_factor) {
  const int lane = threadIdx.x & 31;
  long long warp = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long it = warp; it < n_items; it += nwarps) {
    const ItemDesc d = items[it];
    const long long off = (long long)item_tile[it] * tile_w;
    for (int e = lane; e < d.len; e += 32) {
      const double v = val[d.start + e];
      if (v == 0.0) continue;  // pads and explicit zeros stay zero (log1p(0) * idf = 0)
      const long long other = off + (idx[d.start + e] & 0x7fff);  // bit 15: head flag (flat.cu)
      double t = v;
      if (frequencies) t = t / (ROWS_ARE_CELLS ? colsum[d.row] : colsum[other]);
      if (log_scale_tf) t = log1p(t * (frequencies ? scale_factor : 1.0));
      val[d.start + e] = t * (ROWS_ARE_CELLS ? idf[other] : idf[d.row]);
    }
  }
}

static int item_tiles(m3b_ctx* ctx, const TiledLayout& L, int** out) {
  // tile of every item, from the (tile,row) -> item ranges
  std::vector<int> rip((size_t)L.n_tiles * (L.rows + 1));
  CK(ctx, cudaMemcpyAsync(rip.data(), L.row_item_ptr, rip.size() * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  CK(ctx, cudaStreamSynchronize(ctx->stream));
  std::vector<int> tile((size_t)std::max<long long>(L.n_items, 1), 0);
  for (int t = 0; t < L.n_tiles; ++t) {
    const int b = rip[(size_t)t * (L.rows + 1)], e = rip[(size_t)t * (L.rows + 1) + L.rows];
    for (int k = b; k < e; ++k) tile[k] = t;
  }
  RET(dev_alloc(ctx, out, tile.size()));
  CK(ctx, cudaMemcpyAsync(*out, tile.data(), tile.size() * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
  CK(ctx, cudaStreamSynchronize(ctx->stream));
  return M3B_OK;
}

int run_tfidf(m3b_mat* m, int frequencies, int log_scale_tf, double scale_factor, const double* row_sums_in,
              double num_cols, double* col_sums_out, double* row_sums_out) {
  m3b_ctx* ctx = m->ctx;
  if (!m->selected || !m->layouts_built) {
    m3b_set_error(ctx, "m3b_tfidf before m3b_select_genes");
    return M3B_E_STATE;
  }
  if (m->tfidf_applied) {
    m3b_set_error(ctx, "m3b_tfidf applied twice: call m3b_select_genes again first");
    return M3B_E_STATE;
  }
  if (m->f0 != 0.0) {
    m3b_set_error(ctx, "TF-IDF on the dense pseudo-count path is not supported");
    return M3B_E_STATE;
  }
  // TF-IDF rewrites every stored value per (gene, cell): the index-only unit layouts cannot hold
  // that, so the layouts are re-built with all entries in the general stream first
  if (m->split) RET(rebuild_layouts(m, false));
  const int n = m->n_kept;
  const long long C = m->n_cells;
  if (num_cols <= 0) num_cols = (double)m->n_cells_total;
  double *d_colsum = nullptr, *d_rowpos = nullptr, *d_idf = nullptr;
  int *tileA = nullptr, *tileB = nullptr;
  DevScope guard;
  guard.own(&d_colsum);
  guard.own(&d_rowpos);
  guard.own(&d_idf);
  guard.own(&tileA);
  guard.own(&tileB);
  RET(dev_alloc(ctx, &d_colsum, (size_t)std::max<long long>(C, 1)));
  RET(dev_alloc(ctx, &d_rowpos, (size_t)std::max(n, 1)));
  RET(dev_alloc(ctx, &d_idf, (size_t)std::max(n, 1)));
  int st = M3B_OK;
  std::vector<double> rowpos(std::max(n, 1), 0.0), idf(std::max(n, 1), 0.0);
  do {
    // col_sums = colSums(FM) over the kept genes (local cells)
    CK(ctx, cudaMemsetAsync(d_colsum, 0, std::max<long long>(C, 1) * sizeof(double), ctx->stream));
    if ((st = launch_tile_stream(ctx, m->A, OP_SUM, nullptr)) != M3B_OK) break;
    if ((st = launch_combine_rows(ctx, m->A, d_colsum)) != M3B_OK) break;
    // row_sums = rowSums(FM > 0), summed over the ranks -- or the model's, when projecting
    if (row_sums_in) {
      for (int g = 0; g < n; ++g) rowpos[g] = row_sums_in[g];
    } else {
      CK(ctx, cudaMemsetAsync(d_rowpos, 0, std::max(n, 1) * sizeof(double), ctx->stream));
      if ((st = launch_tile_stream(ctx, m->B, OP_COUNTPOS, nullptr)) != M3B_OK) break;
      if ((st = launch_combine_rows(ctx, m->B, d_rowpos)) != M3B_OK) break;
      if ((st = comm_allreduce_f64(ctx, d_rowpos, n)) != M3B_OK) break;
      if (n) CK(ctx, cudaMemcpyAsync(rowpos.data(), d_rowpos, n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
      CK(ctx, cudaStreamSynchronize(ctx->stream));
    }
    for (int g = 0; g < n; ++g) idf[g] = std::log(1.0 + num_cols / rowpos[g]);  // R/preprocess_cds.R:316
    if (n) CK(ctx, cudaMemcpyAsync(d_idf, idf.data(), n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    if ((st = item_tiles(ctx, m->A, &tileA)) != M3B_OK) break;
    if ((st = item_tiles(ctx, m->B, &tileB)) != M3B_OK) break;
    const int blocks = ctx->sm_count * 8;
    if (m->A.n_items)
      k_tfidf_items<true><<<blocks, 256, 0, ctx->stream>>>(m->A.idx, m->A.val, m->A.items, m->A.n_items, tileA, m->A.tile_w,
                                                           d_colsum, d_idf, frequencies, log_scale_tf, scale_factor);
    if (m->B.n_items)
      k_tfidf_items<false><<<blocks, 256, 0, ctx->stream>>>(m->B.idx, m->B.val, m->B.items, m->B.n_items, tileB, m->B.tile_w,
                                                            d_colsum, d_idf, frequencies, log_scale_tf, scale_factor);
    count_launch(ctx, 2);
    CK(ctx, cudaGetLastError());
    if (col_sums_out && C) CK(ctx, cudaMemcpyAsync(col_sums_out, d_colsum, C * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CK(ctx, cudaStreamSynchronize(ctx->stream));
    if (row_sums_out)
      for (int g = 0; g < n; ++g) row_sums_out[g] = rowpos[g];
    m->tfidf_applied = true;
    m->moments_done = false;
  } while (0);
  return st;
}
