// K10: projection of query cells onto saved loadings (preprocess_transform).
//
// Replaces R/projection.R:135-148: DelayedArray realises each block of
// ((x - center)/scale) densely and multiplies by svd_v with dgemm
// (matrix_multiply_multicore, R/utils.R:823-849).  Here the query stays sparse:
//   Y = A'_q (S^-1 V) - 1 b^T,   b = sum_g ((center_g - f0)/scale_g) V[g,:]
// over the genes kept in the query and present in the model.
//
// One warp per query cell walks the cell's items in layout A; the 32 lanes own
// the nv output columns (up to 4 each), entries are loaded 32 at a time and
// broadcast with shuffles, and each entry pulls one contiguous row of the
// (L2-resident, row-major) scaled loadings.
#include <algorithm>
#include <vector>

#include "m3b_internal.h"

#define PJ_CPL 4  // columns per lane => nv <= 128 per pass

// The row's entries come from up to two layouts: the general one (idx + val) and the unit one
// (count-1 entries: index only, value = uval[row]; pads carry idx == tile_w).
__global__ void __launch_bounds__(256) k_project(const unsigned short* __restrict__ idx, const double* __restrict__ val,
                                                 const ItemDesc* __restrict__ items, const int* __restrict__ rip,
                                                 const unsigned short* __restrict__ uidx,
                                                 const ItemDesc* __restrict__ uitems, const int* __restrict__ urip,
                                                 const double* __restrict__ uval,
                                                 long long rows, int n_tiles, int tile_w,
                                                 const double* __restrict__ Vs /* [G'q][nv] row-major */,
                                                 const double* __restrict__ b, int nv, int col0,
                                                 double* __restrict__ Y, long long ldy) {
  const int lane = threadIdx.x & 31;
  long long warp = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long r = warp; r < rows; r += nwarps) {
    double acc[PJ_CPL];
#pragma unroll
    for (int a = 0; a < PJ_CPL; ++a) acc[a] = 0.0;
    for (int pass = 0; pass < (uidx ? 2 : 1); ++pass)
    for (int t = 0; t < n_tiles; ++t) {
      const int* p = (pass ? urip : rip) + (long long)t * (rows + 1) + r;
      const int ib = p[0], ie = p[1];
      const long long goff = (long long)t * tile_w;
      const double ur = pass ? uval[r] : 0.0;
      for (int it = ib; it < ie; ++it) {
        const ItemDesc d = pass ? uitems[it] : items[it];
        for (int e0 = 0; e0 < d.len; e0 += 32) {
          int myi = 0;
          double myv = 0.0;
          if (e0 + lane < d.len) {
            if (pass) {
              myi = (int)(uidx[d.start + e0 + lane] & 0x7fff);
              myv = (myi < tile_w) ? ur : 0.0;
              if (myi >= tile_w) myi = 0;
            } else {
              myi = (int)(idx[d.start + e0 + lane] & 0x7fff);
              myv = val[d.start + e0 + lane];
            }
          }
          const int cnt = min(32, d.len - e0);
          for (int q = 0; q < cnt; ++q) {
            const int g = __shfl_sync(0xffffffffu, myi, q);
            const double v = __shfl_sync(0xffffffffu, myv, q);
            const double* row = Vs + (goff + g) * (long long)nv + col0;
#pragma unroll
            for (int a = 0; a < PJ_CPL; ++a) {
              const int c = lane + 32 * a;
              if (col0 + c < nv) acc[a] += v * row[c];
            }
          }
        }
      }
    }
#pragma unroll
    for (int a = 0; a < PJ_CPL; ++a) {
      const int c = col0 + lane + 32 * a;
      if (c < nv) Y[(long long)c * ldy + r] = acc[a] - b[c];
    }
  }
}

int run_project(m3b_mat* q, const int* model_row, int n_model, const double* V, const double* center,
                const double* scale, int nv, double* Y, long long ldy) {
  m3b_ctx* ctx = q->ctx;
  if (!q->selected) {
    m3b_set_error(ctx, "m3b_project: call m3b_normalize and m3b_select_genes on the query first");
    return M3B_E_STATE;
  }
  const int n = q->n_kept;
  const long long rows = q->n_cells;
  if (nv <= 0 || ldy < rows) {
    m3b_set_error(ctx, "m3b_project: bad nv/ldy");
    return M3B_E_DIMS;
  }
  for (int g = 0; g < n; ++g)
    if (model_row[g] < 0 || model_row[g] >= n_model) {
      m3b_set_error(ctx, "gene sets differ: genes in the subject matrix are not in the reference matrix");
      return M3B_E_DIMS;
    }
  // scaled loadings (row-major) and the constant row b, built on the host: G'q x nv is tiny
  std::vector<double> Vs((size_t)std::max(n, 1) * nv), b(nv, 0.0);
  std::vector<long double> bacc(nv, 0.0L);
  for (int g = 0; g < n; ++g) {
    const int r = model_row[g];
    const double s = scale ? scale[r] : 1.0;
    const double ce = (center ? center[r] : 0.0) - q->f0;
    for (int c = 0; c < nv; ++c) {
      const double v = V[(size_t)c * n_model + r];
      Vs[(size_t)g * nv + c] = v / s;
      bacc[c] += (long double)(ce / s) * v;
    }
  }
  for (int c = 0; c < nv; ++c) b[c] = (double)bacc[c];
  double *dVs = nullptr, *db = nullptr, *dY = nullptr;
  RET(dev_alloc(ctx, &dVs, Vs.size()));
  RET(dev_alloc(ctx, &db, (size_t)nv));
  RET(dev_alloc(ctx, &dY, (size_t)std::max<long long>(rows, 1) * nv));
  int status = M3B_OK;
  cudaError_t e;
  do {
    if ((e = cudaMemcpyAsync(dVs, Vs.data(), Vs.size() * sizeof(double), cudaMemcpyHostToDevice, ctx->stream)) != cudaSuccess) break;
    if ((e = cudaMemcpyAsync(db, b.data(), nv * sizeof(double), cudaMemcpyHostToDevice, ctx->stream)) != cudaSuccess) break;
    ctx->stats.h2d_bytes += (long long)Vs.size() * 8 + nv * 8;
    if (rows) {
      int blocks = (int)std::min<long long>((rows + 7) / 8, (long long)ctx->sm_count * 8);
      for (int col0 = 0; col0 < nv; col0 += 32 * PJ_CPL) {
        const TiledLayout* U = q->A.unit;
        k_project<<<blocks, 256, 0, ctx->stream>>>(q->A.idx, q->A.val, q->A.items, q->A.row_item_ptr,
                                                   U ? U->idx : nullptr, U ? U->items : nullptr,
                                                   U ? U->row_item_ptr : nullptr, U ? U->rowscale : nullptr, rows,
                                                   q->A.n_tiles, q->A.tile_w, dVs, db, nv, col0, dY, rows);
        count_launch(ctx);
      }
      if ((e = cudaGetLastError()) != cudaSuccess) break;
      if ((e = cudaMemcpy2DAsync(Y, ldy * sizeof(double), dY, rows * sizeof(double), rows * sizeof(double), nv,
                                 cudaMemcpyDeviceToHost, ctx->stream)) != cudaSuccess)
        break;
      ctx->stats.d2h_bytes += rows * nv * 8;
    }
    e = cudaStreamSynchronize(ctx->stream);
  } while (0);
  if (e != cudaSuccess) {
    m3b_set_error(ctx, "m3b_project: %s", cudaGetErrorString(e));
    status = M3B_E_CUDA;
  }
  dev_free(dVs);
  dev_free(db);
  dev_free(dY);
  return status;
}
