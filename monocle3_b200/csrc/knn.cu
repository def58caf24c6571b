// SURVEY section 8(f) row 3 (second half): exact k nearest neighbours of every row of a dense
// n x d matrix among the rows themselves (Euclidean) -- what cluster_cells asks RANN for with
// nn_control method "nn2": RANN::nn2(data, data, k + 1, searchtype = "standard")
// (R/cluster_cells.R:286-287); the embedding it runs on is reducedDims(cds)[["PCA"]], the output
// of this library's PCA path.  The approximate default (annoy) has no exact parity target.
//
// Brute force, exact: distances are accumulated as sum (x_d - y_d)^2 like a kd-tree's leaf scan
// (not as |x|^2 + |y|^2 - 2 x.y, whose cancellation reorders near neighbours), 64 x 64 tiles with
// a 4 x 4 register block per thread; the 64 query rows of a block stay in shared memory for the
// whole sweep, candidate tiles stream from L2.  Every query row keeps its k best in a sorted list
// in shared memory; one thread per row scans a finished tile against the row's current k-th
// distance (after the first tiles almost nothing passes).  Ties go to the smaller index.
#include <algorithm>

#include "m3b_internal.h"

#define KN_T 64       // tile edge
#define KN_KMAX 1024  // (what really bounds k is the shared memory of the sorted per-row lists, checked at launch)
#define KN_DC 32      // dimensions staged per step for the candidate tile

__global__ void __launch_bounds__(256) k_knn_self(const double* __restrict__ X, long long n, int d, int k,
                                                  int* __restrict__ idx_out, double* __restrict__ dist_out) {
  extern __shared__ double kn_sm[];
  double* Qs = kn_sm;                                   // [d][KN_T]
  double* Cs = Qs + (size_t)d * KN_T;                   // [KN_DC][KN_T]
  double* D = Cs + (size_t)KN_DC * KN_T;                // [KN_T][KN_T + 1]
  double* ld = D + (size_t)KN_T * (KN_T + 1);           // [KN_T][k]
  int* li = reinterpret_cast<int*>(ld + (size_t)KN_T * k);  // [KN_T][k]
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const long long q0 = (long long)blockIdx.x * KN_T;
  for (int e = tid; e < d * KN_T; e += 256) {
    const int r = e % KN_T, kk = e / KN_T;
    Qs[e] = (q0 + r < n) ? X[(size_t)kk * n + q0 + r] : 0.0;
  }
  for (int e = tid; e < KN_T * k; e += 256) {
    ld[e] = INFINITY;
    li[e] = -1;
  }
  __syncthreads();
  for (long long c0 = 0; c0 < n; c0 += KN_T) {
    double acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = 0.0;
    for (int k0 = 0; k0 < d; k0 += KN_DC) {
      const int dc = min(KN_DC, d - k0);
      for (int e = tid; e < dc * KN_T; e += 256) {
        const int r = e % KN_T, kk = e / KN_T;
        Cs[e] = (c0 + r < n) ? X[(size_t)(k0 + kk) * n + c0 + r] : 0.0;
      }
      __syncthreads();
      for (int kk = 0; kk < dc; ++kk) {
        double a[4], b[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) a[i] = Qs[(size_t)(k0 + kk) * KN_T + ty + 16 * i];
#pragma unroll
        for (int j = 0; j < 4; ++j) b[j] = Cs[kk * KN_T + tx + 16 * j];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const double t = a[i] - b[j];
            acc[i][j] = fma(t, t, acc[i][j]);
          }
      }
      __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j)
        D[(ty + 16 * i) * (KN_T + 1) + tx + 16 * j] = (c0 + tx + 16 * j < n) ? acc[i][j] : INFINITY;
    __syncthreads();
    if (tid < KN_T && q0 + tid < n) {
      double* rd = ld + (size_t)tid * k;
      int* ri = li + (size_t)tid * k;
      double thr = rd[k - 1];
      for (int c = 0; c < KN_T; ++c) {
        const double v = D[tid * (KN_T + 1) + c];
        if (v < thr) {  // strict: on a tie the earlier (smaller) index stays
          int s = k - 1;
          while (s > 0 && rd[s - 1] > v) {
            rd[s] = rd[s - 1];
            ri[s] = ri[s - 1];
            --s;
          }
          rd[s] = v;
          ri[s] = (int)(c0 + c);
          thr = rd[k - 1];
        }
      }
    }
    __syncthreads();
  }
  for (int e = tid; e < KN_T * k; e += 256) {
    const int r = e % KN_T, s = e / KN_T;
    if (q0 + r < n) {
      idx_out[(size_t)s * n + q0 + r] = li[(size_t)r * k + s] + 1;      // 1-based, column-major n x k
      dist_out[(size_t)s * n + q0 + r] = sqrt(ld[(size_t)r * k + s]);
    }
  }
}

int run_knn_self(m3b_ctx* ctx, const double* X, long long n, int d, int k, int* idx_out, double* dist_out) {
  if (n < 1 || d < 1 || k < 1 || k > KN_KMAX || k > n) {
    m3b_set_error(ctx, "m3b_knn_self: need n >= 1, d >= 1, 1 <= k <= min(n, %d)", KN_KMAX);
    return M3B_E_DIMS;
  }
  const size_t smem = ((size_t)d * KN_T + (size_t)KN_DC * KN_T + (size_t)KN_T * (KN_T + 1) + (size_t)KN_T * k) * sizeof(double) +
                      (size_t)KN_T * k * sizeof(int);
  if (smem > 220 * 1024) {
    m3b_set_error(ctx, "m3b_knn_self: d = %d with k = %d does not fit the shared memory of a query block (%zu KB > 220 KB)", d, k,
                  smem / 1024);
    return M3B_E_DIMS;
  }
  double *d_X = nullptr, *d_dist = nullptr;
  int* d_idx = nullptr;
  DevScope guard;
  guard.own(&d_X);
  guard.own(&d_dist);
  guard.own(&d_idx);
  RET(dev_alloc(ctx, &d_X, (size_t)n * d));
  RET(dev_alloc(ctx, &d_dist, (size_t)n * k));
  RET(dev_alloc(ctx, &d_idx, (size_t)n * k));
  cudaError_t e = cudaSuccess;
  do {
    if ((e = cudaMemcpyAsync(d_X, X, (size_t)n * d * sizeof(double), cudaMemcpyHostToDevice, ctx->stream)) != cudaSuccess) break;
    ctx->stats.h2d_bytes += n * d * 8;
    if ((e = cudaFuncSetAttribute(k_knn_self, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) break;
    k_knn_self<<<(unsigned)((n + KN_T - 1) / KN_T), 256, smem, ctx->stream>>>(d_X, n, d, k, d_idx, d_dist);
    count_launch(ctx);
    if ((e = cudaGetLastError()) != cudaSuccess) break;
    if ((e = cudaMemcpyAsync(idx_out, d_idx, (size_t)n * k * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream)) != cudaSuccess) break;
    if ((e = cudaMemcpyAsync(dist_out, d_dist, (size_t)n * k * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream)) != cudaSuccess) break;
    ctx->stats.d2h_bytes += n * k * 12;
    e = cudaStreamSynchronize(ctx->stream);
  } while (0);
  if (e != cudaSuccess) {
    m3b_set_error(ctx, "m3b_knn_self: %s", cudaGetErrorString(e));
    return M3B_E_CUDA;
  }
  return M3B_OK;
}
