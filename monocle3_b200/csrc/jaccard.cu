// SURVEY section 8(f) row 3 (first half): Jaccard weights of a kNN graph -- the reference's only
// native kernel, jaccard_coeff (src/clustering.cpp:13-58, exported at src/RcppExports.cpp as
// _monocle3_jaccard_coeff; called from R/cluster_cells.R:313 and R/graph_test.R:429,506).
//
//   idx: n x k matrix of 1-based neighbour ids (an R NumericMatrix: doubles, column-major).
//   For every edge (i, j-th neighbour kk):  u = |unique(idx[i,]) ∩ unique(idx[kk,])|,
//   w = u / (2k - u) if weight and u > 0, else 1;  out row r = i*k + j = (i+1, kk+1, w);
//   finally the weight column is divided by its maximum.
//
// One warp per node i: its neighbour list sits in shared memory; for each neighbour kk the lanes
// walk kk's list 32 elements at a time (k <= 2048: the reference accepts any k, cluster_cells'
// default is 20), a value counts once (the first lane that holds it, __match_any_sync, and not seen
// in an earlier chunk) if it occurs in i's list.  Integer work; the two divisions are the same
// IEEE operations as the reference's, so the result is bit-identical.
#include <algorithm>

#include "m3b_internal.h"

#define JC_WARPS 8
#define JC_KMAX 2048

// doubles, column-major, 1-based  ->  int32, row-major, 0-based
__global__ void k_jc_pack(const double* __restrict__ idx, long long n, int k, int* __restrict__ out, int* __restrict__ bad) {
  const long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (e >= n * k) return;
  const long long i = e / k;
  const int j = (int)(e % k);
  const double v = idx[(size_t)j * n + i];
  const long long t = (long long)v - 1;
  if (!(v >= 1.0) || t >= n) atomicOr(bad, 1);
  out[e] = (int)t;
}

__global__ void __launch_bounds__(32 * JC_WARPS) k_jaccard(const int* __restrict__ nb, long long n, int k, int weight,
                                                           double* __restrict__ out, unsigned long long* __restrict__ wmax) {
  extern __shared__ int jc_rows[];  // [JC_WARPS][k]
  int* const myrow = jc_rows + (size_t)(threadIdx.x >> 5) * k;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const long long total = n * (long long)k;
  double* o0 = out;
  double* o1 = out + total;
  double* o2 = out + 2 * total;
  double mx = 0.0;
  for (long long i = blockIdx.x * (long long)JC_WARPS + w; i < n; i += (long long)gridDim.x * JC_WARPS) {
    __syncwarp();
    for (int t = lane; t < k; t += 32) myrow[t] = nb[i * k + t];
    __syncwarp();
    for (int j = 0; j < k; ++j) {
      const int kk = myrow[j];
      double wt = 1.0;
      if (weight) {
        int u = 0;
        for (int half = 0; 32 * half < k; ++half) {  // chunk of 32 elements of kk's list
          const int t = lane + 32 * half;
          const bool have = t < k;
          const int v = have ? nb[(long long)kk * k + t] : -1 - t;  // distinct dummies for idle lanes
          // first occurrence inside kk's list: no lower lane of this chunk with the same value, and
          // not present in an earlier chunk
          const unsigned peers = __match_any_sync(0xffffffffu, v);
          bool first = have && ((peers & ((1u << lane) - 1u)) == 0u);
          if (half >= 1 && first)
            for (int q = 0; q < 32 * half; ++q) first = first && (nb[(long long)kk * k + q] != v);
          bool in_i = false;
          if (first)
            for (int q = 0; q < k; ++q) in_i = in_i || (myrow[q] == v);
          u += __popc(__ballot_sync(0xffffffffu, first && in_i));
        }
        if (u > 0) wt = (double)u / (double)(2 * k - u);
      }
      if (lane == 0) {
        const long long r = i * k + j;
        o0[r] = (double)(i + 1);
        o1[r] = (double)(kk + 1);
        o2[r] = wt;
      }
      mx = fmax(mx, wt);
    }
  }
  // weights are positive: their IEEE bit patterns order like the values
  if (lane == 0 && mx > 0.0) atomicMax(wmax, (unsigned long long)__double_as_longlong(mx));
}

__global__ void k_jc_scale(double* __restrict__ w, long long total, const unsigned long long* __restrict__ wmax) {
  const double mx = __longlong_as_double((long long)*wmax);
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x)
    w[e] = w[e] / mx;
}

int run_jaccard(m3b_ctx* ctx, const double* idx, long long n, int k, int weight, double* out) {
  if (n < 1 || k < 1 || k > JC_KMAX) {
    m3b_set_error(ctx, "m3b_jaccard_coeff: need n >= 1 and 1 <= k <= %d", JC_KMAX);
    return M3B_E_DIMS;
  }
  const long long total = n * (long long)k;
  double *d_idx = nullptr, *d_out = nullptr;
  int* d_nb = nullptr;
  unsigned long long* d_max = nullptr;
  DevScope guard;
  guard.own(&d_idx);
  guard.own(&d_out);
  guard.own(&d_nb);
  guard.own(&d_max);
  RET(dev_alloc(ctx, &d_idx, (size_t)total));
  RET(dev_alloc(ctx, &d_out, (size_t)3 * total));
  RET(dev_alloc(ctx, &d_nb, (size_t)total + 2));
  RET(dev_alloc(ctx, &d_max, (size_t)2));
  int* d_bad = d_nb + total;
  int st = M3B_OK;
  cudaError_t e = cudaSuccess;
  do {
    if ((e = cudaMemcpyAsync(d_idx, idx, (size_t)total * sizeof(double), cudaMemcpyHostToDevice, ctx->stream)) != cudaSuccess) break;
    ctx->stats.h2d_bytes += total * 8;
    if ((e = cudaMemsetAsync(d_bad, 0, 2 * sizeof(int), ctx->stream)) != cudaSuccess) break;
    if ((e = cudaMemsetAsync(d_max, 0, 2 * sizeof(unsigned long long), ctx->stream)) != cudaSuccess) break;
    k_jc_pack<<<(unsigned)((total + 255) / 256), 256, 0, ctx->stream>>>(d_idx, n, k, d_nb, d_bad);
    count_launch(ctx);
    int bad = 0;
    if ((e = cudaMemcpyAsync(&bad, d_bad, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream)) != cudaSuccess) break;
    if ((e = cudaStreamSynchronize(ctx->stream)) != cudaSuccess) break;
    if (bad) {
      m3b_set_error(ctx, "m3b_jaccard_coeff: neighbour ids must lie in 1..n");
      st = M3B_E_DIMS;
      break;
    }
    const int blocks = (int)std::min<long long>((n + JC_WARPS - 1) / JC_WARPS, (long long)ctx->sm_count * 8);
    const size_t jc_smem = (size_t)JC_WARPS * k * sizeof(int);
    if (jc_smem > 48 * 1024 &&
        (e = cudaFuncSetAttribute(k_jaccard, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)jc_smem)) != cudaSuccess)
      break;
    k_jaccard<<<blocks, 32 * JC_WARPS, jc_smem, ctx->stream>>>(d_nb, n, k, weight ? 1 : 0, d_out, d_max);
    k_jc_scale<<<std::min<long long>((total + 255) / 256, (long long)ctx->sm_count * 8), 256, 0, ctx->stream>>>(d_out + 2 * total, total, d_max);
    count_launch(ctx, 2);
    if ((e = cudaGetLastError()) != cudaSuccess) break;
    if ((e = cudaMemcpyAsync(out, d_out, (size_t)3 * total * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream)) != cudaSuccess) break;
    ctx->stats.d2h_bytes += 3 * total * 8;
    e = cudaStreamSynchronize(ctx->stream);
  } while (0);
  if (e != cudaSuccess) {
    m3b_set_error(ctx, "m3b_jaccard_coeff: %s", cudaGetErrorString(e));
    return M3B_E_CUDA;
  }
  return st;
}
