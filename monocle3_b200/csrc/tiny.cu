// irlba's tiny-problem branch: min(nrow, ncol) < 6  =>  a dense SVD instead of the Lanczos process
// (irlba 2.3.x R/irlba.R: "Tiny problem detected, using standard `svd` function"; SURVEY.md
// Appendix A.2).  preprocess_cds reaches it only with toy inputs (fewer than 6 cells or fewer than
// 6 kept genes), but the reference does not fail there and neither does this library.
//
// The operator is the same implicitly centred / scaled matrix as in irlba.cu,
//   Ahat[c, g] = (y[c, g] - ce[g]) / s[g]        (cells x kept genes, y = 0 where nothing is stored),
// materialised here because one of its dimensions is at most 5: Ahat is built on the device, the
// k x k Gram matrix of its short side (k = min dim) is reduced in one block with a fixed order, a
// cyclic Jacobi eigen-solver (one thread, k <= 5) gives the singular values and the short-side
// vectors, and one more pass over Ahat gives the long-side vectors.  Unsharded matrices only.
#include <algorithm>
#include <vector>

#include "m3b_internal.h"

__global__ void k_tiny_fill(double* __restrict__ A, long long m, int n, const double* __restrict__ ce,
                            const double* __restrict__ s) {
  const long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (e >= m * n) return;
  const int g = (int)(e / m);
  double v = ce ? -ce[g] : 0.0;
  if (s) v = v / s[g];
  A[e] = v;
}

__global__ void k_tiny_scatter(const long long* __restrict__ p, const int* __restrict__ gi, const double* __restrict__ y,
                               const int* __restrict__ kept_of_gene, long long m, const double* __restrict__ ce,
                               const double* __restrict__ s, double* __restrict__ A) {
  const int lane = threadIdx.x & 31;
  long long warp = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long c = warp; c < m; c += nwarps)
    for (long long k = p[c] + lane; k < p[c + 1]; k += 32) {
      const int g = kept_of_gene[gi[k]];
      if (g < 0) continue;
      double v = y[k] - (ce ? ce[g] : 0.0);
      if (s) v = v / s[g];
      A[(long long)g * m + c] = v;
    }
}

// G = A^T A (short side = columns, k = n) or A A^T (short side = rows, k = m); one block, fixed order
__global__ void __launch_bounds__(256) k_tiny_gram(const double* __restrict__ A, long long m, int n, int cols_short,
                                                   double* __restrict__ G) {
  __shared__ double sh[256];
  const int k = cols_short ? n : (int)m;
  const long long len = cols_short ? m : n;
  for (int a = 0; a < k; ++a)
    for (int b = a; b < k; ++b) {
      double acc = 0.0;
      for (long long r = threadIdx.x; r < len; r += blockDim.x) {
        const double x = cols_short ? A[(long long)a * m + r] : A[r * m + a];
        const double z = cols_short ? A[(long long)b * m + r] : A[r * m + b];
        acc += x * z;
      }
      sh[threadIdx.x] = acc;
      __syncthreads();
      if (threadIdx.x == 0) {
        double t = 0.0;
        for (int q = 0; q < (int)blockDim.x; ++q) t += sh[q];
        G[a * k + b] = t;
        G[b * k + a] = t;
      }
      __syncthreads();
    }
}

// symmetric k x k (k <= 5): cyclic Jacobi; eigenvalues descending in lam, eigenvectors in the columns of E
__global__ void k_tiny_eig(int k, double* __restrict__ G, double* __restrict__ lam, double* __restrict__ E) {
  if (blockIdx.x || threadIdx.x) return;
  for (int a = 0; a < k; ++a)
    for (int b = 0; b < k; ++b) E[a * k + b] = (a == b) ? 1.0 : 0.0;
  for (int sweep = 0; sweep < 60; ++sweep) {
    double off = 0.0, diag = 0.0;
    for (int a = 0; a < k; ++a)
      for (int b = 0; b < k; ++b) (a == b ? diag : off) += G[a * k + b] * G[a * k + b];
    if (off <= 1e-32 * diag) break;
    for (int pp = 0; pp < k; ++pp)
      for (int qq = pp + 1; qq < k; ++qq) {
        const double apq = G[pp * k + qq];
        if (apq == 0.0) continue;
        const double theta = (G[qq * k + qq] - G[pp * k + pp]) / (2.0 * apq);
        const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
        const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
        for (int r = 0; r < k; ++r) {  // columns pp, qq
          const double gp = G[r * k + pp], gq = G[r * k + qq];
          G[r * k + pp] = c * gp - s * gq;
          G[r * k + qq] = s * gp + c * gq;
        }
        for (int r = 0; r < k; ++r) {  // rows pp, qq
          const double gp = G[pp * k + r], gq = G[qq * k + r];
          G[pp * k + r] = c * gp - s * gq;
          G[qq * k + r] = s * gp + c * gq;
        }
        for (int r = 0; r < k; ++r) {
          const double ep = E[r * k + pp], eq = E[r * k + qq];
          E[r * k + pp] = c * ep - s * eq;
          E[r * k + qq] = s * ep + c * eq;
        }
      }
  }
  // selection sort, descending
  for (int a = 0; a < k; ++a) lam[a] = G[a * k + a];
  for (int a = 0; a < k; ++a) {
    int best = a;
    for (int b = a + 1; b < k; ++b)
      if (lam[b] > lam[best]) best = b;
    if (best != a) {
      const double t = lam[a];
      lam[a] = lam[best];
      lam[best] = t;
      for (int r = 0; r < k; ++r) {
        const double e = E[r * k + a];
        E[r * k + a] = E[r * k + best];
        E[r * k + best] = e;
      }
    }
  }
}

// long-side vectors.  cols_short: X[c, j] = sum_g A[c, g] E[g, j] (= u d), V = E.
// otherwise:  V[g, j] = sum_c A[c, g] E[c, j] / d_j,  X[c, j] = E[c, j] d_j.
__global__ void k_tiny_finish(const double* __restrict__ A, long long m, int n, int cols_short, int nv,
                              const double* __restrict__ lam, const double* __restrict__ E, double* __restrict__ d,
                              double* __restrict__ V, double* __restrict__ X) {
  const int k = cols_short ? n : (int)m;
  const long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (e < nv) d[e] = sqrt(fmax(lam[e], 0.0));
  if (cols_short) {
    if (e < m * nv) {
      const long long c = e % m;
      const int j = (int)(e / m);
      double acc = 0.0;
      for (int g = 0; g < k; ++g) acc += A[(long long)g * m + c] * E[g * k + j];
      X[e] = acc;
    }
    if (e < (long long)n * nv) V[e] = E[(e % n) * k + (e / n)];
  } else {
    if (e < (long long)n * nv) {
      const long long g = e % n;
      const int j = (int)(e / n);
      const double dj = sqrt(fmax(lam[j], 0.0));
      double acc = 0.0;
      for (int c = 0; c < k; ++c) acc += A[g * m + c] * E[c * k + j];
      V[e] = dj > 0.0 ? acc / dj : 0.0;
    }
    if (e < m * nv) X[e] = E[(e % m) * k + (e / m)] * sqrt(fmax(lam[e / m], 0.0));
  }
}

int run_irlba_tiny(m3b_mat* mt, int nv, int center, int scale, double* d_out, double* V_out, double* X_out, long long ldx,
                   double* center_out, double* scale_out, int* iter_out, int* mprod_out, bool copy_x) {
  m3b_ctx* ctx = mt->ctx;
  const int n = mt->n_kept;
  const long long m = mt->n_cells;
  if (ctx->comm.world != 1 || m != mt->n_cells_total) {
    m3b_set_error(ctx, "irlba's dense branch (min dimension < 6) is implemented for unsharded matrices only");
    return M3B_E_DIMS;
  }
  const int cols_short = n <= m;
  const int k = cols_short ? n : (int)m;
  if (center || scale) RET(compute_moments(mt));
  double *A = nullptr, *G = nullptr, *lam = nullptr, *E = nullptr, *dd = nullptr, *V = nullptr, *X = nullptr, *cef = nullptr;
  const double* s = scale ? mt->scale : nullptr;
  const double* ce = center ? mt->mean_sparse : nullptr;
  int st = M3B_OK;
  cudaError_t e = cudaSuccess;
  std::vector<double> fill;
  do {
    if ((st = dev_alloc(ctx, &A, (size_t)std::max<long long>(m * n, 1))) != M3B_OK) break;
    if ((st = dev_alloc(ctx, &G, (size_t)k * k + 1)) != M3B_OK) break;
    if ((st = dev_alloc(ctx, &lam, (size_t)k + 1)) != M3B_OK) break;
    if ((st = dev_alloc(ctx, &E, (size_t)k * k + 1)) != M3B_OK) break;
    if ((st = dev_alloc(ctx, &dd, (size_t)nv)) != M3B_OK) break;
    if ((st = dev_alloc(ctx, &V, (size_t)n * nv)) != M3B_OK) break;
    if ((st = dev_alloc(ctx, &X, (size_t)std::max<long long>(m * nv, 1))) != M3B_OK) break;
    if (!center && mt->f0 != 0.0) {  // dense pseudo-count path without centring: the implicit constant is -f0
      if ((st = dev_alloc(ctx, &cef, (size_t)n)) != M3B_OK) break;
      fill.assign(n, -mt->f0);
      if ((e = cudaMemcpyAsync(cef, fill.data(), n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream)) != cudaSuccess) break;
      ce = cef;
    }
    const long long tot = m * n;
    k_tiny_fill<<<(unsigned)((tot + 255) / 256), 256, 0, ctx->stream>>>(A, m, n, ce, s);
    k_tiny_scatter<<<(unsigned)std::min<long long>((m + 7) / 8, 1184), 256, 0, ctx->stream>>>(mt->p, mt->i, mt->y, mt->kept_of_gene, m,
                                                                                          ce, s, A);
    k_tiny_gram<<<1, 256, 0, ctx->stream>>>(A, m, n, cols_short, G);
    k_tiny_eig<<<1, 32, 0, ctx->stream>>>(k, G, lam, E);
    const long long work = std::max<long long>(std::max<long long>(m, n) * nv, nv);
    k_tiny_finish<<<(unsigned)((work + 255) / 256), 256, 0, ctx->stream>>>(A, m, n, cols_short, nv, lam, E, dd, V, X);
    count_launch(ctx, 5);
    if ((e = cudaGetLastError()) != cudaSuccess) break;
    if ((e = cudaMemcpyAsync(d_out, dd, nv * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream)) != cudaSuccess) break;
    if (V_out && (e = cudaMemcpyAsync(V_out, V, (size_t)n * nv * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream)) != cudaSuccess) break;
    if (copy_x && X_out && m &&
        (e = cudaMemcpy2DAsync(X_out, ldx * sizeof(double), X, m * sizeof(double), m * sizeof(double), nv, cudaMemcpyDeviceToHost,
                               ctx->stream)) != cudaSuccess)
      break;
    if (center_out && center && (e = cudaMemcpyAsync(center_out, mt->center, n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream)) != cudaSuccess) break;
    if (scale_out && scale && (e = cudaMemcpyAsync(scale_out, mt->scale, n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream)) != cudaSuccess) break;
    e = cudaStreamSynchronize(ctx->stream);
  } while (0);
  if (st == M3B_OK && e != cudaSuccess) {
    m3b_set_error(ctx, "irlba (dense branch): %s", cudaGetErrorString(e));
    st = M3B_E_CUDA;
  }
  dev_free(mt->Xdev);
  mt->Xdev = nullptr;
  if (st == M3B_OK && !copy_x) {
    mt->Xdev = X;
    mt->Xdev_nv = nv;
    X = nullptr;
  }
  dev_free(A);
  dev_free(G);
  dev_free(lam);
  dev_free(E);
  dev_free(dd);
  dev_free(V);
  dev_free(X);
  dev_free(cef);
  if (iter_out) *iter_out = 0;
  if (mprod_out) *mprod_out = 0;
  return st;
}
