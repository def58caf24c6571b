// the residual regression of align_cds on the PCA / LSI coordinates.
//
// Replaces  fit <- limma::lmFit(Matrix::t(preproc_res), X.model_mat);  beta <- fit$coefficients[, -1];
//           preproc_res <- t(t(preproc_res) - beta %*% t(X.model_mat[, -1]))       (R/alignment.R:106-117)
// and the re-application of a saved beta in align_beta_transform (R/projection.R:468-478).
// lmFit without weights is ordinary least squares per row of t(X), i.e. per PC: the p x p Gram
// matrix M^T M and the p x nv right-hand sides M^T X are tall-skinny reductions over the cells
// (deterministic two-stage sums on the device, one allreduce when the cells are sharded); the
// p x p solve (p = number of design columns, a handful) is finished on the host in long double on
// the column-equilibrated system with lm.fit's aliasing rule (a design column whose residual norm
// against the columns before it is below 1e-7 of its norm gets NA, which align_cds turns into 0);
// the subtraction is one element-wise pass over X.
#include <math.h>

#include <algorithm>
#include <vector>

#include "m3b_internal.h"

#define AL_THREADS 256
#define AL_ROWS 128  // rows per staged tile
#define AL_LD 129    // shared-memory row stride of a staged column (odd: consecutive columns start on different banks)

// part[blk][a * q + b] = sum over the block's rows of M[r, a] * Z[r, b],  Z = [M | X]  (q = p + nv)
__global__ void __launch_bounds__(AL_THREADS) k_align_gram(const double* __restrict__ M, long long ldm, int p,
                                                           const double* __restrict__ X, long long ldx, int nv, long long n,
                                                           long long rows_per_block, double* __restrict__ part) {
  extern __shared__ double al_sm[];  // [q][AL_LD]
  const int q = p + nv;
  const long long r_lo = (long long)blockIdx.x * rows_per_block, r_hi = min(n, r_lo + rows_per_block);
  const int n_out = p * q;
  // each thread owns outputs o = tid, tid + 256, ... (at most 8 of them are kept in registers per pass)
  for (int o0 = 0; o0 < n_out; o0 += AL_THREADS * 8) {
    double acc[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) acc[u] = 0.0;
    for (long long t0 = r_lo; t0 < r_hi; t0 += AL_ROWS) {
      const int len = (int)min((long long)AL_ROWS, r_hi - t0);
      __syncthreads();
      for (int e = threadIdx.x; e < q * AL_ROWS; e += AL_THREADS) {
        const int c = e / AL_ROWS, r = e % AL_ROWS;
        double v = 0.0;
        if (r < len) v = (c < p) ? M[(long long)c * ldm + t0 + r] : X[(long long)(c - p) * ldx + t0 + r];
        al_sm[c * AL_LD + r] = v;
      }
      __syncthreads();
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int o = o0 + u * AL_THREADS + threadIdx.x;
        if (o < n_out) {
          const double* ma = al_sm + (o / q) * AL_LD;
          const double* zb = al_sm + (o % q) * AL_LD;
          double s = 0.0;
          for (int r = 0; r < AL_ROWS; ++r) s += ma[r] * zb[r];  // rows past len are zero
          acc[u] += s;
        }
      }
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int o = o0 + u * AL_THREADS + threadIdx.x;
      if (o < n_out) part[(long long)blockIdx.x * n_out + o] = acc[u];
    }
  }
}

// out[o] = sum_b part[b][o] in block order (fixed => bit-reproducible)
__global__ void k_align_reduce(const double* __restrict__ part, int nblk, int n_out, double* __restrict__ out) {
  const int o = blockIdx.x * blockDim.x + threadIdx.x;
  if (o >= n_out) return;
  double s = 0.0;
  for (int b = 0; b < nblk; ++b) s += part[(long long)b * n_out + o];
  out[o] = s;
}

// X[r, c] -= sum_t M[r, 1 + t] * beta[c, t]
__global__ void __launch_bounds__(256) k_align_apply(double* __restrict__ X, long long ldx, int nv, long long n,
                                                     const double* __restrict__ M, long long ldm, int p,
                                                     const double* __restrict__ beta /* nv x (p-1), column-major */) {
  const long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (r >= n) return;
  for (int c = 0; c < nv; ++c) {
    double s = 0.0;
    for (int t = 0; t + 1 < p; ++t) s += M[(long long)(t + 1) * ldm + r] * beta[(long long)t * nv + c];
    X[(long long)c * ldx + r] -= s;
  }
}

static int align_upload(m3b_ctx* ctx, const double* X, long long n, int nv, long long ldx, const double* M, int p, double** dX,
                        double** dM) {
  RET(dev_alloc(ctx, dX, (size_t)std::max<long long>(n, 1) * nv));
  RET(dev_alloc(ctx, dM, (size_t)std::max<long long>(n, 1) * p));
  if (n) {
    CK(ctx, cudaMemcpy2DAsync(*dX, n * sizeof(double), X, ldx * sizeof(double), n * sizeof(double), nv, cudaMemcpyHostToDevice,
                              ctx->stream));
    CK(ctx, cudaMemcpyAsync(*dM, M, (size_t)n * p * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    ctx->stats.h2d_bytes += n * (long long)(nv + p) * 8;
  }
  return M3B_OK;
}

static int align_apply_and_download(m3b_ctx* ctx, double* dX, double* dM, double* X, long long n, int nv, long long ldx, int p,
                                    const std::vector<double>& beta) {
  double* dB = nullptr;
  RET(dev_alloc(ctx, &dB, std::max<size_t>(beta.size(), 1)));
  int st = M3B_OK;
  cudaError_t e = cudaSuccess;
  do {
    if (!beta.empty() && (e = cudaMemcpyAsync(dB, beta.data(), beta.size() * sizeof(double), cudaMemcpyHostToDevice, ctx->stream)) != cudaSuccess) break;
    if (n && p > 1) {
      k_align_apply<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(dX, n, nv, n, dM, n, p, dB);
      count_launch(ctx);
    }
    if ((e = cudaGetLastError()) != cudaSuccess) break;
    if (n && (e = cudaMemcpy2DAsync(X, ldx * sizeof(double), dX, n * sizeof(double), n * sizeof(double), nv, cudaMemcpyDeviceToHost,
                                    ctx->stream)) != cudaSuccess)
      break;
    ctx->stats.d2h_bytes += n * (long long)nv * 8;
    e = cudaStreamSynchronize(ctx->stream);
  } while (0);
  dev_free(dB);
  if (e != cudaSuccess) {
    m3b_set_error(ctx, "m3b_align: %s", cudaGetErrorString(e));
    st = M3B_E_CUDA;
  }
  return st;
}

int run_align_residuals(m3b_ctx* ctx, double* X, long long n, int nv, long long ldx, const double* M, int p, double* beta_out,
                        bool fit) {
  if (n < 0 || nv <= 0 || p < 1 || ldx < n || p > 64 || p + nv > 216) {
    m3b_set_error(ctx, "m3b_align_residuals: bad dimensions (n=%lld, nv=%d, p=%d, ldx=%lld; p <= 64, p + nv <= 216)", n, nv, p, ldx);
    return M3B_E_DIMS;
  }
  double *dX = nullptr, *dM = nullptr, *part = nullptr, *dG = nullptr;
  int st = align_upload(ctx, X, n, nv, ldx, M, p, &dX, &dM);
  std::vector<double> beta((size_t)nv * std::max(p - 1, 0), 0.0);
  if (st == M3B_OK && !fit) {
    for (size_t k = 0; k < beta.size(); ++k) beta[k] = beta_out[k];  // re-apply a saved beta (align_beta_transform)
  } else if (st == M3B_OK) {
    const int q = p + nv, n_out = p * q;
    const int nblk = (int)std::max<long long>(1, std::min<long long>(ctx->sm_count * 2, (n + 1023) / 1024));
    const long long rpb = (n + nblk - 1) / nblk;
    std::vector<double> G((size_t)n_out, 0.0);
    do {
      if ((st = dev_alloc(ctx, &part, (size_t)nblk * n_out)) != M3B_OK) break;
      if ((st = dev_alloc(ctx, &dG, (size_t)n_out)) != M3B_OK) break;
      const size_t smem = (size_t)q * AL_LD * sizeof(double);
      cudaError_t e = cudaFuncSetAttribute(k_align_gram, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max<size_t>(smem, 1024));
      if (e == cudaSuccess) {
        k_align_gram<<<nblk, AL_THREADS, smem, ctx->stream>>>(dM, n, p, dX, n, nv, n, rpb, part);
        k_align_reduce<<<(n_out + 255) / 256, 256, 0, ctx->stream>>>(part, nblk, n_out, dG);
        count_launch(ctx, 2);
        e = cudaGetLastError();
      }
      if (e != cudaSuccess) {
        m3b_set_error(ctx, "m3b_align_residuals: %s", cudaGetErrorString(e));
        st = M3B_E_CUDA;
        break;
      }
      if ((st = comm_allreduce_f64(ctx, dG, n_out)) != M3B_OK) break;  // cells sharded: sums over all ranks
      if (cudaMemcpyAsync(G.data(), dG, n_out * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess ||
          cudaStreamSynchronize(ctx->stream) != cudaSuccess) {
        m3b_set_error(ctx, "m3b_align_residuals: copy of the Gram matrix failed");
        st = M3B_E_CUDA;
        break;
      }
      // ---- host: p x p solve, lm.fit semantics (limited pivoting: aliased columns -> NA -> 0) ----
      // G[a * q + b]: a < p rows; b < p: M^T M, b >= p: M^T X.  Equilibrate with d_a = 1 / ||m_a||.
      std::vector<long double> d(p, 0.0L), L((size_t)p * p, 0.0L);
      std::vector<int> used;
      for (int a = 0; a < p; ++a) d[a] = G[(size_t)a * q + a] > 0 ? 1.0L / sqrtl((long double)G[(size_t)a * q + a]) : 0.0L;
      auto g = [&](int a, int b) { return (long double)G[(size_t)a * q + b] * d[a] * d[b]; };
      for (int j = 0; j < p; ++j) {
        if (d[j] == 0.0L) continue;  // a zero column is aliased
        // Cholesky row of column j against the used columns: l_jk = (g_jk - sum_m l_jm l_km) / l_kk
        std::vector<long double> row(used.size(), 0.0L);
        long double rem = g(j, j);  // = 1
        for (size_t k = 0; k < used.size(); ++k) {
          long double s = g(j, used[k]);
          for (size_t m2 = 0; m2 < k; ++m2) s -= row[m2] * L[(size_t)used[k] * p + used[m2]];
          row[k] = s / L[(size_t)used[k] * p + used[k]];
          rem -= row[k] * row[k];
        }
        // rem = squared norm of the (unit-norm) column's residual against the used columns
        if (!(rem > 1e-14L)) continue;  // lm.fit: residual norm < 1e-7 x column norm => aliased
        for (size_t k = 0; k < used.size(); ++k) L[(size_t)j * p + used[k]] = row[k];
        L[(size_t)j * p + j] = sqrtl(rem);
        used.push_back(j);
      }
      const int u = (int)used.size();
      for (int c = 0; c < nv; ++c) {
        // solve (L L^T) z = D M^T x_c on the used columns; coefficient = d * z
        std::vector<long double> w(u), z(u);
        for (int k = 0; k < u; ++k) {
          long double s = (long double)G[(size_t)used[k] * q + p + c] * d[used[k]];
          for (int m2 = 0; m2 < k; ++m2) s -= L[(size_t)used[k] * p + used[m2]] * w[m2];
          w[k] = s / L[(size_t)used[k] * p + used[k]];
        }
        for (int k = u - 1; k >= 0; --k) {
          long double s = w[k];
          for (int m2 = k + 1; m2 < u; ++m2) s -= L[(size_t)used[m2] * p + used[k]] * z[m2];
          z[k] = s / L[(size_t)used[k] * p + used[k]];
        }
        for (int k = 0; k < u; ++k)
          if (used[k] >= 1) beta[(size_t)(used[k] - 1) * nv + c] = (double)(z[k] * d[used[k]]);  // intercept dropped
      }
      for (size_t k = 0; k < beta.size(); ++k) beta_out[k] = beta[k];
    } while (0);
  }
  if (st == M3B_OK) st = align_apply_and_download(ctx, dX, dM, X, n, nv, ldx, p, beta);
  cudaStreamSynchronize(ctx->stream);
  dev_free(dX);
  dev_free(dM);
  dev_free(part);
  dev_free(dG);
  return st;
}
