// K8 fast path: SVD of the restarted projected matrix B by its structure.
//
// After a thick restart irlba's B (work x work, upper triangular) is
//     [ diag(BS[:k])   res          ]      rows 0..k-1 : the k kept Ritz values + residual column
//     [      0         bidiagonal   ]      rows k..n-1 : the new Lanczos steps (n-k <= 7 of them)
// LAPACK (dgesdd in irlb.c, reached from R/utils.R:417) and the Jacobi kernel in small_svd.cu
// both treat B as a dense matrix.  Here the structure is used instead:
//   * the leading (k+1)x(k+1) block M = [diag(d) z; 0 a] is an ARROW matrix: M M^T =
//     diag(d^2, 0) + u u^T with u = (z, a), a rank-one update of a diagonal matrix.  Its
//     eigenvalues are the roots of the secular equation 1 + sum_i u_i^2/(p_i^2 - s^2) = 0, one
//     root strictly between each pair of consecutive poles -- every root is found
//     independently by one warp.  The singular vectors follow in closed form once u is
//     replaced by the vector u^ that makes the computed roots exact (Loewner's formula, the
//     Gu-Eisenstat device LAPACK's divide-and-conquer uses to get orthogonal vectors).
//   * every further row/column of the bidiagonal tail turns [S w; 0 a'] (S the current
//     singular values, w = beta * last row of the current left vectors) into another arrow,
//     solved the same way and folded in with two small GEMMs.
// Work per restart: (n-k) secular solves (O(n^2) each, fully parallel) + 2(n-k-1) n^3 GEMMs
// spread over the whole chip, instead of ~6 Jacobi sweeps of n^3 serialised on one SM.
// Any irregularity (non-descending or coincident poles, a zero weight, a root that does not
// converge, non-finite values) raises a flag and the Jacobi kernel redoes that restart.
#include <algorithm>
#include <utility>

#include "m3b_internal.h"

#define AS_THREADS 1024
#define AS_NMAX 256
#define AS_MAXIT 60

__device__ __forceinline__ double wsum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double wprod(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v *= __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// One arrow solve.  Poles pf[0..m-2] (descending, > 0) and pf[m-1] = 0; weights uf[0..m-1].
//   mode 0: pf = diag(B)[0..m-2], uf = (B[0..m-2, m-1], B[m-1, m-1])          (first block)
//   mode 1: pf = sig_in[0..m-2],  uf = (beta * Ycur[m-2, 0..m-2], alpha)      (tail merge)
// Weights below 8 eps max(|pf|,|uf|) are DEFLATED (LAPACK dlasd2's rule): their pole is a
// singular value with unit singular vectors e_i (converged Ritz values do this all the time);
// the secular equation is solved on the remaining ACTIVE poles only.
// Outputs: sig_out[m] (descending), Yp / Xp (m x m, ld n): left / right singular vectors of the
// arrow.  *flag |= reason on any irregularity (1 size, 2 poles/weights, 4 root, 8 Loewner).
__global__ void __launch_bounds__(AS_THREADS, 1)
k_arrow_solve(int mode, int n, const int* __restrict__ ctl_k, int tail_j, const double* __restrict__ B,
              const double* __restrict__ sig_in, const double* __restrict__ Ycur, double* __restrict__ sig_out,
              double* __restrict__ Yp, double* __restrict__ Xp, int* __restrict__ flag) {
  __shared__ double pf[AS_NMAX], uf[AS_NMAX];                 // full problem
  __shared__ double p[AS_NMAX], u[AS_NMAX], mu[AS_NMAX], sg[AS_NMAX], uh[AS_NMAX];  // active problem
  __shared__ int org[AS_NMAX], act[AS_NMAX], defl[AS_NMAX], rnk[AS_NMAX];
  __shared__ int s_bad, s_ma;
  __shared__ double s_tol;
  if (mode == 1 && *flag) return;  // an earlier solve of this restart already gave up
  const int k = ctl_k[0];
  const int m = (mode == 0) ? k + 1 : tail_j + 1;  // size of this arrow
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
  if (tid == 0) s_bad = 0;
  if (m > n || m < 2 || m > AS_NMAX) {
    if (tid == 0) atomicOr(flag, 1);  // reason 1: size
    return;
  }
  for (int i = tid; i < m; i += blockDim.x) {
    double pi, ui;
    if (mode == 0) {
      pi = (i < m - 1) ? B[(size_t)i * n + i] : 0.0;
      ui = B[(size_t)(m - 1) * n + i];  // column m-1 = (z, alpha)
    } else {
      const int j = tail_j;  // new row/column index (m - 1)
      pi = (i < m - 1) ? sig_in[i] : 0.0;
      ui = (i < m - 1) ? B[(size_t)j * n + (j - 1)] * Ycur[(size_t)i * n + (m - 2)] : B[(size_t)j * n + j];
    }
    pf[i] = pi;
    uf[i] = ui;
  }
  __syncthreads();
  // deflation + regularity checks (m <= 256: one thread walks the list)
  if (tid == 0) {
    double umax = 0.0;
    bool bad = false;
    for (int i = 0; i < m; ++i) {
      umax = fmax(umax, fabs(uf[i]));
      bad = bad || !isfinite(pf[i]) || !isfinite(uf[i]);
      if (i < m - 1) bad = bad || !(pf[i] > 0.0) || !(pf[i] > pf[i + 1]);
    }
    const double tol = 8.0 * 2.220446049250313e-16 * fmax(pf[0], umax);
    int ma = 0, nd = 0;
    for (int i = 0; i < m; ++i) {
      if (fabs(uf[i]) > tol) act[ma++] = i;
      else defl[nd++] = i;
    }
    if (ma < 1 || act[ma - 1] != m - 1) bad = true;  // alpha itself negligible: singular B
    for (int r = 0; r + 1 < ma && !bad; ++r)
      if (!(pf[act[r]] - pf[act[r + 1]] > tol)) bad = true;  // coincident poles
    s_ma = ma;
    s_tol = tol;
    if (bad) s_bad = 1;
  }
  __syncthreads();
  if (s_bad) {
    if (tid == 0) atomicOr(flag, 2);  // reason 2: irregular poles / weights
    return;
  }
  const int ma = s_ma;
  for (int r = tid; r < ma; r += blockDim.x) {
    p[r] = pf[act[r]];
    u[r] = uf[act[r]];
  }
  __syncthreads();
  double un2 = 0.0;
  for (int i = lane; i < ma; i += 32) un2 += u[i] * u[i];
  un2 = wsum(un2);
  const double unorm = sqrt(un2);
  // ---- phase 1: one warp per root of the active problem ----
  // The root r lies in (p_r, p_{r-1}).  h(x) = f * (s - p_r) * (p_{r-1} - s) has no pole in
  // that interval, is negative at the lower pole and positive at the upper one; it is solved
  // with the Illinois variant of regula falsi in the offset x = s - p_o from the nearer pole o
  // (all pole differences are formed as (p_i - p_o) - x, never as p_i - s).
  for (int r = warp; r < ma; r += nwarps) {
    int o, up;
    double lo, hi;
    if (r == 0) {
      o = 0;
      up = -1;
      lo = 0.0;
      hi = unorm * (1.0 + 1e-12) + 1e-300;
    } else {
      up = r - 1;
      const double mid = 0.5 * (p[r - 1] - p[r]);
      double f = 0.0;
      for (int i = lane; i < ma; i += 32) f += u[i] * u[i] / (((p[i] - p[r]) - mid) * ((p[i] + p[r]) + mid));
      f = 1.0 + wsum(f);
      if (f > 0.0) {
        o = r;
        lo = 0.0;
        hi = mid;
      } else {
        o = r - 1;
        lo = -mid;
        hi = 0.0;
      }
    }
    const double po = p[o], pr = p[r], ur2 = u[r] * u[r];
    const double pu = (up >= 0) ? p[up] : 0.0, uu2 = (up >= 0) ? u[up] * u[up] : 0.0;
    auto hfun = [&](double x) -> double {
      double psi = 0.0;
      for (int i = lane; i < ma; i += 32)
        if (i != r && i != up) psi += u[i] * u[i] / (((p[i] - po) - x) * ((p[i] + po) + x));
      psi = 1.0 + wsum(psi);
      const double sv = po + x;
      const double dl = (po - pr) + x;
      double L, tl;
      if (pr > 0.0) {
        L = dl;
        tl = -ur2 / (pr + sv);
      } else {
        L = sv * sv;
        tl = -ur2;
      }
      if (up >= 0) {
        const double du = (pu - po) - x;
        return du * L * psi + du * tl + uu2 * L / (pu + sv);
      }
      return L * psi + tl;
    };
    double xa = lo, xb = hi, ha = hfun(xa), hb = hfun(xb), x = xa;
    bool ok = true;
    if (ha == 0.0) x = xa;
    else if (hb == 0.0) x = xb;
    else if (!(ha < 0.0 && hb > 0.0)) ok = false;
    else {
      int side = 0;
      bool conv = false;
      for (int it = 0; it < 100; ++it) {
        double xn = (xa * hb - xb * ha) / (hb - ha);
        if (!(xn > xa && xn < xb)) xn = 0.5 * (xa + xb);
        const double hn = hfun(xn);
        x = xn;
        if (hn == 0.0) {
          conv = true;
          break;
        }
        if (hn < 0.0) {
          xa = xn;
          ha = hn;
          if (side == -1) hb *= 0.5;
          side = -1;
        } else {
          xb = xn;
          hb = hn;
          if (side == 1) ha *= 0.5;
          side = 1;
        }
        if (fabs(xb - xa) <= 4.0 * 2.220446049250313e-16 * fmax(fabs(xa), fabs(xb))) {
          conv = true;
          break;
        }
      }
      if (!conv && !(fabs(xb - xa) <= 1e-10 * fmax(fabs(xa), fabs(xb)))) ok = false;
    }
    if (lane == 0) {
      org[r] = o;
      mu[r] = x;
      sg[r] = po + x;
      if (!ok || !isfinite(x) || x == 0.0) s_bad = 1;
    }
  }
  __syncthreads();
  if (s_bad) {
    if (tid == 0) atomicOr(flag, 4);  // reason 4: a secular root did not converge
    return;
  }
  // ---- phase 2: Loewner weights  uh_i^2 = prod_j (s_j^2 - p_i^2) / prod_{j != i} (p_j^2 - p_i^2) ----
  for (int i = warp; i < ma; i += nwarps) {
    double pr = 1.0;
    for (int j = lane; j < ma; j += 32) {
      const double num = ((p[org[j]] - p[i]) + mu[j]) * (sg[j] + p[i]);
      pr *= (j == i) ? num : num / ((p[j] - p[i]) * (p[j] + p[i]));
    }
    pr = wprod(pr);
    if (lane == 0) {
      uh[i] = copysign(sqrt(fabs(pr)), u[i]);
      if (!isfinite(pr) || pr == 0.0) s_bad = 1;
    }
  }
  __syncthreads();
  if (s_bad) {
    if (tid == 0) atomicOr(flag, 8);  // reason 8: Loewner product degenerate
    return;
  }
  // ---- phase 3: output order.  entry e < ma: active root e; e >= ma: deflated pole defl[e - ma] ----
  for (int e = tid; e < m; e += blockDim.x) {
    const double me = (e < ma) ? sg[e] : pf[defl[e - ma]];
    int rk = 0;
    for (int q = 0; q < m; ++q) {
      const double ot = (q < ma) ? sg[q] : pf[defl[q - ma]];
      rk += (ot > me || (ot == me && q < e)) ? 1 : 0;
    }
    rnk[e] = rk;
  }
  __syncthreads();
  // ---- phase 4: vectors.  y_j(i) = uh_i / ((p_i - s_j)(p_i + s_j)); x_j = M^T y_j / s_j ----
  for (int e = warp; e < m; e += nwarps) {
    const int dst = rnk[e];
    double* yc = Yp + (size_t)dst * n;
    double* xc = Xp + (size_t)dst * n;
    for (int i = lane; i < m; i += 32) {
      yc[i] = 0.0;
      xc[i] = 0.0;
    }
    __syncwarp();
    if (e >= ma) {  // deflated: unit vectors
      const int i0 = defl[e - ma];
      if (lane == 0) {
        yc[i0] = 1.0;
        xc[i0] = 1.0;
        sig_out[dst] = pf[i0];
      }
      continue;
    }
    const double po = p[org[e]], mj = mu[e], sj = sg[e];
    double ny = 0.0, dot = 0.0;
    for (int i = lane; i < ma; i += 32) {
      const double y = uh[i] / (((p[i] - po) - mj) * (p[i] + sj));
      ny += y * y;
      dot += uh[i] * y;
    }
    ny = wsum(ny);
    dot = wsum(dot);
    const double iy = 1.0 / sqrt(ny);
    double nx = 0.0;
    for (int i = lane; i < ma; i += 32) {
      const double y = uh[i] / (((p[i] - po) - mj) * (p[i] + sj));
      const double xv = (i < ma - 1) ? p[i] * y : dot;  // last active index is the pole 0 (alpha row)
      nx += xv * xv;
    }
    nx = wsum(nx);
    const double ix = 1.0 / sqrt(nx);
    for (int i = lane; i < ma; i += 32) {
      const double y = uh[i] / (((p[i] - po) - mj) * (p[i] + sj));
      const double xv = (i < ma - 1) ? p[i] * y : dot;
      yc[act[i]] = y * iy;
      xc[act[i]] = xv * ix;
    }
    if (lane == 0) sig_out[dst] = sj;
  }
}

// Fold an arrow's vectors into the accumulated ones:  Ynew[0:m-1, :] = Ycur[0:m-1, 0:m-1] * Yp[0:m-1, :],
// Ynew[m-1, :] = Yp[m-1, :]  (and the same for X).  All matrices col-major with ld n; m = tail_j + 1.
// grid = (ceil(m/16), ceil(m/16), 2); block = 16 x 16; blockIdx.z picks Y or X.
__global__ void __launch_bounds__(256) k_arrow_fold(int n, int tail_j, const double* __restrict__ Ycur,
                                                    const double* __restrict__ Xcur, const double* __restrict__ Yp,
                                                    const double* __restrict__ Xp, double* __restrict__ Ynew,
                                                    double* __restrict__ Xnew, const int* __restrict__ flag) {
  if (*flag) return;
  const int m = tail_j + 1, mm = m - 1;
  const double* A = blockIdx.z ? Xcur : Ycur;
  const double* P = blockIdx.z ? Xp : Yp;
  double* C = blockIdx.z ? Xnew : Ynew;
  __shared__ double As[16][17], Ps[16][17];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int r = blockIdx.x * 16 + tx, c = blockIdx.y * 16 + ty;
  double acc = 0.0;
  for (int k0 = 0; k0 < mm; k0 += 16) {
    // As[kk][row], Ps[col][kk]
    const int ka = k0 + ty, kb = k0 + tx;
    As[ty][tx] = (r < mm && ka < mm) ? A[(size_t)ka * n + r] : 0.0;
    Ps[ty][tx] = (c < m && kb < mm) ? P[(size_t)c * n + kb] : 0.0;
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) acc += As[kk][tx] * Ps[ty][kk];
    __syncthreads();
  }
  if (c < m) {
    if (r < mm) C[(size_t)c * n + r] = acc;
    else if (r == mm) C[(size_t)c * n + r] = P[(size_t)c * n + mm];
  }
}

// same sign convention as the Jacobi kernel: largest |u| component of every left vector positive
__global__ void k_svd_signfix(int n, double* __restrict__ U, double* __restrict__ V, const int* __restrict__ flag) {
  if (*flag) return;
  const int lane = threadIdx.x & 31;
  const int j = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (j >= n) return;
  double best = 0.0;
  int bi = n;
  for (int i = lane; i < n; i += 32) {
    const double x = fabs(U[(size_t)j * n + i]);
    if (x > best) {
      best = x;
      bi = i;
    }
  }
  for (int o = 16; o > 0; o >>= 1) {
    const double ob = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ob > best || (ob == best && oi < bi)) {
      best = ob;
      bi = oi;
    }
  }
  if (bi < n && U[(size_t)j * n + bi] < 0.0)
    for (int i = lane; i < n; i += 32) {
      U[(size_t)j * n + i] = -U[(size_t)j * n + i];
      V[(size_t)j * n + i] = -V[(size_t)j * n + i];
    }
}

// Structured SVD of the restarted B (k = ctl[CT_K] kept Ritz values, known to the host as k_host).
// Buffers: U/V (n x n outputs), S (n), scratch >= 4 n^2 doubles + 2 n, flag = device int (zeroed here).
// On return *flag != 0 means "not done: run the Jacobi kernel".
int launch_arrow_svd(m3b_ctx* ctx, int n, int k_host, const int* ctl_k, const double* B, double* U, double* S,
                     double* V, double* scratch, int* flag) {
  if (n > AS_NMAX || k_host < 1 || k_host + 1 > n) return M3B_E_DIMS;
  double* Ya = scratch;                       // ping
  double* Xa = scratch + (size_t)n * n;
  double* Yp = scratch + 2 * (size_t)n * n;   // per-arrow vectors
  double* Xp = scratch + 3 * (size_t)n * n;
  double* sa = scratch + 4 * (size_t)n * n;   // sigma ping
  double* sb = sa + n;                        // sigma pong
  CK(ctx, cudaMemsetAsync(flag, 0, sizeof(int), ctx->stream));
  const int ntail = n - (k_host + 1);
  // the accumulated vectors ping-pong between (U,V) and (Ya,Xa) so that the last fold lands in U,V
  double *Yc, *Xc, *Yn, *Xn;
  if (ntail % 2 == 0) {
    Yc = U;
    Xc = V;
    Yn = Ya;
    Xn = Xa;
  } else {
    Yc = Ya;
    Xc = Xa;
    Yn = U;
    Xn = V;
  }
  double* s_cur = (ntail == 0) ? S : sa;
  k_arrow_solve<<<1, AS_THREADS, 0, ctx->stream>>>(0, n, ctl_k, 0, B, nullptr, nullptr, s_cur, Yc, Xc, flag);
  count_launch(ctx);
  for (int t = 0; t < ntail; ++t) {
    const int j = k_host + 1 + t;  // new row / column
    double* s_next = (t == ntail - 1) ? S : (s_cur == sa ? sb : sa);
    k_arrow_solve<<<1, AS_THREADS, 0, ctx->stream>>>(1, n, ctl_k, j, B, s_cur, Yc, s_next, Yp, Xp, flag);
    dim3 grid((j + 1 + 15) / 16, (j + 1 + 15) / 16, 2);
    k_arrow_fold<<<grid, 256, 0, ctx->stream>>>(n, j, Yc, Xc, Yp, Xp, Yn, Xn, flag);
    count_launch(ctx, 2);
    std::swap(Yc, Yn);
    std::swap(Xc, Xn);
    s_cur = s_next;
  }
  k_svd_signfix<<<(n * 32 + 255) / 256, 256, 0, ctx->stream>>>(n, U, V, flag);
  count_launch(ctx);
  CK(ctx, cudaGetLastError());
  return M3B_OK;
}
