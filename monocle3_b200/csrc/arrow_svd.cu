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
//
// ONE secular solve per restart (mode 2, the default).  With T the (n-k)^2 bidiagonal tail
// block (T[0,0] = a) and T' = T with that corner removed,
//     B B^T = blockdiag(D^2, T' T'^T) + u u^T,   u = (z, a e_1):
// the tail couples to the kept Ritz values through its first column only.  T' is at most 7x7:
// its SVD T' = P S' R^T comes from the Jacobi kernel in a few microseconds (one singular value is
// exactly 0, T' has a zero column), and in the basis blockdiag(I, P) the whole of B B^T is
// diag(d^2, s'^2) + u~ u~^T with u~ = (z, a P^T e_1): a single rank-one update, one secular
// solve of size n over the merged poles.  Left vectors U = blockdiag(I, P) Y (Loewner), right
// vectors V = B^T U S^-1.  The sequential tail merges above (kept as M3B_SVD_TAILMERGE=1) cost
// n-k dependent solves; this costs one.
// Any irregularity (non-descending or coincident poles, a zero weight, a root that does not
// converge, non-finite values) raises a flag and the Jacobi kernel redoes that restart.
#include <algorithm>
#include <utility>

#include "m3b_internal.h"

#define AS_NMAX 256

int launch_jacobi_svd(m3b_ctx* ctx, int n, const double* B, double* U, double* s, double* V, double* scratch, int* info,
                      const int* only_if);

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

// ---------------------------------------------------------------------------------------------
// One arrow solve = two multi-CTA kernels (the first version ran everything in ONE block and was
// instruction-bound on a single SM: ~190 us per solve, profiles/):
//   k_arrow_roots   : 4 warps per CTA, one secular root per warp  -> org / mu / sg (global)
//   k_arrow_vectors : 4 warps per CTA, one output column per warp (Loewner weights and the output
//                     order are recomputed by every CTA: O(m^2) and cheaper than another launch)
// Both start by rebuilding the problem (poles, weights, deflation) from B in shared memory with
// the same deterministic code, so nothing but the roots has to travel between them.
//
// Poles pf[0..m-2] (descending, > 0) and pf[m-1] = 0; weights uf[0..m-1].
//   mode 0: pf = diag(B)[0..m-2], uf = (B[0..m-2, m-1], B[m-1, m-1])          (first block)
//   mode 1: pf = sig_in[0..m-2],  uf = (beta * Ycur[m-2, 0..m-2], alpha)      (tail merge)
// Weights below 8 eps max(|pf|,|uf|) are DEFLATED (LAPACK dlasd2's rule): their pole is a
// singular value with unit singular vectors e_i (converged Ritz values do this all the time);
// the secular equation is solved on the remaining ACTIVE poles only.
// *flag |= reason on any irregularity (1 size, 2 poles/weights, 4 root, 8 Loewner).
// ---------------------------------------------------------------------------------------------
#define AR_THREADS 128
#define AR_WARPS 4

struct ArrowProblem {  // shared-memory image of one arrow problem
  double pf[AS_NMAX], uf[AS_NMAX];  // full
  double p[AS_NMAX], u2[AS_NMAX];   // active poles, squared active weights
  int act[AS_NMAX], defl[AS_NMAX];
  int src[AS_NMAX];  // mode 2: index in the concatenated (D, tail) basis of every merged pole
  int m, ma, bad;
};

// returns false (block-uniform) if the problem is irregular
__device__ __forceinline__ bool arrow_setup(ArrowProblem& P, int mode, int n, int k, int tail_j,
                                            const double* __restrict__ B, const double* __restrict__ sig_in,
                                            const double* __restrict__ Ycur) {
  const int tid = threadIdx.x;
  const int m = (mode == 2) ? n : ((mode == 0) ? k + 1 : tail_j + 1);
  if (m > n || m < 2 || m > AS_NMAX) return false;
  if (mode == 2) {
    // merged poles: the k kept Ritz values (descending) and the t-1 non-zero singular values of T'
    // (sig_in, descending), then the zero pole; Ycur = P (t x t, col-major, ld t)
    const int t = n - k;
    if (k < 1 || t < 2 || t > 32) return false;
    const double alpha = B[(size_t)k * n + k];
    for (int q = tid; q < n; q += blockDim.x) {
      double pole, wt;
      int pos;
      if (q < k) {
        pole = B[(size_t)q * n + q];
        wt = B[(size_t)k * n + q];
        int c = 0;
        for (int j = 0; j < t - 1; ++j) c += (sig_in[j] > pole) ? 1 : 0;
        pos = q + c;
      } else {
        const int j = q - k;
        wt = alpha * Ycur[(size_t)j * t];
        if (j < t - 1) {
          pole = sig_in[j];
          int c = 0;
          for (int qq = 0; qq < k; ++qq) c += (B[(size_t)qq * n + qq] >= pole) ? 1 : 0;
          pos = j + c;
        } else {
          pole = 0.0;
          pos = n - 1;
        }
      }
      P.pf[pos] = pole;
      P.uf[pos] = wt;
      P.src[pos] = q;
    }
  }
  for (int i = tid; i < m && mode != 2; i += blockDim.x) {
    double pi, ui;
    if (mode == 0) {
      pi = (i < m - 1) ? B[(size_t)i * n + i] : 0.0;
      ui = B[(size_t)(m - 1) * n + i];  // column m-1 = (z, alpha)
    } else {
      const int j = tail_j;  // new row/column index (m - 1)
      pi = (i < m - 1) ? sig_in[i] : 0.0;
      ui = (i < m - 1) ? B[(size_t)j * n + (j - 1)] * Ycur[(size_t)i * n + (m - 2)] : B[(size_t)j * n + j];
    }
    P.pf[i] = pi;
    P.uf[i] = ui;
  }
  __syncthreads();
  // deflation + regularity checks, by warp 0 (ballot / popc compaction keeps the index order)
  if (tid == 0) P.bad = 0;
  __syncthreads();
  if (tid < 32) {
    const int lane = tid;
    double umax = 0.0;
    bool bad = false;
    for (int i = lane; i < m; i += 32) {
      umax = fmax(umax, fabs(P.uf[i]));
      bad = bad || !isfinite(P.pf[i]) || !isfinite(P.uf[i]);
      if (i < m - 1) bad = bad || !(P.pf[i] > 0.0) || !(P.pf[i] > P.pf[i + 1]);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) umax = fmax(umax, __shfl_xor_sync(0xffffffffu, umax, o));
    const double tol = 8.0 * 2.220446049250313e-16 * fmax(P.pf[0], umax);
    int ma = 0, nd = 0;
    for (int i0 = 0; i0 < m; i0 += 32) {
      const int i = i0 + lane;
      const bool in = i < m;
      const bool a = in && fabs(P.uf[i]) > tol;
      const unsigned ba = __ballot_sync(0xffffffffu, a), bd = __ballot_sync(0xffffffffu, in && !a);
      const unsigned lt = (1u << lane) - 1u;
      if (a) P.act[ma + __popc(ba & lt)] = i;
      if (in && !a) P.defl[nd + __popc(bd & lt)] = i;
      ma += __popc(ba);
      nd += __popc(bd);
    }
    __syncwarp();
    if (ma < 1 || P.act[ma - 1] != m - 1) bad = true;  // alpha itself negligible: singular B
    for (int r = lane; r + 1 < ma; r += 32)
      if (!(P.pf[P.act[r]] - P.pf[P.act[r + 1]] > tol)) bad = true;  // coincident poles
    if (__any_sync(0xffffffffu, bad) && lane == 0) P.bad = 1;
    if (lane == 0) {
      P.m = m;
      P.ma = ma;
    }
  }
  __syncthreads();
  if (P.bad) return false;
  for (int r = tid; r < P.ma; r += blockDim.x) {
    P.p[r] = P.pf[P.act[r]];
    const double w = P.uf[P.act[r]];
    P.u2[r] = w * w;
  }
  __syncthreads();
  return true;
}

__global__ void __launch_bounds__(AR_THREADS)
k_arrow_roots(int mode, int n, const int* __restrict__ ctl_k, int tail_j, const double* __restrict__ B,
              const double* __restrict__ sig_in, const double* __restrict__ Ycur, int* __restrict__ org_g,
              double* __restrict__ mu_g, double* __restrict__ sg_g, double* __restrict__ uh_g,
              unsigned int* __restrict__ done_ctr, int* __restrict__ flag) {
  __shared__ ArrowProblem P;
  __shared__ int s_last;
  if (mode == 1 && *flag) return;  // an earlier solve of this restart already gave up
  if (!arrow_setup(P, mode, n, ctl_k[0], tail_j, B, sig_in, Ycur)) {
    if (threadIdx.x == 0) atomicOr(flag, 2);
    return;
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int ma = P.ma;
  const int r = blockIdx.x * AR_WARPS + warp;
  const double* p = P.p;
  const double* u2 = P.u2;
  if (r < ma) {
  // The root r lies in (p_r, p_{r-1}).  h(x) = f * (s - p_r) * (p_{r-1} - s) has no pole in
  // that interval, is negative at the lower pole and positive at the upper one; it is solved
  // with the Illinois variant of regula falsi in the offset x = s - p_o from the nearer pole o
  // (all pole differences are formed as (p_i - p_o) - x, never as p_i - s).
  int o, up;
  double lo, hi;
  if (r == 0) {
    double un2 = 0.0;
    for (int i = lane; i < ma; i += 32) un2 += u2[i];
    un2 = wsum(un2);
    o = 0;
    up = -1;
    lo = 0.0;
    hi = sqrt(un2) * (1.0 + 1e-12) + 1e-300;
  } else {
    up = r - 1;
    const double mid = 0.5 * (p[r - 1] - p[r]);
    double f = 0.0;
    for (int i0 = 0; i0 < ma; i0 += 128) {
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int i = i0 + lane + 32 * q;
        if (i < ma) f += u2[i] / (((p[i] - p[r]) - mid) * ((p[i] + p[r]) + mid));
      }
    }
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
  const double po = p[o], pr = p[r], ur2 = u2[r];
  const double pu = (up >= 0) ? p[up] : 0.0, uu2 = (up >= 0) ? u2[up] : 0.0;
  auto hfun = [&](double x) -> double {
    double psi = 0.0;
    for (int i0 = 0; i0 < ma; i0 += 128) {
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int i = i0 + lane + 32 * q;
        if (i < ma && i != r && i != up) psi += u2[i] / (((p[i] - po) - x) * ((p[i] + po) + x));
      }
    }
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
    // probe the dominant-pole estimate  x ~ u_o^2 / (2 p_o c0)  (c0 = the rest of the secular
    // function at the pole): converged Ritz values sit 1e-20 gaps from their pole
    if (po > 0.0) {
      double c0 = 0.0;
      for (int i0 = 0; i0 < ma; i0 += 128) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int i = i0 + lane + 32 * q;
          if (i < ma && i != o) c0 += u2[i] / ((p[i] - po) * (p[i] + po));
        }
      }
      c0 = 1.0 + wsum(c0);
      const double xg = u2[o] / (2.0 * po * c0);
      if (isfinite(xg) && xg > xa && xg < xb) {
        const double hg = hfun(xg);
        if (hg < 0.0) {
          xa = xg;
          ha = hg;
        } else if (hg > 0.0) {
          xb = xg;
          hb = hg;
        } else {
          xa = xb = xg;
        }
      }
    }
    if (xa == xb) x = xa;
    else {
      // Dekker's method with Brent's minimal step: bq = best iterate, aq = end of the bracket with
      // the opposite sign, cq = previous iterate.  The secant runs through the two LATEST
      // iterates, so a stale far end of the bracket does not slow it down (regula falsi /
      // Illinois crept for ~50 iterations on roots that sit 1e-20 gaps from a pole).
      double aq, hq_a, bq, hq_b;
      if (fabs(ha) < fabs(hb)) {
        aq = xb; hq_a = hb; bq = xa; hq_b = ha;
      } else {
        aq = xa; hq_a = ha; bq = xb; hq_b = hb;
      }
      double cq = aq, hq_c = hq_a;
      bool conv = false;
      for (int it = 0; it < 100; ++it) {
        if (hq_b == 0.0) {
          conv = true;
          break;
        }
        const double tol1 = 2.0 * 2.220446049250313e-16 * fabs(bq);
        const double mid = 0.5 * (aq - bq);
        if (fabs(mid) <= tol1) {
          conv = true;
          break;
        }
        double st = (hq_b != hq_c) ? -hq_b * (bq - cq) / (hq_b - hq_c) : mid;
        const double ratio = st / mid;
        if (!(ratio > 0.0 && ratio < 1.0)) st = mid;  // keep the step inside (b, b + mid)
        if (fabs(st) < tol1) st = (mid > 0.0) ? tol1 : -tol1;
        cq = bq;
        hq_c = hq_b;
        bq = bq + st;
        hq_b = hfun(bq);
        if ((hq_b < 0.0) == (hq_a < 0.0)) {
          aq = cq;
          hq_a = hq_c;
        }
        if (fabs(hq_a) < fabs(hq_b)) {
          const double ta = aq, tha = hq_a;
          aq = bq; hq_a = hq_b;
          bq = ta; hq_b = tha;
          cq = aq; hq_c = hq_a;
        }
      }
      x = bq;
      if (!conv) ok = false;
    }
  }
  if (lane == 0) {
    org_g[r] = o;
    mu_g[r] = x;
    sg_g[r] = po + x;
    if (!ok || !isfinite(x) || x == 0.0) atomicOr(flag, 4);  // reason 4: a secular root did not converge
  }
  }  // r < ma
  // ---- the last CTA to finish computes the Loewner weights for everyone (one thread per i):
  //   uh_i^2 = prod_j (s_j^2 - p_i^2) / prod_{j != i} (p_j^2 - p_i^2),  one division per 4 factors
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned int d = atomicInc(done_ctr, gridDim.x - 1);
    s_last = (d == gridDim.x - 1) ? 1 : 0;
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  if (*((volatile int*)flag)) return;
  for (int i = threadIdx.x; i < ma; i += blockDim.x) {
    const double pi = p[i];
    double pr = 1.0;
    for (int j0 = 0; j0 < ma; j0 += 4) {
      double num = 1.0, den = 1.0;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int j = j0 + q;
        if (j < ma) {
          const int oj = __ldcg(org_g + j);
          num *= ((p[oj] - pi) + __ldcg(mu_g + j)) * (__ldcg(sg_g + j) + pi);
          if (j != i) den *= (p[j] - pi) * (p[j] + pi);
        }
      }
      pr *= num / den;
    }
    uh_g[i] = copysign(sqrt(fabs(pr)), P.uf[P.act[i]]);
    if (!isfinite(pr) || pr == 0.0) atomicOr(flag, 8);  // reason 8: Loewner product degenerate
  }
}

// Outputs: sig_out[m] (descending), Yp / Xp (m x m, ld n): left / right singular vectors of the arrow.
__global__ void __launch_bounds__(AR_THREADS)
k_arrow_vectors(int mode, int n, const int* __restrict__ ctl_k, int tail_j, const double* __restrict__ B,
                const double* __restrict__ sig_in, const double* __restrict__ Ycur, const int* __restrict__ org_g,
                const double* __restrict__ mu_g, const double* __restrict__ sg_g, const double* __restrict__ uh_g,
                double* __restrict__ sig_out, double* __restrict__ Yp, double* __restrict__ Xp,
                int* __restrict__ flag) {
  __shared__ ArrowProblem P;
  __shared__ double mu[AS_NMAX], sg[AS_NMAX], uh[AS_NMAX];
  __shared__ int org[AS_NMAX];
  if (*flag) return;
  if (!arrow_setup(P, mode, n, ctl_k[0], tail_j, B, sig_in, Ycur)) return;  // (roots already flagged it)
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int m = P.m, ma = P.ma;
  const double* p = P.p;
  if (mode == 2 && blockIdx.x == 0) {  // the merge order, for k_tail_finish (stored behind the roots' org array)
    int* src_g = const_cast<int*>(org_g) + n;
    for (int r = tid; r < m; r += blockDim.x) src_g[r] = P.src[r];
  }
  for (int r = tid; r < ma; r += blockDim.x) {
    org[r] = org_g[r];
    mu[r] = mu_g[r];
    sg[r] = sg_g[r];
    uh[r] = uh_g[r];
  }
  __syncthreads();
  // this warp's output entry.  entry e < ma: active root e; e >= ma: deflated pole defl[e - ma]
  const int e = blockIdx.x * AR_WARPS + warp;
  if (e >= m) return;
  const double me = (e < ma) ? sg[e] : P.pf[P.defl[e - ma]];
  int rk = 0;
  for (int q = lane; q < m; q += 32) {
    const double ot = (q < ma) ? sg[q] : P.pf[P.defl[q - ma]];
    rk += (ot > me || (ot == me && q < e)) ? 1 : 0;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) rk += __shfl_xor_sync(0xffffffffu, rk, o);
  const int dst = rk;
  double* yc = Yp + (size_t)dst * n;
  double* xc = Xp + (size_t)dst * n;
  for (int i = lane; i < m; i += 32) {
    yc[i] = 0.0;
    xc[i] = 0.0;
  }
  __syncwarp();
  if (e >= ma) {  // deflated: unit vectors
    const int i0 = P.defl[e - ma];
    if (lane == 0) {
      yc[i0] = 1.0;
      xc[i0] = 1.0;
      sig_out[dst] = P.pf[i0];
    }
    return;
  }
  // y_j(i) = uh_i / ((p_i - s_j)(p_i + s_j)); x_j = M^T y_j / s_j
  const double po = p[org[e]], mj = mu[e], sj = sg[e];
  double yv[AS_NMAX / 32];
  double ny = 0.0, dot = 0.0;
#pragma unroll
  for (int q = 0; q < AS_NMAX / 32; ++q) {
    const int i = lane + 32 * q;
    yv[q] = (i < ma) ? uh[i] / (((p[i] - po) - mj) * (p[i] + sj)) : 0.0;
    ny += yv[q] * yv[q];
    dot += (i < ma) ? uh[i] * yv[q] : 0.0;
  }
  ny = wsum(ny);
  dot = wsum(dot);
  const double iy = 1.0 / sqrt(ny);
  double nx = 0.0;
#pragma unroll
  for (int q = 0; q < AS_NMAX / 32; ++q) {
    const int i = lane + 32 * q;
    const double xv = (i < ma - 1) ? p[i] * yv[q] : ((i == ma - 1) ? dot : 0.0);  // last active = pole 0 (alpha row)
    nx += xv * xv;
  }
  nx = wsum(nx);
  const double ix = 1.0 / sqrt(nx);
#pragma unroll
  for (int q = 0; q < AS_NMAX / 32; ++q) {
    const int i = lane + 32 * q;
    if (i < ma) {
      const double xv = (i < ma - 1) ? p[i] * yv[q] : dot;
      yc[P.act[i]] = yv[q] * iy;
      xc[P.act[i]] = xv * ix;
    }
  }
  if (lane == 0) sig_out[dst] = sj;
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
// SVD of the tail block T' (t <= 8) in ONE warp: Hestenes Jacobi on the ROWS of T' (columns of
// A = T'^T): A Q = W with orthogonal columns, so T' = Q diag(|w_j|) (..)^T and the accumulated
// rotations Q are the left singular vectors -- including the null vector that belongs to the zero
// column of T'.  8 lanes per column pair (one per row), 4 disjoint pairs per step, 7 steps per sweep
// (round-robin).  Outputs: Pu (t x t, col-major, ld t; columns ordered by descending singular
// value), sT (t).
__global__ void __launch_bounds__(32) k_tail_svd8(int n, const int* __restrict__ ctl_k, const double* __restrict__ B,
                                                  double* __restrict__ Pu, double* __restrict__ sT,
                                                  int* __restrict__ flag) {
  __shared__ double A[8][9], Q[8][9];  // [column][row]
  __shared__ double nrm[8];
  const int lane = threadIdx.x, k = ctl_k[0], t = n - k;
  if (t < 2 || t > 8) {
    if (lane == 0) atomicOr(flag, 1);
    return;
  }
  for (int e = lane; e < 64; e += 32) {
    const int c = e >> 3, r = e & 7;  // column c of A = row c of T':  A[c][r] = T'[c, r] = B[k+c, k+r]
    double v = 0.0;
    if (c < t && r < t && !(c == 0 && r == 0)) v = B[(size_t)(k + r) * n + (k + c)];
    A[c][r] = v;
    Q[c][r] = (c == r) ? 1.0 : 0.0;
  }
  __syncwarp();
  // columns whose norm has dropped to rounding level (the null direction) are left alone:
  // against them every inner product is noise and the sweeps would never settle
  double fro = 0.0;
  for (int e = lane; e < 64; e += 32) fro += A[e >> 3][e & 7] * A[e >> 3][e & 7];
  fro = wsum(fro);
  const double tiny2 = 1e-30 * fro;
  const int pr = lane >> 3, row = lane & 7;
  for (int sweep = 0; sweep < 40; ++sweep) {
    int rotated = 0;
    for (int st = 0; st < 7; ++st) {
      int i, j;
      if (pr == 0) {
        i = 7;
        j = st;
      } else {
        i = (st + pr) % 7;
        j = (st + 7 - pr) % 7;
      }
      if (i > j) {
        const int tmp = i;
        i = j;
        j = tmp;
      }
      const double ai = A[i][row], aj = A[j][row];
      double al = ai * ai, be = aj * aj, ga = ai * aj;
#pragma unroll
      for (int o = 4; o > 0; o >>= 1) {
        al += __shfl_xor_sync(0xffffffffu, al, o);
        be += __shfl_xor_sync(0xffffffffu, be, o);
        ga += __shfl_xor_sync(0xffffffffu, ga, o);
      }
      const bool rot = al > tiny2 && be > tiny2 && fabs(ga) > 1e-15 * sqrt(al * be);
      if (rot) {
        const double zeta = (be - al) / (2.0 * ga);
        const double tn = copysign(1.0, zeta) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
        const double c = 1.0 / sqrt(1.0 + tn * tn), sn = c * tn;
        const double qi = Q[i][row], qj = Q[j][row];
        A[i][row] = c * ai - sn * aj;
        A[j][row] = sn * ai + c * aj;
        Q[i][row] = c * qi - sn * qj;
        Q[j][row] = sn * qi + c * qj;
      }
      rotated |= __any_sync(0xffffffffu, rot) ? 1 : 0;
      __syncwarp();
    }
    if (!rotated) break;
  }
  // column norms, rank by descending norm (ties by index), write out the first t
  if (lane < 8) {
    double v = 0.0;
    for (int r = 0; r < 8; ++r) v += A[lane][r] * A[lane][r];
    nrm[lane] = sqrt(v);
  }
  __syncwarp();
  if (lane < t) {
    int rk = 0;
    for (int q = 0; q < t; ++q) rk += (nrm[q] > nrm[lane] || (nrm[q] == nrm[lane] && q < lane)) ? 1 : 0;
    sT[rk] = nrm[lane];
    for (int r = 0; r < t; ++r) Pu[(size_t)rk * t + r] = Q[lane][r];
  }
}

// T' (t x t, col-major, ld t): the bidiagonal tail block of B without its corner
__global__ void k_tail_prep(int n, const int* __restrict__ ctl_k, const double* __restrict__ B, double* __restrict__ Tm) {
  const int k = ctl_k[0], t = n - k;
  for (int e = threadIdx.x; e < t * t; e += blockDim.x) {
    const int r = e % t, c = e / t;
    Tm[e] = (r == 0 && c == 0) ? 0.0 : B[(size_t)(k + c) * n + (k + r)];
  }
}

// Column c of the outputs: U[:, c] = blockdiag(I, P) y_c (y_c un-merged through src),
// V[:, c] = B^T U[:, c] / s_c.  One block per column.
__global__ void __launch_bounds__(128) k_tail_finish(int n, const int* __restrict__ ctl_k, const double* __restrict__ B,
                                                     const double* __restrict__ Pu, const int* __restrict__ src_g,
                                                     const double* __restrict__ Ym, const double* __restrict__ S,
                                                     double* __restrict__ U, double* __restrict__ V,
                                                     int* __restrict__ flag) {
  __shared__ double ycat[AS_NMAX], ucol[AS_NMAX];
  __shared__ double sh[4];
  __shared__ int s_flip;
  if (*flag) return;
  const int k = ctl_k[0], t = n - k, c = blockIdx.x, tid = threadIdx.x;
  for (int pos = tid; pos < n; pos += blockDim.x) ycat[src_g[pos]] = Ym[(size_t)c * n + pos];
  __syncthreads();
  for (int q = tid; q < n; q += blockDim.x) {
    double v;
    if (q < k) {
      v = ycat[q];
    } else {
      v = 0.0;
      const int a = q - k;
      for (int j = 0; j < t; ++j) v += Pu[(size_t)j * t + a] * ycat[k + j];
    }
    ucol[q] = v;
    U[(size_t)c * n + q] = v;
  }
  __syncthreads();
  const double sc = S[c];
  double nn = 0.0;
  for (int i = tid; i < n; i += blockDim.x) {
    const double* bc = B + (size_t)i * n;  // column i of B = row i of B^T
    double a0 = 0.0, a1 = 0.0;
    int r = 0;
    for (; r + 1 < n; r += 2) {
      a0 += bc[r] * ucol[r];
      a1 += bc[r + 1] * ucol[r + 1];
    }
    if (r < n) a0 += bc[r] * ucol[r];
    const double v = (a0 + a1) / sc;
    V[(size_t)c * n + i] = v;
    nn += v * v;
  }
  // ||v|| must be 1: anything else means s_c was too small for B^T u / s (fall back to Jacobi)
  nn = wsum(nn);
  if ((tid & 31) == 0) sh[tid >> 5] = nn;
  __syncthreads();
  if (tid == 0) {
    const double tot = ((sh[0] + sh[1]) + sh[2]) + sh[3];
    if (!(fabs(tot - 1.0) < 1e-9)) atomicOr(flag, 16);
  }
  // sign convention of the Jacobi kernel (k_svd_signfix): largest |u| component positive
  if (tid < 32) {
    double best = 0.0;
    int bi = n;
    for (int i = tid; i < n; i += 32) {
      const double x = fabs(ucol[i]);
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
    if (tid == 0) s_flip = (bi < n && ucol[bi] < 0.0) ? 1 : 0;
  }
  __syncthreads();
  if (s_flip)
    for (int i = tid; i < n; i += blockDim.x) {
      U[(size_t)c * n + i] = -U[(size_t)c * n + i];
      V[(size_t)c * n + i] = -V[(size_t)c * n + i];
    }
}

// Buffers: U/V (n x n outputs), S (n), scratch >= 4 n^2 + 6 n doubles, flag = device int (zeroed here).
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
  CK(ctx, cudaMemsetAsync(scratch + 4 * (size_t)n * n + 5 * (size_t)n, 0, 2 * n * sizeof(double), ctx->stream));
  const int tt = n - k_host;
  if (!getenv("M3B_SVD_TAILMERGE") && tt >= 2 && tt <= 32 && 5 * tt * tt + tt + 2 <= n * n && 2 * n + 4 <= 2 * n * n) {
    // one secular solve over the merged poles (see the header); small buffers live in the Ya / Xa areas
    const int t = tt;
    double* Tm = Ya;           // t*t
    double* Pu = Tm + t * t;   // t*t   left vectors of T'
    double* Pv = Pu + t * t;   // t*t
    double* sT = Pv + t * t;   // t
    double* jscr = sT + t;     // 2*t*t
    int* jinfo = reinterpret_cast<int*>(jscr + 2 * t * t);
    double* mu_g = sb + n;
    double* sg_g = mu_g + n;
    double* uh_g = sg_g + n;
    int* org_g = reinterpret_cast<int*>(Xa);  // org[n], src[n], done counter
    unsigned int* done_ctr = reinterpret_cast<unsigned int*>(org_g + 2 * n);
    CK(ctx, cudaMemsetAsync(done_ctr, 0, sizeof(unsigned int), ctx->stream));
    if (t <= 8) {
      k_tail_svd8<<<1, 32, 0, ctx->stream>>>(n, ctl_k, B, Pu, sT, flag);
      count_launch(ctx);
    } else {
      k_tail_prep<<<1, 64, 0, ctx->stream>>>(n, ctl_k, B, Tm);
      count_launch(ctx);
      RET(launch_jacobi_svd(ctx, t, Tm, Pu, sT, Pv, jscr, jinfo, nullptr));
    }
    const int grid = (n + AR_WARPS - 1) / AR_WARPS;
    k_arrow_roots<<<grid, AR_THREADS, 0, ctx->stream>>>(2, n, ctl_k, 0, B, sT, Pu, org_g, mu_g, sg_g, uh_g, done_ctr, flag);
    k_arrow_vectors<<<grid, AR_THREADS, 0, ctx->stream>>>(2, n, ctl_k, 0, B, sT, Pu, org_g, mu_g, sg_g, uh_g, S, Yp, Xp,
                                                          flag);
    k_tail_finish<<<n, 128, 0, ctx->stream>>>(n, ctl_k, B, Pu, org_g + n, Yp, S, U, V, flag);
    count_launch(ctx, 3);
    CK(ctx, cudaGetLastError());
    return M3B_OK;
  }
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
  // root buffers (org as ints, mu, sg) live behind the sigma ping-pong in the scratch area
  double* mu_g = sb + n;
  double* sg_g = mu_g + n;
  double* uh_g = sg_g + n;
  int* org_g = reinterpret_cast<int*>(uh_g + n);
  unsigned int* done_ctr = reinterpret_cast<unsigned int*>(org_g + n);  // zeroed with the flag; self re-arming
  auto solve = [&](int mode, int tail_j, int m, const double* sig_in, const double* Ycur_, double* sig_out_, double* Yo,
                   double* Xo) {
    const int grid = (m + AR_WARPS - 1) / AR_WARPS;
    k_arrow_roots<<<grid, AR_THREADS, 0, ctx->stream>>>(mode, n, ctl_k, tail_j, B, sig_in, Ycur_, org_g, mu_g, sg_g, uh_g,
                                                        done_ctr, flag);
    k_arrow_vectors<<<grid, AR_THREADS, 0, ctx->stream>>>(mode, n, ctl_k, tail_j, B, sig_in, Ycur_, org_g, mu_g, sg_g,
                                                          uh_g, sig_out_, Yo, Xo, flag);
    count_launch(ctx, 2);
  };
  solve(0, 0, k_host + 1, nullptr, nullptr, s_cur, Yc, Xc);
  for (int t = 0; t < ntail; ++t) {
    const int j = k_host + 1 + t;  // new row / column
    double* s_next = (t == ntail - 1) ? S : (s_cur == sa ? sb : sa);
    solve(1, j, j + 1, s_cur, Yc, s_next, Yp, Xp);
    dim3 grid((j + 1 + 15) / 16, (j + 1 + 15) / 16, 2);
    k_arrow_fold<<<grid, 256, 0, ctx->stream>>>(n, j, Yc, Xc, Yp, Xp, Yn, Xn, flag);
    count_launch(ctx);
    std::swap(Yc, Yn);
    std::swap(Xc, Xn);
    s_cur = s_next;
  }
  k_svd_signfix<<<(n * 32 + 255) / 256, 256, 0, ctx->stream>>>(n, U, V, flag);
  count_launch(ctx);
  CK(ctx, cudaGetLastError());
  return M3B_OK;
}
