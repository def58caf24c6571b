// K8: SVD of the small (work x work) projected matrix B, on the device.
//
// Replaces LAPACK dgesdd("O") inside irlba's restart (irlb.c; reached from
// R/utils.R:417).  One thread block, one-sided (Hestenes) Jacobi.
//
// Two things make it fast enough to sit inside every restart:
//  * it runs on X = B^T, not on B.  B after a thick restart is diag(BS[:k]) plus
//    one spike COLUMN plus a short bidiagonal tail; the columns of B^T (= rows
//    of B) are then almost orthogonal already (their inner products are
//    res_i*res_j against norms BS_i*BS_j), so Jacobi needs only a few sweeps
//    (the Drmac-Veselic observation for triangular factors).
//      X J = Y (orthogonal columns),  X = P S J^T  =>  B = J S P^T
//    so U_B = J (accumulated rotations) and V_B = normalised columns of Y.
//  * a half-warp (16 lanes) per column pair: all n/2 <= 64 disjoint pairs of a
//    round-robin round run in ONE pass of the 1024-thread block; the squared
//    column norms are carried along (alpha -= t*gamma, beta += t*gamma) and
//    refreshed once per sweep, so a pair costs one dot product, not three.
//
// Output convention matches dgesdd: U (n x n), s descending, V (n x n, columns
// are right singular vectors).  Signs: each pair (u_i, v_i) is normalised so that
// the largest-magnitude component of u_i is positive (the optional host callback
// supplies LAPACK when exact sign fidelity with R is wanted).
#include "m3b_internal.h"

#define SVD_THREADS 1024
#define SVD_MAX_SWEEPS 60
#define SVD_NMAX 256

__device__ __forceinline__ double half_sum(double v) {  // sum over a 16-lane half-warp
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o, 16);
  return v;
}
__device__ __forceinline__ double warp_sum32(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Bin (n x n col-major).  Uout/Vout (n x n col-major), sout (n).  gX/gJ: global scratch used
// when 2 n^2 doubles do not fit in shared memory.  NIT = ceil(n/16) when n <= 128 (the column
// loops are then fully unrolled and predicated, so the shared-memory loads of a pair are all
// in flight together), 0 = generic loops.
template <int NIT>
__global__ void __launch_bounds__(SVD_THREADS, 1)
k_jacobi_svd(int n, const double* __restrict__ Bin, double* __restrict__ Uout, double* __restrict__ sout,
             double* __restrict__ Vout, double* __restrict__ gX, double* __restrict__ gJ, int use_smem,
             int* __restrict__ info, const int* __restrict__ only_if) {
  // only_if != nullptr: run only when the structured fast path (arrow_svd.cu) raised its flag
  if (only_if && *only_if == 0) return;
  extern __shared__ double sm[];
  __shared__ int s_rot;
  __shared__ double s_nrm[SVD_NMAX];  // squared column norms of X (carried along)
  __shared__ int s_rank[SVD_NMAX];
  double* X = use_smem ? sm : gX;
  double* J = use_smem ? (sm + (size_t)n * n) : gJ;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
  const int hl = tid & 15, half = tid >> 4, nhalves = blockDim.x >> 4;
  for (int k = tid; k < n * n; k += blockDim.x) {
    const int r = k % n, c = k / n;
    X[k] = Bin[(size_t)r * n + c];  // X = B^T
    J[k] = (r == c) ? 1.0 : 0.0;
  }
  __syncthreads();
  const int np = (n + 1) & ~1;  // players (even)
  const int rounds = np - 1, pairs = np / 2;
  const double tol = 1.0e-15 * sqrt((double)n);
  const double tol2 = tol * tol;
  int sweep = 0;
  for (; sweep < SVD_MAX_SWEEPS; ++sweep) {
    if (tid == 0) s_rot = 0;
    // refresh the squared column norms
    for (int j = warp; j < n; j += nwarps) {
      double a = 0.0;
      for (int i = lane; i < n; i += 32) {
        const double x = X[(size_t)j * n + i];
        a += x * x;
      }
      a = warp_sum32(a);
      if (lane == 0) s_nrm[j] = a;
    }
    __syncthreads();
    // round-robin: player a of pair k in round r (players 0..np-2 rotate, np-1 is fixed)
    int ra = half % (np - 1), rb = (np - 1 - half % (np - 1)) % (np - 1);  // r = 0
    for (int r = 0; r < rounds; ++r) {
      for (int k0 = 0; k0 < pairs; k0 += nhalves) {
        const int k = k0 + half;
        int a = 0, b = 0;
        bool valid = k < pairs;
        if (valid) {
          if (k == 0) {
            a = r;
            b = np - 1;
          } else if (k0 == 0) {
            a = ra;
            b = rb;
          } else {
            a = (r + k) % (np - 1);
            b = (r - k + (np - 1)) % (np - 1);
          }
        }
        const int p = min(a, b), q = max(a, b);
        valid = valid && (q < n);  // dummy player of an odd n
        double* xp = X + (size_t)p * n;
        double* xq = X + (size_t)q * n;
        double ga = 0.0;
        double xv[NIT > 0 ? NIT : 1], yv[NIT > 0 ? NIT : 1];
        if (NIT > 0) {
#pragma unroll
          for (int u = 0; u < NIT; ++u) {
            const int i = hl + 16 * u;
            const bool ok = valid && i < n;
            xv[u] = ok ? xp[i] : 0.0;
            yv[u] = ok ? xq[i] : 0.0;
          }
#pragma unroll
          for (int u = 0; u < NIT; ++u) ga += xv[u] * yv[u];
        } else if (valid) {
          for (int i = hl; i < n; i += 16) ga += xp[i] * xq[i];
        }
        ga = half_sum(ga);
        if (valid) {
          const double al = s_nrm[p], be = s_nrm[q];
          if (ga * ga > tol2 * (al * be) && al > 0.0 && be > 0.0) {
            // t = tan(theta) of the rotation that zeroes <x_p, x_q>: smaller root of
            // t^2 + 2 zeta t - 1 = 0 with zeta = (be - al) / (2 ga), written with one sqrt, one
            // division and one rsqrt:  t = sign(a) g2 / (|a| + hypot(a, g2)),  c = rsqrt(1 + t^2)
            const double a_ = be - al, g2 = 2.0 * ga;
            const double h = sqrt(a_ * a_ + g2 * g2);
            const double t = copysign(g2, a_ * 1.0 >= 0.0 ? g2 : -g2) / (fabs(a_) + h);
            const double c = rsqrt(1.0 + t * t), s = c * t;
            double* jp = J + (size_t)p * n;
            double* jq = J + (size_t)q * n;
            if (NIT > 0) {
#pragma unroll
              for (int u = 0; u < NIT; ++u) {
                const int i = hl + 16 * u;
                if (i < n) {
                  xp[i] = c * xv[u] - s * yv[u];
                  xq[i] = s * xv[u] + c * yv[u];
                }
              }
              // xv/yv are dead now: reuse the registers for the J pair
#pragma unroll
              for (int u = 0; u < NIT; ++u) {
                const int i = hl + 16 * u;
                xv[u] = (i < n) ? jp[i] : 0.0;
                yv[u] = (i < n) ? jq[i] : 0.0;
              }
#pragma unroll
              for (int u = 0; u < NIT; ++u) {
                const int i = hl + 16 * u;
                if (i < n) {
                  jp[i] = c * xv[u] - s * yv[u];
                  jq[i] = s * xv[u] + c * yv[u];
                }
              }
            } else {
              for (int i = hl; i < n; i += 16) {
                const double x = xp[i], y = xq[i];
                xp[i] = c * x - s * y;
                xq[i] = s * x + c * y;
              }
              for (int i = hl; i < n; i += 16) {
                const double x = jp[i], y = jq[i];
                jp[i] = c * x - s * y;
                jq[i] = s * x + c * y;
              }
            }
            if (hl == 0) {
              s_nrm[p] = al - t * ga;
              s_nrm[q] = be + t * ga;
              s_rot = 1;
            }
          }
        }
      }
      // advance the rotating players of this half-warp's first-pass pair
      ra = (ra + 1 == np - 1) ? 0 : ra + 1;
      rb = (rb + 1 == np - 1) ? 0 : rb + 1;
      __syncthreads();
    }
    const int rot = s_rot;
    __syncthreads();
    if (!rot) break;
  }
  // singular values = column norms of Y (recomputed exactly)
  for (int j = warp; j < n; j += nwarps) {
    double a = 0.0;
    for (int i = lane; i < n; i += 32) {
      const double x = X[(size_t)j * n + i];
      a += x * x;
    }
    a = warp_sum32(a);
    if (lane == 0) s_nrm[j] = sqrt(a);
  }
  __syncthreads();
  // rank sort, descending (ties broken by index)
  for (int j = tid; j < n; j += blockDim.x) {
    int rk = 0;
    const double me = s_nrm[j];
    for (int i = 0; i < n; ++i) {
      const double o = s_nrm[i];
      rk += (o > me || (o == me && i < j)) ? 1 : 0;
    }
    s_rank[j] = rk;
  }
  __syncthreads();
  for (int j = warp; j < n; j += nwarps) {
    const int dst = s_rank[j];
    const double sg = s_nrm[j];
    const double inv = sg > 0.0 ? 1.0 / sg : 0.0;
    // sign convention: largest |u| component positive (u = column of J)
    double best = 0.0;
    int bi = n;
    for (int i = lane; i < n; i += 32) {
      const double x = fabs(J[(size_t)j * n + i]);
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
    double sgn = 1.0;
    if (bi < n && J[(size_t)j * n + bi] < 0.0) sgn = -1.0;
    for (int i = lane; i < n; i += 32) {
      Uout[(size_t)dst * n + i] = J[(size_t)j * n + i] * sgn;
      Vout[(size_t)dst * n + i] = X[(size_t)j * n + i] * inv * sgn;
    }
    if (lane == 0) sout[dst] = sg;
  }
  if (tid == 0) info[0] = sweep;
}

int svd_init(m3b_ctx* ctx) {
  const int bytes = 216 * 1024;
  CK(ctx, cudaFuncSetAttribute(k_jacobi_svd<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  CK(ctx, cudaFuncSetAttribute(k_jacobi_svd<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  CK(ctx, cudaFuncSetAttribute(k_jacobi_svd<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  CK(ctx, cudaFuncSetAttribute(k_jacobi_svd<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  CK(ctx, cudaFuncSetAttribute(k_jacobi_svd<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  CK(ctx, cudaFuncSetAttribute(k_jacobi_svd<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  CK(ctx, cudaFuncSetAttribute(k_jacobi_svd<6>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  CK(ctx, cudaFuncSetAttribute(k_jacobi_svd<7>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  CK(ctx, cudaFuncSetAttribute(k_jacobi_svd<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  return M3B_OK;
}

int launch_jacobi_svd(m3b_ctx* ctx, int n, const double* B, double* U, double* s, double* V, double* scratch,
                      int* info, const int* only_if) {
  if (n > SVD_NMAX) {
    m3b_set_error(ctx, "small SVD: work=%d > %d is not supported", n, SVD_NMAX);
    return M3B_E_DIMS;
  }
  size_t need = 2 * (size_t)n * n * sizeof(double);
  int use_smem = need <= 216 * 1024;
  const int nit = (n <= 128) ? (n + 15) / 16 : 0;
  void (*fn)(int, const double*, double*, double*, double*, double*, double*, int, int*, const int*);
  switch (nit) {
    case 1: fn = k_jacobi_svd<1>; break;
    case 2: fn = k_jacobi_svd<2>; break;
    case 3: fn = k_jacobi_svd<3>; break;
    case 4: fn = k_jacobi_svd<4>; break;
    case 5: fn = k_jacobi_svd<5>; break;
    case 6: fn = k_jacobi_svd<6>; break;
    case 7: fn = k_jacobi_svd<7>; break;
    case 8: fn = k_jacobi_svd<8>; break;
    default: fn = k_jacobi_svd<0>; break;
  }
  fn<<<1, SVD_THREADS, use_smem ? need : 0, ctx->stream>>>(n, B, U, s, V, scratch, scratch + (size_t)n * n, use_smem, info, only_if);
  count_launch(ctx);
  CK(ctx, cudaGetLastError());
  return M3B_OK;
}
