// K8: SVD of the small (work x work) projected matrix B, on the device.
//
// Replaces LAPACK dgesdd("O") inside irlba's restart (irlb.c; reached from
// R/utils.R:417).  One thread block, one-sided (Hestenes) Jacobi with a
// round-robin pair schedule: each warp orthogonalises one column pair per round.
// B after a thick restart is diagonal + one spike column + a short bidiagonal
// tail, so a handful of sweeps suffice.  Keeping this on the device removes the
// host round trip (D2H of B, LAPACK, H2D of U/V) from every restart; the host
// only reads back {k, converged}.
//
// Output convention matches dgesdd: U (n x n), s descending, V (n x n, columns are
// right singular vectors; the host-callback path hands over Vt and is transposed
// on upload).  Signs: each pair (u_i, v_i) is normalised so that the largest-
// magnitude component of u_i is positive (LAPACK's signs are not reproducible
// without LAPACK; the optional host callback gives exact sign fidelity).
#include "m3b_internal.h"

#define SVD_THREADS 1024
#define SVD_MAX_SWEEPS 60

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// G (n x n col-major, in/out: B -> U*diag(s) -> U), V (n x n col-major out), s (n) out.
// scratch: order[n] (int) + norms[n].  G and V may live in shared or global memory.
__global__ void __launch_bounds__(SVD_THREADS, 1)
k_jacobi_svd(int n, const double* __restrict__ Bin, double* __restrict__ Uout, double* __restrict__ sout,
             double* __restrict__ Vout, double* __restrict__ gG, double* __restrict__ gV, int use_smem,
             int* __restrict__ info) {
  extern __shared__ double sm[];
  __shared__ int s_rot;
  __shared__ double s_norm[256];
  __shared__ int s_rank[256];
  double* G = use_smem ? sm : gG;
  double* V = use_smem ? (sm + (size_t)n * n) : gV;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
  for (int k = tid; k < n * n; k += blockDim.x) {
    G[k] = Bin[k];
    V[k] = ((k / n) == (k % n)) ? 1.0 : 0.0;
  }
  __syncthreads();
  const int np = (n + 1) & ~1;  // players (even)
  const int rounds = np - 1, pairs = np / 2;
  const double tol = 1e-15;
  int sweep = 0;
  for (; sweep < SVD_MAX_SWEEPS; ++sweep) {
    if (tid == 0) s_rot = 0;
    __syncthreads();
    for (int r = 0; r < rounds; ++r) {
      for (int k = warp; k < pairs; k += nwarps) {
        int a, b;
        if (k == 0) {
          a = r;
          b = np - 1;
        } else {
          a = (r + k) % (np - 1);
          b = (r - k + (np - 1)) % (np - 1);
        }
        int p = min(a, b), q = max(a, b);
        if (q >= n) continue;  // dummy player of an odd n
        double* gp = G + (size_t)p * n;
        double* gq = G + (size_t)q * n;
        double al = 0.0, be = 0.0, ga = 0.0;
        for (int i = lane; i < n; i += 32) {
          double x = gp[i], y = gq[i];
          al += x * x;
          be += y * y;
          ga += x * y;
        }
        al = warp_sum(al);
        be = warp_sum(be);
        ga = warp_sum(ga);
        if (fabs(ga) > tol * sqrt(al * be) && al > 0.0 && be > 0.0) {
          double zeta = (be - al) / (2.0 * ga);
          double t = copysign(1.0, zeta) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
          double c = 1.0 / sqrt(1.0 + t * t), s = c * t;
          for (int i = lane; i < n; i += 32) {
            double x = gp[i], y = gq[i];
            gp[i] = c * x - s * y;
            gq[i] = s * x + c * y;
          }
          double* vp = V + (size_t)p * n;
          double* vq = V + (size_t)q * n;
          for (int i = lane; i < n; i += 32) {
            double x = vp[i], y = vq[i];
            vp[i] = c * x - s * y;
            vq[i] = s * x + c * y;
          }
          if (lane == 0) s_rot = 1;
        }
      }
      __syncthreads();
    }
    int rot = s_rot;
    __syncthreads();
    if (!rot) break;
  }
  // singular values = column norms
  for (int j = warp; j < n; j += nwarps) {
    double a = 0.0;
    for (int i = lane; i < n; i += 32) {
      double x = G[(size_t)j * n + i];
      a += x * x;
    }
    a = warp_sum(a);
    if (lane == 0) s_norm[j] = sqrt(a);
  }
  __syncthreads();
  // rank sort, descending (ties broken by index)
  for (int j = tid; j < n; j += blockDim.x) {
    int rk = 0;
    double me = s_norm[j];
    for (int i = 0; i < n; ++i) {
      double o = s_norm[i];
      rk += (o > me || (o == me && i < j)) ? 1 : 0;
    }
    s_rank[j] = rk;
  }
  __syncthreads();
  for (int j = warp; j < n; j += nwarps) {
    const int dst = s_rank[j];
    const double sg = s_norm[j];
    const double inv = sg > 0.0 ? 1.0 / sg : 0.0;
    // sign convention: largest |u| component positive
    double best = 0.0;
    int bi = n;
    for (int i = lane; i < n; i += 32) {
      double x = fabs(G[(size_t)j * n + i]);
      if (x > best) {
        best = x;
        bi = i;
      }
    }
    for (int o = 16; o > 0; o >>= 1) {
      double ob = __shfl_xor_sync(0xffffffffu, best, o);
      int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ob > best || (ob == best && oi < bi)) {
        best = ob;
        bi = oi;
      }
    }
    double sgn = 1.0;
    if (bi < n && G[(size_t)j * n + bi] < 0.0) sgn = -1.0;
    for (int i = lane; i < n; i += 32) {
      Uout[(size_t)dst * n + i] = G[(size_t)j * n + i] * inv * sgn;
      Vout[(size_t)dst * n + i] = V[(size_t)j * n + i] * sgn;
    }
    if (lane == 0) sout[dst] = sg;
  }
  if (tid == 0) {
    info[0] = sweep;
  }
}

int launch_jacobi_svd(m3b_ctx* ctx, int n, const double* B, double* U, double* s, double* V, double* scratch,
                      int* info) {
  if (n > 256) {
    m3b_set_error(ctx, "small SVD: work=%d > 256 is not supported", n);
    return M3B_E_DIMS;
  }
  size_t need = 2 * (size_t)n * n * sizeof(double);
  int use_smem = need <= 220 * 1024;
  static bool attr = false;
  if (!attr) {
    CK(ctx, cudaFuncSetAttribute(k_jacobi_svd, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    attr = true;
  }
  k_jacobi_svd<<<1, SVD_THREADS, use_smem ? need : 0, ctx->stream>>>(n, B, U, s, V, scratch, scratch + (size_t)n * n,
                                                                      use_smem, info);
  count_launch(ctx);
  CK(ctx, cudaGetLastError());
  return M3B_OK;
}
