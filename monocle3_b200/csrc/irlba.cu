// K3 (gene moments), K6/K7 (orthogonalisation, norms), K9 (restart GEMMs) and the
// IRLBA driver that strings them together with K4/K5 (spmv.cu) and K8 (small_svd.cu).
//
// Replaces irlba::irlba(A=t(FM), nv, center, scale) and the post-processing of
// sparse_prcomp_irlba (R/utils.R:383-428).  The iteration is irlba's fast C path
// step for step: single-vector implicitly restarted
// Lanczos bidiagonalisation, work = nv+7, one classical Gram-Schmidt pass per
// vector, thick restart on the Ritz vectors, irlba's convergence test.
//
// Everything between two restarts is queued on one stream without host
// synchronisation: all scalars (S, R_F, beta, sum(w)) live in device memory and
// every reduction is a fixed-order two-stage reduction (bit-reproducible).  The
// host reads back two ints per restart ({k, converged}).
//
// Operator (m = local cells, n = kept genes):
//   Ax(x):  xs = x / s;  w = A' xs - (ce . xs)
//   Atw(w): f = (A'^T w) / s - sum(w) * ce / s
// where A' is the stored sparse matrix, s the gene sd (or 1), and ce the
// effective centre = mean(A') when scaling, -f0 on the reference's dense
// pseudo-count paths without scaling, 0 otherwise (DESIGN.md "implicit constant").
#include <algorithm>
#include <cmath>
#include <chrono>
#include <cstring>
#include <vector>

#include "m3b_internal.h"

int launch_jacobi_svd(m3b_ctx* ctx, int n, const double* B, double* U, double* s, double* V, double* scratch, int* info,
                      const int* only_if);
int launch_arrow_svd(m3b_ctx* ctx, int n, int k_host, const int* ctl_k, const double* B, double* U, double* S, double* V,
                     double* scratch, int* flag);

#define NT 256
#define M3B_OF_SMEM_MAX (200 * 1024)  // dynamic shared memory the fused vector kernel may ask for
#define MAX_PART 592  // upper bound on per-kernel partial counts (4 * 148)

// scalar slots in the device scalar buffer
enum { SC_S = 0, SC_RF = 1, SC_SUMW = 2, SC_NRM2 = 3, SC_SUM = 4, SC_COUNT = 8 };
// control block (ints)
enum { CT_K = 0, CT_CONV = 1, CT_FLAG = 2, CT_SWEEPS = 3, CT_SVDFLAG = 4, CT_COUNT = 8 };

// --------------------------------------------------------------------------
// deterministic block reduction (fixed tree): result valid in thread 0
// --------------------------------------------------------------------------
__device__ __forceinline__ double block_sum(double v, double* sh /* >= 32 */) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if (lane == 0) sh[warp] = v;
  __syncthreads();
  double r = 0.0;
  if (threadIdx.x == 0) {
    const int nw = blockDim.x >> 5;
    for (int w = 0; w < nw; ++w) r += sh[w];
  }
  __syncthreads();
  return r;
}

// sum of a short partial array, identical in every block that calls it (broadcast via smem)
__device__ __forceinline__ double block_sum_partials(const double* __restrict__ p, int np, double* sh) {
  double v = 0.0;
  for (int k = threadIdx.x; k < np; k += blockDim.x) v += p[k];
  double r = block_sum(v, sh);
  __shared__ double bc;
  if (threadIdx.x == 0) bc = r;
  __syncthreads();
  return bc;
}

// sum of a row's partials: its slots are contiguous; 8 lanes per row, fixed shuffle tree.
// Must be called by all 32 lanes of a warp (rows beyond the end pass b == e).
__device__ __forceinline__ double row_sum8(const double* __restrict__ partial, int b, int e, int sub) {
  double a = 0.0;
  for (int k = b + sub; k < e; k += 8) a += partial[k];
  a += __shfl_xor_sync(0xffffffffu, a, 4);
  a += __shfl_xor_sync(0xffffffffu, a, 2);
  a += __shfl_xor_sync(0xffffffffu, a, 1);
  return a;
}

// --------------------------------------------------------------------------
// gene-side vector kernels
// --------------------------------------------------------------------------
// xs = v / s ; pb[block] = partial dot(xs, ce)
__global__ void __launch_bounds__(NT) k_make_xs(const double* __restrict__ v, const double* __restrict__ s,
                                                const double* __restrict__ ce, int n, double* __restrict__ xs,
                                                double* __restrict__ pb) {
  __shared__ double sh[32];
  double acc = 0.0;
  for (int g = blockIdx.x * blockDim.x + threadIdx.x; g < n; g += gridDim.x * blockDim.x) {
    double x = v[g];
    if (s) x = x / s[g];
    xs[g] = x;
    if (ce) acc += x * ce[g];
  }
  double r = block_sum(acc, sh);
  if (threadIdx.x == 0) pb[blockIdx.x] = r;
}

// w[r] = sum(partials of row r) - beta - R_F * wprev[r]      (K4 epilogue)
__global__ void __launch_bounds__(NT) k_combine_A(const double* __restrict__ partial, const int* __restrict__ rsp,
                                                  long long rows, const double* __restrict__ pb, int npb,
                                                  const double* __restrict__ wprev, const double* __restrict__ scal,
                                                  double* __restrict__ w) {
  __shared__ double sh[32];
  const double beta = pb ? block_sum_partials(pb, npb, sh) : 0.0;
  const double rf = wprev ? scal[SC_RF] : 0.0;
  const int sub = threadIdx.x & 7;
  const long long rpb = blockDim.x >> 3;
  for (long long r0 = blockIdx.x * rpb; r0 < rows; r0 += (long long)gridDim.x * rpb) {
    const long long r = r0 + (threadIdx.x >> 3);
    const int b = (r < rows) ? rsp[r] : 0, e = (r < rows) ? rsp[r + 1] : 0;
    double acc = row_sum8(partial, b, e, sub);
    if (r < rows && sub == 0) {
      acc -= beta;
      if (wprev) acc -= rf * wprev[r];
      w[r] = acc;
    }
  }
}

// Fraw[g] = sum(partials of row g)                           (K5 combine, local shard)
// F[g] = Fraw[g]/s - (sumw*ce)/s - S*vj[g]                   (fix-ups, after the allreduce)
__global__ void __launch_bounds__(NT) k_fix_F(const double* __restrict__ Fraw, const double* __restrict__ s,
                                              const double* __restrict__ ce, const double* __restrict__ vj,
                                              const double* __restrict__ scal, int n, double* __restrict__ F) {
  const double sumw = scal[SC_SUMW], S = scal[SC_S];
  for (int g = blockIdx.x * blockDim.x + threadIdx.x; g < n; g += gridDim.x * blockDim.x) {
    double f = Fraw[g];
    if (s) f = f / s[g];
    if (ce) {
      double c = sumw * ce[g];
      if (s) c = c / s[g];
      f = f - c;
    }
    f = f - S * vj[g];
    F[g] = f;
  }
}

// single-rank fusion of the K5 combine with the fix-ups (no allreduce in between):
// F[g] = (sum of row g's partials)/s - (sumw*ce)/s - S*vj[g]
__global__ void __launch_bounds__(NT) k_combine_fix_F(const double* __restrict__ partial, const int* __restrict__ rsp,
                                                      int rows, const double* __restrict__ s,
                                                      const double* __restrict__ ce, const double* __restrict__ vj,
                                                      const double* __restrict__ scal, double* __restrict__ F) {
  const double sumw = scal[SC_SUMW], S = scal[SC_S];
  const int sub = threadIdx.x & 7;
  const int rpb = blockDim.x >> 3;
  for (int g0 = blockIdx.x * rpb; g0 < rows; g0 += gridDim.x * rpb) {
    const int g = g0 + (threadIdx.x >> 3);
    const int b = (g < rows) ? rsp[g] : 0, e = (g < rows) ? rsp[g + 1] : 0;
    double f = row_sum8(partial, b, e, sub);
    if (g < rows && sub == 0) {
      if (s) f = f / s[g];
      if (ce) {
        double c = sumw * ce[g];
        if (s) c = c / s[g];
        f = f - c;
      }
      F[g] = f - S * vj[g];
    }
  }
}

// --------------------------------------------------------------------------
// K5 combine + cross-rank sum + fix-ups in ONE kernel over peer memory (NVLink / NVSwitch).
//
// Cells are sharded, so every rank holds a partial A^T w (a gene-length vector, 160-240 KB).
// Instead of  combine -> ncclAllReduce -> fix-ups  (three launches, one NCCL round trip of
// ~15-20 us per Lanczos step) each rank
//   1. adds the partials of its gene rows and writes them into ITS slot of the exchange buffer,
//   2. the last CTA to finish publishes a sequence number into every peer's flag array
//      (system-scope release stores through the peers' mapped pointers),
//   3. every CTA waits until all peers have published the same number, then reads the peers'
//      slots DIRECTLY (ld over NVLink), adds the world's partials in rank order -- the same
//      order on every rank, so the replicated gene-space state stays bit-identical -- and
//      applies  /s - sum(w) ce / s - S v_j.
// Two slots alternate (a peer can lag at most one step behind: it cannot publish step k+1
// before it has finished reading step k).  The grid (<= 296 CTAs) is always co-resident, so the
// in-kernel wait cannot starve the publishing CTA.  A bounded spin (about a second) turns a
// missing peer into an error flag instead of a hang.
// --------------------------------------------------------------------------
#define XF_CTAS 296  // two CTAs per SM: always co-resident (256 threads, no shared memory to speak of)

__device__ __forceinline__ unsigned long long ld_acquire_sys_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys_u64(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ double ld_relaxed_sys_f64(const double* p) {
  double v;
  asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
  return v;
}

__global__ void __launch_bounds__(NT)
k_combine_xchg_fix_F(const double* __restrict__ partial, const int* __restrict__ rsp, int rows,
                     const double* __restrict__ s, const double* __restrict__ ce, const double* __restrict__ vj,
                     const double* __restrict__ scal, double* __restrict__ F, double* const* __restrict__ peers,
                     int rank, int world, unsigned long long seq, unsigned int* __restrict__ done_ctr,
                     int* __restrict__ ctl, long long timeout) {
  __shared__ int s_ok;
  if (*((volatile int*)(ctl + CT_FLAG)) & 4) return;  // an earlier exchange already timed out: do not wait again
  const int slot = (int)(seq & 1ull);
  const size_t flags_doubles = (size_t)world * 2 * M3B_XCHG_FLAG_STRIDE;
  double* mine = peers[rank] + flags_doubles + (size_t)slot * M3B_XCHG_CAP;
  // contiguous row range of this CTA
  const int per = (rows + gridDim.x - 1) / gridDim.x;
  const int g_lo = blockIdx.x * per, g_hi = min(rows, g_lo + per);
  // 1. local combine (8 lanes per gene row, fixed shuffle tree)
  {
    const int sub = threadIdx.x & 7, rpb = blockDim.x >> 3;
    for (int g0 = g_lo; g0 < g_hi; g0 += rpb) {
      const int g = g0 + (threadIdx.x >> 3);
      const int b = (g < g_hi) ? rsp[g] : 0, e = (g < g_hi) ? rsp[g + 1] : 0;
      const double f = row_sum8(partial, b, e, sub);
      if (g < g_hi && sub == 0) mine[g] = f;
    }
  }
  // 2. publish: the last CTA of this rank writes seq into every peer's flag for (rank, slot)
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int d = atomicInc(done_ctr, gridDim.x - 1);
    if (d == gridDim.x - 1) {
      __threadfence_system();
      for (int q = 0; q < world; ++q) {
        unsigned long long* fl =
            reinterpret_cast<unsigned long long*>(peers[q] + ((size_t)rank * 2 + slot) * M3B_XCHG_FLAG_STRIDE);
        st_release_sys_u64(fl, seq);
      }
    }
    // 3a. wait for every rank's flag in MY buffer (bounded spin)
    int ok = 1;
    const long long t0 = clock64();
    for (int q = 0; q < world && ok; ++q) {
      const unsigned long long* fl =
          reinterpret_cast<const unsigned long long*>(peers[rank] + ((size_t)q * 2 + slot) * M3B_XCHG_FLAG_STRIDE);
      while (ld_acquire_sys_u64(fl) != seq) {
        if (clock64() - t0 > timeout) {  // a peer never arrived
          ok = 0;
          break;
        }
        __nanosleep(64);
      }
    }
    s_ok = ok;
    if (!ok) atomicOr(ctl + CT_FLAG, 4);
  }
  __syncthreads();
  if (!s_ok) return;
  // 3b. sum the world's partials in rank order straight out of the peers' memory, then fix up
  const double sumw = scal[SC_SUMW], S = scal[SC_S];
  for (int g = g_lo + threadIdx.x; g < g_hi; g += blockDim.x) {
    double f = 0.0;
    for (int q = 0; q < world; ++q)
      f += ld_relaxed_sys_f64(peers[q] + flags_doubles + (size_t)slot * M3B_XCHG_CAP + g);
    if (s) f = f / s[g];
    if (ce) {
      double c = sumw * ce[g];
      if (s) c = c / s[g];
      f = f - c;
    }
    F[g] = f - S * vj[g];
  }
}

// --------------------------------------------------------------------------
// K6: tall-skinny orthogonalisation  y -= X (X^T y), one classical GS pass
// --------------------------------------------------------------------------
// part[blk*ncols + c] = sum over the block's rows of X[r,c]*y[r]
// A block owns `chunk` rows (y chunk staged in shared memory); a warp takes 4 columns at a
// time so that every lane keeps 4 independent load streams in flight.
__global__ void __launch_bounds__(NT) k_dots(const double* __restrict__ X, long long ld, long long rows, int ncols,
                                             const double* __restrict__ y, int chunk, double* __restrict__ part) {
  extern __shared__ double ys[];  // chunk doubles
  const long long r0 = (long long)blockIdx.x * chunk;
  const int len = (int)min((long long)chunk, rows - r0);
  for (int k = threadIdx.x; k < len; k += blockDim.x) ys[k] = y[r0 + k];
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  for (int c0 = warp * 4; c0 < ncols; c0 += nwarps * 4) {
    const int nc = min(4, ncols - c0);
    const double* x0 = X + (long long)c0 * ld + r0;
    const double* x1 = x0 + (nc > 1 ? ld : 0);
    const double* x2 = x0 + (nc > 2 ? 2 * ld : 0);
    const double* x3 = x0 + (nc > 3 ? 3 * ld : 0);
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
    for (int k = lane; k < len; k += 32) {
      const double yv = ys[k];
      a0 += x0[k] * yv;
      a1 += x1[k] * yv;
      a2 += x2[k] * yv;
      a3 += x3[k] * yv;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      a0 += __shfl_xor_sync(0xffffffffu, a0, o);
      a1 += __shfl_xor_sync(0xffffffffu, a1, o);
      a2 += __shfl_xor_sync(0xffffffffu, a2, o);
      a3 += __shfl_xor_sync(0xffffffffu, a3, o);
    }
    if (lane == 0) {
      double* pp = part + (long long)blockIdx.x * ncols + c0;
      pp[0] = a0;
      if (nc > 1) pp[1] = a1;
      if (nc > 2) pp[2] = a2;
      if (nc > 3) pp[3] = a3;
    }
  }
}

// t[c] = sum_b part[b*ncols + c]: one warp per column, lanes stride over the blocks, fixed tree
__global__ void __launch_bounds__(NT) k_reduce_dots(const double* __restrict__ part, int nblk, int ncols,
                                                    double* __restrict__ t) {
  const int lane = threadIdx.x & 31;
  const int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (c >= ncols) return;
  double a = 0.0;
  for (int b = lane; b < nblk; b += 32) a += part[(long long)b * ncols + c];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
  if (lane == 0) t[c] = a;
}

// y[r] -= sum_c X[r,c]*t[c];  pn[blk] = partial ||y||^2, ps[blk] = partial sum(y)
// Block = 64 rows x 4 column groups: each thread sums its quarter of the columns with four
// independent accumulators (16 loads in flight per row), the groups are added in a fixed
// order through shared memory.
__global__ void __launch_bounds__(NT) k_apply_orth(const double* __restrict__ X, long long ld, long long rows, int ncols,
                                                   const double* __restrict__ t, double* __restrict__ y,
                                                   double* __restrict__ pn, double* __restrict__ ps) {
  extern __shared__ double ts[];  // ncols
  __shared__ double red[4][64];
  __shared__ double sh[32];
  for (int k = threadIdx.x; k < ncols; k += blockDim.x) ts[k] = t[k];
  __syncthreads();
  const int rl = threadIdx.x & 63, cg = threadIdx.x >> 6;
  double n2 = 0.0, sm = 0.0;
  for (long long rb = blockIdx.x; rb * 64 < rows; rb += gridDim.x) {
    const long long r = rb * 64 + rl;
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
    if (r < rows) {
      const double* xr = X + r;
      int c = cg;
      for (; c + 12 < ncols; c += 16) {
        a0 += xr[(long long)c * ld] * ts[c];
        a1 += xr[(long long)(c + 4) * ld] * ts[c + 4];
        a2 += xr[(long long)(c + 8) * ld] * ts[c + 8];
        a3 += xr[(long long)(c + 12) * ld] * ts[c + 12];
      }
      for (; c < ncols; c += 4) a0 += xr[(long long)c * ld] * ts[c];
    }
    red[cg][rl] = (a0 + a1) + (a2 + a3);
    __syncthreads();
    if (cg == 0 && r < rows) {
      const double v = y[r] - (((red[0][rl] + red[1][rl]) + red[2][rl]) + red[3][rl]);
      y[r] = v;
      n2 += v * v;
      sm += v;
    }
    __syncthreads();
  }
  double r1 = block_sum(n2, sh);
  double r2 = block_sum(sm, sh);
  if (threadIdx.x == 0) {
    pn[blockIdx.x] = r1;
    ps[blockIdx.x] = r2;
  }
}


// --------------------------------------------------------------------------
// Single-rank fusion of one half Lanczos step's vector work into ONE cooperative kernel:
//   combine (row sums of the product's partials + operator fix-ups)  ->  y chunk in shared memory
//   partial dots X^T y            | grid barrier
//   t = sum of the partial dots, y -= X t, partial ||y||^2 and sum(y)   | grid barrier
//   norm, scaling and the scalars of the step (k_scale_w_fused / k_finish_F / k_finish_last)
// The separate kernels (k_combine_*, k_dots, k_reduce_dots, k_apply_orth, k_finish_*) each cost a
// launch gap plus a ramp and a tail on vectors of 20-40 k elements: ~50 us per half step for
// ~10 us of memory time.  Every CTA owns one contiguous chunk of rows for the whole kernel; all
// reductions keep a fixed order (bit-reproducible).  Cell-sharded runs use the same kernel with the
// cross-rank sums inside it (flag-in-data pushes over peer memory, see ll_store / ll_wait below);
// the separate kernels with NCCL calls between them remain as the fallback (no peer access, more
// than M3B_LL_CAPF kept genes, M3B_NO_MR_FUSED=1) and as the careful path of the invariant-subspace
// recovery.
// --------------------------------------------------------------------------
#define OF_THREADS 1024
enum { OF_W = 0, OF_F = 1, OF_FLAST = 2 };

struct OrthArgs {
  int mode;
  long long rows;
  int chunk;
  // combine
  const double* partial;
  const int* rsp;
  const double* pb_in;  // W side: partials of beta (nullptr: beta = 0)
  int npb_in;
  const double* wprev;  // W side: nullptr or the previous W column
  const double* s;      // gene sd or nullptr
  const double* ce;     // effective centre or nullptr
  const double* vj;     // F side: V[:, j]
  // orthogonalisation
  const double* X;
  long long ld;
  int ncols;
  double* y;
  double* part;  // [grid][ncols]
  double* pn;    // [grid]
  double* ps;    // [grid]
  // finish
  double* vnext;
  double* xs;
  double* pb_out;  // [grid]
  double* Bm;
  int work, j;
  double eps;
  double* scal;
  int* ctl;
  unsigned long long* bar;
  unsigned long long bar_base;
  // cell-sharded runs (world > 1): peer-memory exchange buffers (Comm::xchg of every rank), see ll_area
  double* const* peers;
  int rank, world;
  unsigned long long seq;  // number of the first exchange of this launch (F side uses 1, W side 2)
  long long timeout;       // bound of the in-kernel waits, SM cycles
  int xpar;                // parity of the flag-in-data areas this launch uses (alternates per launch and side)
  unsigned long long* diag;  // [8] cycle counters of CTA 0 (see XD_*), or nullptr
};
// diag slots: cycles CTA 0 spent waiting for the peers in the three exchanges, whole-kernel cycles and launch counts
enum { XD_WAIT_F = 0, XD_WAIT_W1 = 1, XD_WAIT_W2 = 2, XD_KERNEL_F = 3, XD_KERNEL_W = 4, XD_N_F = 5, XD_N_W = 6, XD_COUNT = 8 };

// ---- cross-rank exchange inside the fused kernel (one box, CUDA-IPC peer memory) ----
// (Round 1 pulled: every rank wrote its contribution into its own slot, published a flag to every peer
// behind a system-scope fence, and whoever needed the sum read the peers' slots over NVLink.  That
// protocol survives in k_combine_xchg_fix_F, the non-fused fallback; the fused kernel pushes.)
// ---- flag-in-data exchange ("LL": every 16-byte entry carries its own validity flag) ----
// The pull protocol costs, per exchange, a local grid barrier (the slot must be complete before
// it is published), a system-scope fence, a flag store, and then a dependent round trip over NVLink
// for the data.  Here the data is PUSHED: a value travels as {lo32, flag, hi32, flag} in one 16-byte
// store into the receiver's own memory (8-byte halves are single-copy atomic, so a half whose flag
// matches is valid; the flag is the exchange's sequence number, which only grows), the receiver polls
// its local copy.  No fence, no barrier, one one-way trip; every CTA exchanges its own rows.
// Parity: an area is reused by the launch after next of the same side, which a rank can only reach
// after every peer has pushed the launch in between, i.e. has finished reading this one.
__device__ __forceinline__ double* ll_area(double* const* peers, int q, int world, int par) {
  return peers[q] + (size_t)world * 2 * M3B_XCHG_FLAG_STRIDE + 2 * (size_t)M3B_XCHG_CAP +
         (size_t)par * 2 * M3B_LL_ENTRIES(world);
}
__device__ __forceinline__ double* ll_entry_f(double* area, int src, long long row) {
  return area + 2 * ((size_t)src * M3B_LL_CAPF + (size_t)row);
}
__device__ __forceinline__ double* ll_entry_w1(double* area, int world, int src, int c) {
  return area + 2 * ((size_t)world * M3B_LL_CAPF + (size_t)src * M3B_LL_CAPW + c);
}
__device__ __forceinline__ double* ll_entry_w2(double* area, int world, int src, int blk, int k) {
  return area + 2 * ((size_t)world * (M3B_LL_CAPF + M3B_LL_CAPW) + ((size_t)src * M3B_LL_CAPG + blk) * 2 + k);
}
__device__ __forceinline__ void ll_store(double* e, double v, unsigned flag) {
  const unsigned lo = (unsigned)__double2loint(v), hi = (unsigned)__double2hiint(v);
  asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(e), "r"(lo), "r"(flag), "r"(hi), "r"(flag) : "memory");
}
// spins (bounded) until the entry carries `flag`; on time-out raises CT_FLAG bit 4 and returns 0
__device__ __forceinline__ double ll_wait(const double* e, unsigned flag, int* ctl, long long timeout) {
  unsigned a, b, c, d;
  long long t0 = 0;
  for (unsigned spin = 0;; ++spin) {
    asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "l"(e) : "memory");
    if (b == flag && d == flag) break;
    if ((spin & 1023u) == 1023u) {
      if (t0 == 0) t0 = clock64();
      else if (clock64() - t0 > timeout) {
        atomicOr(ctl + CT_FLAG, 4);
        return 0.0;
      }
    }
  }
  return __hiloint2double((int)c, (int)a);
}

// All CTAs are co-resident (cooperative launch).  The spin is bounded (~2 s) so that a lost CTA
// turns into an error (CT_FLAG bit 8) instead of a hung device.
__device__ __forceinline__ void grid_barrier(unsigned long long* bar, unsigned long long target, int* ctl,
                                             long long timeout) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(bar, 1ULL);
    unsigned long long v;
    const long long t0 = clock64();
    for (;;) {
      asm volatile("ld.acquire.gpu.u64 %0, [%1];" : "=l"(v) : "l"(bar) : "memory");
      if (v >= target) break;
      if (clock64() - t0 > timeout) {
        atomicOr(ctl + CT_FLAG, 8);
        break;
      }
    }
    __threadfence();
  }
  __syncthreads();
}

// deterministic sum of a short array written by other CTAs of this launch (L1-bypassing loads)
__device__ __forceinline__ double block_sum_cg(const double* p, int np, double* sh, double* bc) {
  double v = 0.0;
  for (int k = threadIdx.x; k < np; k += blockDim.x) v += __ldcg(p + k);
  double r = block_sum(v, sh);
  if (threadIdx.x == 0) *bc = r;
  __syncthreads();
  const double out = *bc;
  __syncthreads();
  return out;
}

__global__ void __launch_bounds__(OF_THREADS, 1) k_orth_fused(const OrthArgs a) {
  extern __shared__ double of_smem[];
  double* ys = of_smem;                      // [chunk]   this CTA's rows of y
  double* ts = ys + a.chunk;                 // [ncols]
  double* red = ts + ((a.ncols + 1) & ~1);   // [1024]    scratch of the two reductions
  __shared__ double sh[32];
  __shared__ double bc;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const long long r0 = (long long)blockIdx.x * a.chunk;
  const int len = (int)max(0LL, min((long long)a.chunk, a.rows - r0));
  const int ncols = a.ncols;
  const long long t_kernel0 = clock64();
  // ---- phase 0: combine (4 lanes per row)
  {
    double beta = 0.0, rf = 0.0, sumw = 0.0, S = 0.0;
    if (a.mode == OF_W) {
      if (a.pb_in) beta = block_sum_cg(a.pb_in, a.npb_in, sh, &bc);
      if (a.wprev) rf = a.scal[SC_RF];
    } else {
      sumw = a.scal[SC_SUMW];
      S = a.scal[SC_S];
    }
    const int sub = tid & 3;
    for (int q0 = 0; q0 < len; q0 += OF_THREADS / 4) {
      const int q = q0 + (tid >> 2);
      const long long r = r0 + q;
      int b = 0, e = 0;
      if (q < len) {
        b = a.rsp[r];
        e = a.rsp[r + 1];
      }
      double f = 0.0;
      for (int k = b + sub; k < e; k += 4) f += a.partial[k];
      f += __shfl_xor_sync(0xffffffffu, f, 2);
      f += __shfl_xor_sync(0xffffffffu, f, 1);
      if (q < len && sub == 0 && a.world > 1 && a.mode != OF_W) {
        ys[q] = f;  // raw local sum; pushed to the peers and summed over the ranks below
      } else if (q < len && sub == 0) {
        if (a.mode == OF_W) {
          f -= beta;
          if (a.wprev) f -= rf * a.wprev[r];
        } else {
          if (a.s) f = f / a.s[r];
          if (a.ce) {
            double c = sumw * a.ce[r];
            if (a.s) c = c / a.s[r];
            f = f - c;
          }
          f = f - S * a.vj[r];
        }
        ys[q] = f;
      }
    }
    __syncthreads();
    if (a.world > 1 && a.mode != OF_W) {
      // this CTA pushes the raw sums of ITS rows into every peer's area (consecutive threads ->
      // consecutive 16-byte entries of one peer), then adds its rows over the ranks in rank order
      // (=> identical on every rank) out of its own area and applies the fix-ups.  No grid barrier:
      // the same CTA of every rank owns the same rows.
      const unsigned flag = (unsigned)a.seq;
      const int npush = len * (a.world - 1);
      for (int k = tid; k < npush; k += OF_THREADS) {
        const int pr = k / len, q = k - pr * len;
        const int peer = pr + (pr >= a.rank ? 1 : 0);
        ll_store(ll_entry_f(ll_area(a.peers, peer, a.world, a.xpar), a.rank, r0 + q), ys[q], flag);
      }
      __syncthreads();  // ys is rewritten below
      const long long tw0 = clock64();
      double* mine = ll_area(a.peers, a.rank, a.world, a.xpar);
      for (int q = tid; q < len; q += OF_THREADS) {
        const long long r = r0 + q;
        double f = 0.0;
        for (int rk = 0; rk < a.world; ++rk)
          f += (rk == a.rank) ? ys[q] : ll_wait(ll_entry_f(mine, rk, r), flag, a.ctl, a.timeout);
        if (a.s) f = f / a.s[r];
        if (a.ce) {
          double c = sumw * a.ce[r];
          if (a.s) c = c / a.s[r];
          f = f - c;
        }
        ys[q] = f - S * a.vj[r];
      }
      __syncthreads();
      if (a.diag && blockIdx.x == 0 && tid == 0) atomicAdd(a.diag + XD_WAIT_F, (unsigned long long)(clock64() - tw0));
    }
  }
  const unsigned long long nb0 = 0ull;  // local grid barriers used so far
  // ---- phase 1: partial dots of this chunk, a warp takes 4 columns at a time
  for (int c0 = warp * 4; c0 < ncols; c0 += (OF_THREADS / 32) * 4) {
    const int nc = min(4, ncols - c0);
    const double* x0 = a.X + (long long)c0 * a.ld + r0;
    const double* x1 = x0 + (nc > 1 ? a.ld : 0);
    const double* x2 = x0 + (nc > 2 ? 2 * a.ld : 0);
    const double* x3 = x0 + (nc > 3 ? 3 * a.ld : 0);
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
    // four row steps at a time with all sixteen loads issued before the first use (the chunk is
    // a few hundred rows: the loop is pure load latency)
    for (int k0 = lane; k0 < len; k0 += 128) {
      double v0[4], v1[4], v2[4], v3[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        // unconditional loads from a clamped row (a per-lane "in range ? load : 0" makes the
        // compiler select right behind every load, which waits for it); y is 0 past the end
        const int k = min(k0 + 32 * u, len - 1);
        v0[u] = x0[k];
        v1[u] = x1[k];
        v2[u] = x2[k];
        v3[u] = x3[k];
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int k = k0 + 32 * u;
        const double yv = (k < len) ? ys[k] : 0.0;
        a0 += v0[u] * yv;
        a1 += v1[u] * yv;
        a2 += v2[u] * yv;
        a3 += v3[u] * yv;
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      a0 += __shfl_xor_sync(0xffffffffu, a0, o);
      a1 += __shfl_xor_sync(0xffffffffu, a1, o);
      a2 += __shfl_xor_sync(0xffffffffu, a2, o);
      a3 += __shfl_xor_sync(0xffffffffu, a3, o);
    }
    if (lane == 0) {
      double* pp = a.part + (long long)blockIdx.x * ncols + c0;
      pp[0] = a0;
      if (nc > 1) pp[1] = a1;
      if (nc > 2) pp[2] = a2;
      if (nc > 3) pp[3] = a3;
    }
  }
  grid_barrier(a.bar, a.bar_base + (nb0 + 1) * gridDim.x, a.ctl, a.timeout);
  // ---- phase 2a: t = sum over the CTAs' partial dots (group q of threads sums blocks q, q+G, ...)
  {
    const int G = OF_THREADS / max(ncols, 1);
    const int c = tid % max(ncols, 1), q = tid / max(ncols, 1);
    if (q < G && ncols > 0) {
      double v = 0.0;
      for (int b = q; b < (int)gridDim.x; b += G) v += __ldcg(a.part + (long long)b * ncols + c);
      red[q * ncols + c] = v;
    }
    __syncthreads();
    if (tid < ncols) {
      double v = 0.0;
      for (int q2 = 0; q2 < G; ++q2) v += red[q2 * ncols + tid];
      ts[tid] = v;
    }
    __syncthreads();
    if (a.world > 1 && a.mode == OF_W) {
      // the rows of W are sharded: t is the sum of the ranks' local coefficients (exchange a.seq).
      // CTA 0 pushes the local sums to every peer; every CTA adds them in rank order out of its
      // own rank's area (its own rank's term it has just computed itself).
      const unsigned flag = (unsigned)a.seq;
      if (blockIdx.x == 0) {
        const int npush = ncols * (a.world - 1);
        for (int k = tid; k < npush; k += OF_THREADS) {
          const int pr = k / ncols, c2 = k - pr * ncols;
          const int peer = pr + (pr >= a.rank ? 1 : 0);
          ll_store(ll_entry_w1(ll_area(a.peers, peer, a.world, a.xpar), a.world, a.rank, c2), ts[c2], flag);
        }
      }
      const long long tw0 = clock64();
      double* mine = ll_area(a.peers, a.rank, a.world, a.xpar);
      double v = 0.0;
      if (tid < ncols)
        for (int rk = 0; rk < a.world; ++rk)
          v += (rk == a.rank) ? ts[tid] : ll_wait(ll_entry_w1(mine, a.world, rk, tid), flag, a.ctl, a.timeout);
      __syncthreads();
      if (tid < ncols) ts[tid] = v;
      __syncthreads();
      if (a.diag && blockIdx.x == 0 && tid == 0) atomicAdd(a.diag + XD_WAIT_W1, (unsigned long long)(clock64() - tw0));
    }
  }
  // ---- phase 2b: y -= X t on this chunk (128 rows x 8 column groups per pass)
  double n2 = 0.0, sm = 0.0;
  {
    const int rl = tid & 127, cg = tid >> 7;
    for (int q0 = 0; q0 < len; q0 += 128) {
      const int q = q0 + rl;
      double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
      if (q < len) {
        const double* xr = a.X + r0 + q;
        // columns cg, cg+8, ...: sixteen loads (128 columns) issued together per round
        for (int cb = cg; cb < ncols; cb += 128) {
          double xv[16];
#pragma unroll
          for (int u = 0; u < 16; ++u) {
            const int c = min(cb + 8 * u, ncols - 1);  // clamped, unconditional (the factor is 0 past the end)
            xv[u] = xr[(long long)c * a.ld];
          }
#pragma unroll
          for (int u = 0; u < 16; u += 4) {
            const int c = cb + 8 * u;
            a0 += xv[u] * ((c < ncols) ? ts[c] : 0.0);
            a1 += xv[u + 1] * ((c + 8 < ncols) ? ts[c + 8] : 0.0);
            a2 += xv[u + 2] * ((c + 16 < ncols) ? ts[c + 16] : 0.0);
            a3 += xv[u + 3] * ((c + 24 < ncols) ? ts[c + 24] : 0.0);
          }
        }
      }
      red[cg * 128 + rl] = (a0 + a1) + (a2 + a3);
      __syncthreads();
      if (cg == 0 && q < len) {
        double d = 0.0;
#pragma unroll
        for (int g = 0; g < 8; ++g) d += red[g * 128 + rl];
        const double v = ys[q] - d;
        ys[q] = v;
        n2 += v * v;
        sm += v;
      }
      __syncthreads();
    }
  }
  const bool ll_norms = a.world > 1 && a.mode == OF_W;
  double nn, ssum = 0.0;
  {
    const double r1 = block_sum(n2, sh);
    const double r2 = block_sum(sm, sh);
    if (!ll_norms) {
      if (tid == 0) {
        a.pn[blockIdx.x] = r1;
        a.ps[blockIdx.x] = r2;
      }
      grid_barrier(a.bar, a.bar_base + (nb0 + 2) * gridDim.x, a.ctl, a.timeout);
      nn = block_sum_cg(a.pn, gridDim.x, sh, &bc);
      if (a.mode == OF_W) ssum = block_sum_cg(a.ps, gridDim.x, sh, &bc);
    } else {
      // exchange a.seq + 1: every CTA pushes its {|w|^2, sum w} partial to EVERY rank (its own
      // included); having received all world x grid partials IS the barrier, so the local grid
      // barrier and the separate cross-rank exchange of the pull protocol collapse into one one-way
      // trip.  Every CTA of every rank adds the same values in the same order (thread t takes the
      // entries t, t + 1024, ... of the rank-major list, then the fixed block tree).
      const unsigned flag = (unsigned)(a.seq + 1);
      __shared__ double pr2[2];
      if (tid == 0) {
        pr2[0] = r1;
        pr2[1] = r2;
      }
      __syncthreads();
      if (tid < 2 * a.world) {
        const int peer = tid >> 1, k = tid & 1;
        ll_store(ll_entry_w2(ll_area(a.peers, peer, a.world, a.xpar), a.world, a.rank, blockIdx.x, k), pr2[k], flag);
      }
      const long long tw0 = clock64();
      double* mine = ll_area(a.peers, a.rank, a.world, a.xpar);
      const int ntot = a.world * (int)gridDim.x;
      double v1 = 0.0, v2 = 0.0;
      for (int e = tid; e < ntot; e += OF_THREADS) {
        const int rk = e / (int)gridDim.x, b = e - rk * (int)gridDim.x;
        v1 += ll_wait(ll_entry_w2(mine, a.world, rk, b, 0), flag, a.ctl, a.timeout);
        v2 += ll_wait(ll_entry_w2(mine, a.world, rk, b, 1), flag, a.ctl, a.timeout);
      }
      const double g1 = block_sum(v1, sh);
      const double g2 = block_sum(v2, sh);
      if (tid == 0) {
        pr2[0] = g1;
        pr2[1] = g2;
      }
      __syncthreads();
      nn = pr2[0];
      ssum = pr2[1];
      if (a.diag && blockIdx.x == 0 && tid == 0) atomicAdd(a.diag + XD_WAIT_W2, (unsigned long long)(clock64() - tw0));
    }
  }
  // ---- phase 3: norm, scaling, scalars
  if (a.mode == OF_W) {
    const double S = sqrt(nn), SS = 1.0 / S;
    for (int q = tid; q < len; q += OF_THREADS) a.y[r0 + q] = ys[q] * SS;
    if (blockIdx.x == 0 && tid == 0) {
      a.scal[SC_S] = S;
      a.scal[SC_SUMW] = ssum * SS;
      if (!(S >= a.eps)) a.ctl[CT_FLAG] |= 2;
    }
  } else if (a.mode == OF_F) {
    const double rf = sqrt(nn), R = 1.0 / rf;
    double acc = 0.0;
    for (int q = tid; q < len; q += OF_THREADS) {
      const long long g = r0 + q;
      const double v = ys[q] * R;
      a.y[g] = ys[q];
      a.vnext[g] = v;
      const double x = a.s ? v / a.s[g] : v;
      a.xs[g] = x;
      if (a.ce) acc += x * a.ce[g];
    }
    const double r = block_sum(acc, sh);
    if (tid == 0) {
      a.pb_out[blockIdx.x] = r;
      if (blockIdx.x == 0) {
        a.scal[SC_RF] = rf;
        a.Bm[(long long)a.j * a.work + a.j] = a.scal[SC_S];
        a.Bm[(long long)(a.j + 1) * a.work + a.j] = rf;
        if (!(rf >= a.eps)) a.ctl[CT_FLAG] |= 2;
      }
    }
  } else {
    const double rf = sqrt(nn), R = 1.0 / rf;
    for (int q = tid; q < len; q += OF_THREADS) a.y[r0 + q] = ys[q] * R;
    if (blockIdx.x == 0 && tid == 0) {
      a.scal[SC_RF] = (rf < a.eps) ? 0.0 : rf;
      a.Bm[(long long)a.j * a.work + a.j] = a.scal[SC_S];
    }
  }
  if (a.diag && blockIdx.x == 0 && tid == 0) {
    atomicAdd(a.diag + (a.mode == OF_W ? XD_KERNEL_W : XD_KERNEL_F), (unsigned long long)(clock64() - t_kernel0));
    atomicAdd(a.diag + (a.mode == OF_W ? XD_N_W : XD_N_F), 1ULL);
  }
}

// plain norm/sum partials of a vector (first Lanczos vector: no orthogonalisation)
__global__ void __launch_bounds__(NT) k_norm_partials(const double* __restrict__ y, long long rows,
                                                      double* __restrict__ pn, double* __restrict__ ps) {
  __shared__ double sh[32];
  double n2 = 0.0, sm = 0.0;
  for (long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x; r < rows; r += (long long)gridDim.x * blockDim.x) {
    double v = y[r];
    n2 += v * v;
    sm += v;
  }
  double r1 = block_sum(n2, sh);
  double r2 = block_sum(sm, sh);
  if (threadIdx.x == 0) {
    pn[blockIdx.x] = r1;
    ps[blockIdx.x] = r2;
  }
}

// buf2 = {sum pn, sum ps}   (single block; the allreduce of the W side rides on buf2)
__global__ void __launch_bounds__(NT) k_reduce2(const double* __restrict__ pn, const double* __restrict__ ps, int np,
                                                double* __restrict__ buf2) {
  __shared__ double sh[32];
  double a = 0.0, b = 0.0;
  for (int k = threadIdx.x; k < np; k += blockDim.x) {
    a += pn[k];
    b += ps[k];
  }
  double r1 = block_sum(a, sh);
  double r2 = block_sum(b, sh);
  if (threadIdx.x == 0) {
    buf2[0] = r1;
    buf2[1] = r2;
  }
}

// W side: S = sqrt(buf2[0]); w *= 1/S; scal[S], scal[SUMW] = buf2[1]/S
__global__ void __launch_bounds__(NT) k_scale_w(double* __restrict__ w, long long rows, const double* __restrict__ buf2,
                                                double* __restrict__ scal, int* __restrict__ ctl, double eps,
                                                int first, int force_zero) {
  const double S = sqrt(buf2[0]);
  const double SS = 1.0 / S;
  for (long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x; r < rows; r += (long long)gridDim.x * blockDim.x)
    w[r] *= SS;
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    scal[SC_S] = force_zero ? 0.0 : S;  // force_zero: a random replacement vector (see k_scale_w_fused)
    scal[SC_SUMW] = buf2[1] * SS;
    if (!(S >= eps) && !force_zero) ctl[CT_FLAG] |= first ? 1 : 2;  // 1: start vector in null space, 2: breakdown
  }
}

// single-rank fusion of k_reduce2 + k_scale_w: every block reduces the partials itself
// force_zero (invariant-subspace recovery, irlb.c): w is a random replacement vector -- it is
// normalised, but the recurrence coefficient it stands for is 0
__global__ void __launch_bounds__(NT) k_scale_w_fused(double* __restrict__ w, long long rows, const double* __restrict__ pn,
                                                      const double* __restrict__ ps, int np, double* __restrict__ scal,
                                                      int* __restrict__ ctl, double eps, int first, int force_zero) {
  __shared__ double sh[32];
  const double n2 = block_sum_partials(pn, np, sh);
  const double sm = block_sum_partials(ps, np, sh);
  const double S = sqrt(n2);
  const double SS = 1.0 / S;
  for (long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x; r < rows; r += (long long)gridDim.x * blockDim.x)
    w[r] *= SS;
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    scal[SC_S] = force_zero ? 0.0 : S;
    scal[SC_SUMW] = sm * SS;
    if (!(S >= eps) && !force_zero) ctl[CT_FLAG] |= first ? 1 : 2;
  }
}

// V side: R_F = ||F|| from pn; vnext = F * (1/R_F); xs = vnext / s; pb = partial dot(xs, ce);
// B[j,j] = S, B[j,j+1] = R_F
__global__ void __launch_bounds__(NT) k_finish_F(const double* __restrict__ F, const double* __restrict__ pn, int np,
                                                 const double* __restrict__ s, const double* __restrict__ ce, int n,
                                                 double* __restrict__ vnext, double* __restrict__ xs,
                                                 double* __restrict__ pb, double* __restrict__ Bm, int work, int j,
                                                 double* __restrict__ scal, int* __restrict__ ctl, double eps,
                                                 int force_zero) {
  __shared__ double sh[32];
  double rf = sqrt(block_sum_partials(pn, np, sh));
  const double R = 1.0 / rf;
  if (force_zero) rf = 0.0;  // F is a random replacement vector: normalised, coefficient 0
  double acc = 0.0;
  for (int g = blockIdx.x * blockDim.x + threadIdx.x; g < n; g += gridDim.x * blockDim.x) {
    double v = F[g] * R;
    vnext[g] = v;
    double x = s ? v / s[g] : v;
    xs[g] = x;
    if (ce) acc += x * ce[g];
  }
  double r = block_sum(acc, sh);
  if (threadIdx.x == 0) {
    pb[blockIdx.x] = r;
    if (blockIdx.x == 0) {
      scal[SC_RF] = rf;
      Bm[(long long)j * work + j] = scal[SC_S];
      Bm[(long long)(j + 1) * work + j] = rf;
      if (!(rf >= eps) && !force_zero) ctl[CT_FLAG] |= 2;
    }
  }
}

// end of a Lanczos cycle: B[j,j] = S; R_F = ||F||; F *= 1/R_F  (F becomes the restart vector)
__global__ void __launch_bounds__(NT) k_finish_last(double* __restrict__ F, const double* __restrict__ pn, int np, int n,
                                                    double* __restrict__ Bm, int work, int j, double* __restrict__ scal,
                                                    double eps) {
  __shared__ double sh[32];
  const double rf = sqrt(block_sum_partials(pn, np, sh));
  const double R = 1.0 / rf;
  for (int g = blockIdx.x * blockDim.x + threadIdx.x; g < n; g += gridDim.x * blockDim.x) F[g] *= R;
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    scal[SC_RF] = (rf < eps) ? 0.0 : rf;
    Bm[(long long)j * work + j] = scal[SC_S];
  }
}

// --------------------------------------------------------------------------
// restart logic on the device (irlb.c after dgesdd + convtests), one thread block
//   BU (work x work), BS, res, svratio; k, converged -> ctl
//   new B: diag(BS[:k]) + column k = res[:k]
// --------------------------------------------------------------------------
__global__ void k_restart_logic(int work, int nu, double tol, double svtol, const double* __restrict__ BU,
                                const double* __restrict__ BS, double* __restrict__ svratio, double* __restrict__ res,
                                double* __restrict__ Bm, const double* __restrict__ scal, int* __restrict__ ctl) {
  __shared__ int s_len;
  __shared__ int s_k;
  if (ctl[CT_FLAG] & 2) return;  // the cycle broke down: leave svratio / B / k alone, the host re-runs it
  const double rf = scal[SC_RF];
  const double S = scal[SC_S];
  if (threadIdx.x == 0) s_len = 0;
  __syncthreads();
  double smax = 0.0;
  for (int i = 0; i < work; ++i) smax = fmax(smax, BS[i]);
  for (int i = threadIdx.x; i < work; i += blockDim.x) {
    double sr = fabs(svratio[i] - BS[i]) / BS[i];
    double r = rf * BU[(long long)i * work + (work - 1)];  // last ROW of BU
    res[i] = r;
    svratio[i] = sr;
    if (fabs(r) < tol * smax && sr < svtol) atomicAdd(&s_len, 1);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int k = ctl[CT_K];
    int len = s_len;
    int conv = (len >= nu || S == 0.0) ? 1 : 0;
    if (!conv) {
      if (k < nu + len) k = nu + len;
      if (k > work - 3) k = work - 3;
      if (k < 1) k = 1;
    }
    ctl[CT_K] = k;
    ctl[CT_CONV] = conv;
    s_k = k;
  }
  __syncthreads();
  if (ctl[CT_CONV]) return;
  const int k = s_k;
  for (int i = threadIdx.x; i < work; i += blockDim.x) svratio[i] = BS[i];
  for (int e = threadIdx.x; e < work * work; e += blockDim.x) Bm[e] = 0.0;
  __syncthreads();
  for (int i = threadIdx.x; i < k; i += blockDim.x) {
    Bm[(long long)i * work + i] = BS[i];
    Bm[(long long)k * work + i] = res[i];
  }
}

// Q[:, c] = BU[:, c] * BS[c]  (final embedding X = W * BU[:, :nu] * diag(d))
__global__ void k_scale_cols(const double* __restrict__ BU, const double* __restrict__ BS, int work, int nu,
                             double* __restrict__ Q) {
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e < work * nu) Q[e] = BU[e] * BS[e / work];
}

// transpose a small matrix (host-callback path hands Vt)
__global__ void k_transpose_small(const double* __restrict__ in, int n, double* __restrict__ out) {
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e < n * n) {
    int r = e % n, c = e / n;
    out[(long long)r * n + c] = in[(long long)c * n + r];
  }
}

// --------------------------------------------------------------------------
// K9: in-place tall-skinny GEMM  X[:, :kout] = X[:, :kin] * Q   (Q: kin x kout, ld kin)
// on the fp64 tensor-core path (DMMA, mma.sync.m8n8k4.f64) -- the one GEMM-shaped
// operation of this path, and the only tensor-core use: tcgen05 has no fp64 kind.
//
// Why tensor cores here: the FMA version of this kernel (round 1, profiles/
// r01_full_k_restart_gemm.txt) kept the fp64 pipe 36% busy and was bound by instruction
// issue and shared-memory operand traffic (67 instructions and 7.7 KB of LDS per 1664 FMA
// per warp); one DMMA does 256 FMA from two 8-byte operands per lane, i.e. ~6.5x fewer
// instructions and 8x less shared-memory traffic for the same flops.
//
// Persistent CTAs (one per SM, 128 threads = 4 warps).  Q is staged once per CTA as
// Qs[col][kinp]; the CTA walks 64-row blocks of X, each staged TRANSPOSED as Xs[row][kinp]
// with 8-byte cp.async, double-buffered (block i+1 lands while block i is multiplied), so the
// update can be done in place: a block's rows are complete in shared memory before its
// outputs are written.  16 warps = 4 row groups (16 rows) x 4 column groups (fragments cg,
// cg+4, ...): per k-step of 4 a warp loads 2 A and <= 4 B fragments (one LDS.64 per lane each)
// and issues <= 8 DMMAs; 4 warps per scheduler hide the DMMA latency (with 4 warps per SM the
// tensor pipe was 29% busy, profiles/).  kinp is padded to
// kinp % 16 == 4 doubles so that the 8 rows (or columns) of a fragment land on disjoint banks.
// --------------------------------------------------------------------------
#ifndef GEMM_RB
#define GEMM_RB 32                 // rows per block (16 per row group)
#endif
#define GEMM_RG (GEMM_RB / 16)     // row groups
#define GEMM_NF 13                 // B fragments (8 columns each) per pass
#define GEMM_CPP (8 * GEMM_NF)     // columns per pass = 104
#define GEMM_NCG (32 / GEMM_RG)     // column groups: 32 warps = GEMM_RG row groups x GEMM_NCG
#define GEMM_THREADS 1024
#define GEMM_FPW ((GEMM_NF + GEMM_NCG - 1) / GEMM_NCG)  // fragments per warp (max)

__device__ __forceinline__ void cp_async8(void* smem, const void* gmem) {
  unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(sa), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N));
}
__device__ __forceinline__ void dmma_m8n8k4(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

__device__ __forceinline__ void gemm_stage_rows(double* Xs, const double* X, long long ld, long long rows, long long r0,
                                                int kin, int kinp) {
  // Xs[r][k] = X[r0 + r, k]; global reads coalesced along r
  const int total = kin * GEMM_RB;
  for (int e = threadIdx.x; e < total; e += blockDim.x) {
    const int c = e / GEMM_RB, r = e % GEMM_RB;
    if (r0 + r < rows) cp_async8(Xs + r * kinp + c, X + (long long)c * ld + r0 + r);
  }
}

__global__ void __launch_bounds__(GEMM_THREADS, 1)
k_restart_gemm(double* __restrict__ X, long long ld, long long rows, int kin, int kinp, int kout,
               const double* __restrict__ Q, int nbuf) {
  extern __shared__ double smg[];
  double* Qs = smg;                               // [GEMM_CPP][kinp]
  double* Xs0 = smg + (size_t)GEMM_CPP * kinp;    // [GEMM_RB][kinp] x nbuf
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int gid = lane >> 2, tig = lane & 3;      // fragment coordinates
  const long long nblk = (rows + GEMM_RB - 1) / GEMM_RB;
  const int npass = (kout + GEMM_CPP - 1) / GEMM_CPP;
  const int ksteps = (kin + 3) >> 2;              // 4 * ksteps <= kinp
  // zero the k padding once (never written by the staging) -- also rows of Xs that stay unloaded
  for (int e = threadIdx.x; e < (GEMM_CPP + nbuf * GEMM_RB) * kinp; e += blockDim.x) smg[e] = 0.0;
  __syncthreads();
  auto stage_q = [&](int pass) {
    const int c0 = pass * GEMM_CPP;
    for (int e = threadIdx.x; e < GEMM_CPP * kin; e += blockDim.x) {
      const int c = e / kin, k = e % kin;
      Qs[c * kinp + k] = (c0 + c < kout) ? Q[(long long)(c0 + c) * kin + k] : 0.0;
    }
  };
  if (npass == 1) stage_q(0);
  long long blk = blockIdx.x;
  int buf = 0;
  if (blk < nblk) gemm_stage_rows(Xs0, X, ld, rows, blk * GEMM_RB, kin, kinp);
  cp_async_commit();
  for (; blk < nblk; blk += gridDim.x) {
    const long long nxt = blk + gridDim.x;
    double* Xs = Xs0 + (size_t)buf * GEMM_RB * kinp;
    if (nbuf == 2) {
      if (nxt < nblk) gemm_stage_rows(Xs0 + (size_t)(buf ^ 1) * GEMM_RB * kinp, X, ld, rows, nxt * GEMM_RB, kin, kinp);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    const int rg = warp % GEMM_RG, cg = warp / GEMM_RG;  // row group, column group
    const long long r0 = blk * GEMM_RB + rg * 16;       // this warp's 16 rows
    const double* xa0 = Xs + (size_t)(rg * 16 + gid) * kinp + tig;  // A fragment, rows 0-7
    const double* xa1 = xa0 + (size_t)8 * kinp;                     // rows 8-15
    for (int pass = 0; pass < npass; ++pass) {
      if (npass > 1) {
        __syncthreads();
        stage_q(pass);
        __syncthreads();
      }
      double c[2][GEMM_FPW][2];
#pragma unroll
      for (int f = 0; f < GEMM_FPW; ++f) c[0][f][0] = c[0][f][1] = c[1][f][0] = c[1][f][1] = 0.0;
      // B fragment slot f of this warp: fragment cg + GEMM_NCG f, i.e. column 8 (cg + NCG f) + gid
      const double* qb = Qs + (size_t)(8 * cg + gid) * kinp + tig;
#pragma unroll 2
      for (int ks = 0; ks < ksteps; ++ks) {
        const double a0 = xa0[4 * ks], a1 = xa1[4 * ks];
#pragma unroll
        for (int f = 0; f < GEMM_FPW; ++f) {
          if (cg + GEMM_NCG * f < GEMM_NF) {  // warp-uniform
            const double b = qb[(size_t)f * 8 * GEMM_NCG * kinp + 4 * ks];
            dmma_m8n8k4(c[0][f][0], c[0][f][1], a0, b);
            dmma_m8n8k4(c[1][f][0], c[1][f][1], a1, b);
          }
        }
      }
      // C fragment: row gid (+8), columns 8 (cg + NCG f) + 2 tig + {0,1}
#pragma unroll
      for (int f = 0; f < GEMM_FPW; ++f) {
        if (cg + GEMM_NCG * f < GEMM_NF) {
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int col = pass * GEMM_CPP + 8 * (cg + GEMM_NCG * f) + 2 * tig + h;
            if (col < kout) {
              if (r0 + gid < rows) X[(long long)col * ld + r0 + gid] = c[0][f][h];
              if (r0 + gid + 8 < rows) X[(long long)col * ld + r0 + gid + 8] = c[1][f][h];
            }
          }
        }
      }
    }
    __syncthreads();  // everyone is done with Xs[buf] before it is refilled
    if (nbuf == 1 && nxt < nblk) {
      gemm_stage_rows(Xs0, X, ld, rows, nxt * GEMM_RB, kin, kinp);
      cp_async_commit();
    }
    if (nbuf == 2) buf ^= 1;
  }
  cp_async_wait<0>();
}

// --------------------------------------------------------------------------
// K9, second version: the same in-place X[:, :kout] = X[:, :kin] Q^T with DMMA, re-tiled.
//
// What the first version (above) left on the table, measured at the 2M-cell shape where X no longer
// sits in L2: 4.1 ms per cell-side GEMM = 10 TFLOP/s and 0.8 TB/s -- neither the tensor pipe nor
// DRAM.  (a) A 32-row block is staged with 8-byte cp.async, one per element and thread: 256-byte
// runs scattered over kin DRAM pages; (b) with 16 column groups per 16 rows every warp re-reads the
// same A fragments: 3 LDS per 2 DMMA, 20 KB of shared-memory traffic per row.
// Here: a STAGE is 64 rows, kept COLUMN-major in shared memory (stride 72 doubles: the four k of a
// fragment land on disjoint bank halves), so a column of a stage is ONE contiguous 512-byte run in
// global memory and is fetched with ONE bulk async copy (cp.async.bulk + mbarrier complete_tx,
// issued by a single thread; two stages double-buffered); a CTA walks a CONTIGUOUS range of
// stages.  16 warps = 4 row groups (16 rows) x 4 column groups (4/3/3/3 fragments of 8 columns): 6 LDS
// per 8 DMMA, 6.5 KB of shared-memory traffic per row.  Requires kout <= 104 (one pass) and the
// tile to fit; otherwise the first version runs.
// Column starts that are only 8-byte aligned (odd ld): the copy starts one element early and the
// fragment loads of that column are shifted by one row.
// --------------------------------------------------------------------------
#define G2_ROWS 64
#define G2_RS 72
#define G2_THREADS 512
__device__ __forceinline__ void g2_mbar_init(unsigned bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void g2_expect_tx(unsigned bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void g2_bulk(unsigned dst, const void* src, unsigned bytes, unsigned bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
               "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void g2_wait(unsigned bar, unsigned parity) {
  unsigned ok;
  do {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
  } while (!ok);
}

__global__ void __launch_bounds__(G2_THREADS, 1)
k_restart_gemm2(double* __restrict__ X, long long ld, long long rows, int kin, int kinp, int kout,
                const double* __restrict__ Q, long long stages_per_cta) {
  extern __shared__ __align__(128) double smg2[];
  const int k4 = ((kin + 3) >> 2) << 2;
  double* Qs = smg2;                                   // [104][kinp]
  double* Xs0 = smg2 + (size_t)GEMM_CPP * kinp;        // [2][k4][G2_RS]
  __shared__ __align__(8) unsigned long long mbar[2];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int gid = lane >> 2, tig = lane & 3;
  const int rg = warp & 3, cg = warp >> 2;
  const int f0 = (cg == 0) ? 0 : 1 + 3 * cg, nf = (cg == 0) ? 4 : 3;   // fragments 0-3 | 4-6 | 7-9 | 10-12
  const unsigned bar0 = (unsigned)__cvta_generic_to_shared(&mbar[0]);
  const long long n_stages = (rows + G2_ROWS - 1) / G2_ROWS;
  const long long s_begin = (long long)blockIdx.x * stages_per_cta, s_end = min(n_stages, s_begin + stages_per_cta);
  if (s_begin >= s_end) return;
  const int odd = (int)(ld & 1);
  if (threadIdx.x == 0) {
    g2_mbar_init(bar0, 1);
    g2_mbar_init(bar0 + 8, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // Q (zero-padded in k and in the columns past kout) and the k padding of both stage buffers
  for (int e = threadIdx.x; e < GEMM_CPP * kinp; e += G2_THREADS) {
    const int c = e / kinp, k = e - c * kinp;
    Qs[e] = (c < kout && k < kin) ? Q[(long long)c * kin + k] : 0.0;
  }
  for (int e = threadIdx.x; e < 2 * (k4 - kin) * G2_RS; e += G2_THREADS) {
    const int b = e / ((k4 - kin) * G2_RS), q = e - b * (k4 - kin) * G2_RS;
    Xs0[(size_t)b * k4 * G2_RS + (size_t)kin * G2_RS + q] = 0.0;
  }
  __syncthreads();
  // issue the copies of a FULL stage (one thread); a partial last stage is copied by everybody at use
  auto prefetch = [&](long long st, int buf) {
    const long long r0 = st * G2_ROWS;
    if (r0 + G2_ROWS > rows) return;  // partial
    if (warp == 0) {  // the lanes of warp 0 share the columns; lane 0 announces the byte count first
      const unsigned bar = bar0 + 8u * buf;
      const unsigned dst0 = (unsigned)__cvta_generic_to_shared(Xs0 + (size_t)buf * k4 * G2_RS);
      const int n_odd = odd ? (kin >> 1) : 0;   // columns 1, 3, 5, ... start on an 8-byte boundary only
      if (lane == 0) g2_expect_tx(bar, (unsigned)((kin - n_odd) * (G2_ROWS * 8) + n_odd * (G2_ROWS * 8 + 16)));
      __syncwarp();
      for (int c = lane; c < kin; c += 32) {
        const double* src = X + (long long)c * ld + r0;
        if (odd && (c & 1)) g2_bulk(dst0 + (unsigned)(c * G2_RS * 8), src - 1, G2_ROWS * 8 + 16, bar);
        else g2_bulk(dst0 + (unsigned)(c * G2_RS * 8), src, G2_ROWS * 8, bar);
      }
    }
  };
  unsigned ph = 0u;  // bit b: parity of the next completion of buffer b's barrier
  prefetch(s_begin, 0);
  const int oa = (odd && (tig & 1)) ? 1 : 0;   // row shift of this lane's k columns (k = 4 ks + tig)
  for (long long st = s_begin; st < s_end; ++st) {
    const int buf = (int)((st - s_begin) & 1);
    if (st + 1 < s_end) prefetch(st + 1, buf ^ 1);
    const long long r0 = st * G2_ROWS;
    double* Xs = Xs0 + (size_t)buf * k4 * G2_RS;
    if (r0 + G2_ROWS > rows) {
      // partial stage: plain copies, same (possibly shifted) placement
      const int nr = (int)(rows - r0);
      for (int e = threadIdx.x; e < kin * G2_ROWS; e += G2_THREADS) {
        const int c = e / G2_ROWS, r = e - c * G2_ROWS;
        Xs[(size_t)c * G2_RS + r + ((odd && (c & 1)) ? 1 : 0)] = (r < nr) ? X[(long long)c * ld + r0 + r] : 0.0;
      }
      __syncthreads();
    } else {
      g2_wait(bar0 + 8u * buf, (ph >> buf) & 1u);
      ph ^= 1u << buf;
    }
    double c[2][4][2];
#pragma unroll
    for (int f = 0; f < 4; ++f) c[0][f][0] = c[0][f][1] = c[1][f][0] = c[1][f][1] = 0.0;
    const double* xa = Xs + (size_t)tig * G2_RS + 16 * rg + gid + oa;
    const double* qb = Qs + (size_t)(8 * f0 + gid) * kinp + tig;
    const int ksteps = k4 >> 2;
#pragma unroll 2
    for (int ks = 0; ks < ksteps; ++ks) {
      const double a0 = xa[(size_t)(4 * ks) * G2_RS], a1 = xa[(size_t)(4 * ks) * G2_RS + 8];
#pragma unroll
      for (int f = 0; f < 4; ++f) {
        if (f < nf) {  // warp-uniform
          const double b = qb[(size_t)f * 8 * kinp + 4 * ks];
          dmma_m8n8k4(c[0][f][0], c[0][f][1], a0, b);
          dmma_m8n8k4(c[1][f][0], c[1][f][1], a1, b);
        }
      }
    }
    __syncthreads();  // every warp has read its fragments of this stage: the rows may be overwritten,
                      // and the buffer refilled by the prefetch of the next iteration
    const long long rr = r0 + 16 * rg + gid;
#pragma unroll
    for (int f = 0; f < 4; ++f) {
      if (f < nf) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int col = 8 * (f0 + f) + 2 * tig + h;
          if (col < kout) {
            if (rr < rows) X[(long long)col * ld + rr] = c[0][f][h];
            if (rr + 8 < rows) X[(long long)col * ld + rr + 8] = c[1][f][h];
          }
        }
      }
    }
  }
}

int irlba_init(m3b_ctx* ctx) {
  CK(ctx, cudaFuncSetAttribute(k_restart_gemm2, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024));
  CK(ctx, cudaFuncSetAttribute(k_restart_gemm, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024));
  CK(ctx, cudaFuncSetAttribute(k_orth_fused, cudaFuncAttributeMaxDynamicSharedMemorySize, M3B_OF_SMEM_MAX));
  return M3B_OK;
}

// ==========================================================================
// host driver
// ==========================================================================
struct Irlba {
  m3b_ctx* ctx;
  m3b_mat* m;
  int n, work, nu;
  long long mrows;
  double *V = nullptr, *W = nullptr, *F = nullptr, *Fraw = nullptr, *xs = nullptr;
  double *s = nullptr, *ce = nullptr;  // nullable
  double *part = nullptr, *t = nullptr, *pn = nullptr, *ps = nullptr, *pb = nullptr, *buf2 = nullptr;
  double *scal = nullptr, *Bm = nullptr, *BU = nullptr, *BV = nullptr, *BS = nullptr, *svr = nullptr, *res = nullptr;
  double *Q = nullptr, *svd_scratch = nullptr;
  int* ctl = nullptr;
  unsigned int* xf_done = nullptr;  // CTA counter of the fused exchange kernel (self re-arming)
  // single-rank fused vector kernel (k_orth_fused)
  bool fused = false;
  int of_grid = 0, of_chunk_n = 0, of_chunk_m = 0;
  int pb_n = 0;  // valid entries of pb (k_make_xs: nb_n, k_orth_fused: of_grid)
  unsigned long long* bar = nullptr;
  unsigned long long bar_seq = 0;
  unsigned long long* xdiag = nullptr;  // [XD_COUNT] in-kernel exchange timing of the sharded fused kernel
  double* s_rms = nullptr;              // scale for center = FALSE, scale = TRUE (k_rms_scale)
  int nb_n, nb_m;        // blocks of the vector kernels (gene side / cell side)
  int nb_ca;             // blocks of k_combine_A (32 rows per block)
  int nbo_n, nbo_m;      // blocks of k_apply_orth (= number of pn/ps partials it writes)
  int chunk_n, chunk_m;  // rows per block of k_dots
  int nblk_n, nblk_m;
  double eps;
  std::vector<void*> owned;

  template <typename T>
  int alloc(T** p, size_t cnt) {
    RET(dev_alloc(ctx, p, cnt));
    owned.push_back(*p);
    return M3B_OK;
  }
  void release() {
    for (void* p : owned) dev_free(p);
    owned.clear();
  }

  static int pick_chunk(long long rows, int sms) {
    long long c = (rows + 2LL * sms - 1) / (2LL * sms);
    c = (c + 31) & ~31LL;
    c = std::max<long long>(256, std::min<long long>(c, 4096));
    return (int)c;
  }

  int orthog(double* X, long long ld, long long rows, int ncols, double* y, int chunk, int nblk, int nb, bool allreduce) {
    spmv_timer_begin(ctx, 2);
    if (ncols > 0) {
      k_dots<<<nblk, NT, chunk * sizeof(double), ctx->stream>>>(X, ld, rows, ncols, y, chunk, part);
      k_reduce_dots<<<(ncols * 32 + NT - 1) / NT, NT, 0, ctx->stream>>>(part, nblk, ncols, t);
      count_launch(ctx, 2);
      if (allreduce) RET(comm_allreduce_f64(ctx, t, ncols));
    }
    k_apply_orth<<<nb, NT, std::max(ncols, 1) * sizeof(double), ctx->stream>>>(X, ld, rows, ncols, t, y, pn, ps);
    count_launch(ctx);
    spmv_timer_end(ctx);
    CK(ctx, cudaGetLastError());
    return M3B_OK;
  }

  // one half Lanczos step's vector work after a product, single rank (see k_orth_fused)
  int orth_fused(int mode, int jcol, int ncols, bool sub_prev) {
    OrthArgs a;
    memset(&a, 0, sizeof(a));
    a.mode = mode;
    const bool wside = mode == OF_W;
    const TiledLayout& L = wside ? m->A : m->B;
    a.rows = wside ? mrows : n;
    a.chunk = wside ? of_chunk_m : of_chunk_n;
    a.partial = L.partial;
    a.rsp = L.row_slot_ptr;
    a.s = s;
    a.ce = ce;
    if (wside) {
      a.pb_in = ce ? pb : nullptr;
      a.npb_in = pb_n;
      a.y = W + (long long)jcol * mrows;
      a.wprev = sub_prev ? a.y - mrows : nullptr;
      a.X = W;
      a.ld = mrows;
    } else {
      a.vj = V + (long long)jcol * n;
      a.y = F;
      a.X = V;
      a.ld = n;
      a.vnext = V + (long long)(jcol + 1) * n;  // unused by OF_FLAST
      a.xs = xs;
      a.pb_out = pb;
    }
    a.ncols = ncols;
    a.part = part;
    a.pn = pn;
    a.ps = ps;
    a.Bm = Bm;
    a.work = work;
    a.j = jcol;
    a.eps = eps;
    a.scal = scal;
    a.ctl = ctl;
    a.bar = bar;
    a.bar_base = bar_seq;
    a.world = ctx->comm.world;
    a.rank = ctx->comm.rank;
    a.peers = ctx->comm.peers_dev;
    a.timeout = ctx->spin_timeout_cycles;
    int n_bar = 2;
    if (a.world > 1) {
      // exchanges of this launch: F side 1 (the gene vector), W side 2 (coefficients, then norm and sum);
      // the W side's second grid barrier is the all-to-all of the norm partials itself
      a.seq = ctx->comm.xchg_seq + 1;
      ctx->comm.xchg_seq += wside ? 2 : 1;
      a.xpar = (int)(ctx->comm.ll_launch[wside ? 1 : 0]++ & 1u);
      a.diag = xdiag;
      if (wside) n_bar = 1;
    }
    bar_seq += (unsigned long long)n_bar * of_grid;
    void* args[] = {&a};
    const size_t smem = ((size_t)a.chunk + ((ncols + 1) & ~1) + 1024) * sizeof(double);
    spmv_timer_begin(ctx, 2);
    CK(ctx, cudaLaunchCooperativeKernel((void*)k_orth_fused, dim3(of_grid), dim3(OF_THREADS), args, smem, ctx->stream));
    count_launch(ctx);
    spmv_timer_end(ctx);
    if (mode == OF_F) pb_n = of_grid;
    return M3B_OK;
  }

  // W[:, jw] = Ax(V[:, jv]) given xs/pb already prepared; optional "- R_F * W[:, jw-1]"
  int product_A(int jw, bool sub_prev, bool prelaunched = false, bool combine = true) {
    if (!prelaunched) {
      spmv_timer_begin(ctx, 0);
      RET(launch_tile_stream(ctx, m->A, OP_DOT, xs));
      spmv_timer_end(ctx);
    }
    ctx->stats.spmv_cells_out++;
    if (!combine) return M3B_OK;
    double* w = W + (long long)jw * mrows;
    spmv_timer_begin(ctx, 5);
    if (mrows) {
      k_combine_A<<<nb_ca, NT, 0, ctx->stream>>>(m->A.partial, m->A.row_slot_ptr, mrows, ce ? pb : nullptr,
                                                pb_n, sub_prev ? (w - mrows) : nullptr, scal, w);
      count_launch(ctx);
    }
    spmv_timer_end(ctx);
    CK(ctx, cudaGetLastError());
    return M3B_OK;
  }

  // F = Atw(W[:, j]) - S * V[:, j]
  int product_B(int j, bool combine = true) {
    spmv_timer_begin(ctx, 1);
    RET(launch_tile_stream(ctx, m->B, OP_DOT, W + (long long)j * mrows));
    spmv_timer_end(ctx);
    ctx->stats.spmv_genes_out++;
    if (!combine) return M3B_OK;
    if (ctx->comm.world == 1) {
      spmv_timer_begin(ctx, 5);
      k_combine_fix_F<<<std::max(1, std::min((n + 31) / 32, 8 * ctx->sm_count)), NT, 0, ctx->stream>>>(m->B.partial, m->B.row_slot_ptr, n, s, ce,
                                                    V + (long long)j * n, scal, F);
      count_launch(ctx);
      spmv_timer_end(ctx);
      CK(ctx, cudaGetLastError());
      return M3B_OK;
    }
    if (ctx->comm.p2p && n <= M3B_XCHG_CAP) {
      spmv_timer_begin(ctx, 6);
      ctx->comm.xchg_seq++;
      k_combine_xchg_fix_F<<<std::max(1, std::min(XF_CTAS, (n + 31) / 32)), NT, 0, ctx->stream>>>(
          m->B.partial, m->B.row_slot_ptr, n, s, ce, V + (long long)j * n, scal, F, ctx->comm.peers_dev,
          ctx->comm.rank, ctx->comm.world, ctx->comm.xchg_seq, xf_done, ctl, ctx->spin_timeout_cycles);
      count_launch(ctx);
      spmv_timer_end(ctx);
      CK(ctx, cudaGetLastError());
      return M3B_OK;
    }
    spmv_timer_begin(ctx, 5);
    RET(launch_combine_rows(ctx, m->B, Fraw));
    spmv_timer_end(ctx);
    if (ctx->comm.world > 1) {
      spmv_timer_begin(ctx, 6);
      RET(comm_allreduce_f64(ctx, Fraw, n));
      spmv_timer_end(ctx);
    }
    spmv_timer_begin(ctx, 5);
    k_fix_F<<<nb_n, NT, 0, ctx->stream>>>(Fraw, s, ce, V + (long long)j * n, scal, n, F);
    count_launch(ctx);
    spmv_timer_end(ctx);
    CK(ctx, cudaGetLastError());
    return M3B_OK;
  }

  // normalise W[:, jw] using pn/ps partials (cell side): S, sum(w)
  // ||y||^2 from the partials the last orthog / norm kernel left in pn (host value; slow path only)
  int read_norm2(int np, double* out, bool sharded_rows = false) {
    k_reduce2<<<1, NT, 0, ctx->stream>>>(pn, ps, np, buf2);
    count_launch(ctx);
    if (sharded_rows) RET(comm_allreduce_f64(ctx, buf2, 2));  // cell side of a sharded run: sum over the ranks
    CK(ctx, cudaMemcpyAsync(out, buf2, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CK(ctx, cudaStreamSynchronize(ctx->stream));
    return M3B_OK;
  }

  // irlb.c's answer to an invariant subspace: y = rnorm(len) from the caller's stream, y -= X X^T y
  // Sharded runs: every rank's callback continues the SAME stream (they all drew the start vector
  // from it), so a gene-length vector is simply drawn everywhere (replicas stay identical), and for
  // a cell-length vector every rank draws all n_cells_total values and keeps its own cell range --
  // the streams stay in step and the result is the unsharded run's vector, sliced.
  long long cell_offset = -1;  // first global cell of this rank (lazily: one allreduce)
  int find_cell_offset() {
    if (cell_offset >= 0) return M3B_OK;
    cell_offset = 0;
    const int world = ctx->comm.world;
    if (world == 1) return M3B_OK;
    long long* d = nullptr;
    RET(dev_alloc(ctx, &d, (size_t)world));
    std::vector<long long> h((size_t)world, 0);
    h[ctx->comm.rank] = mrows;
    cudaError_t e = cudaMemcpyAsync(d, h.data(), world * sizeof(long long), cudaMemcpyHostToDevice, ctx->stream);
    int st = e == cudaSuccess ? comm_allreduce_i64(ctx, d, (size_t)world) : M3B_E_CUDA;
    if (st == M3B_OK) e = cudaMemcpyAsync(h.data(), d, world * sizeof(long long), cudaMemcpyDeviceToHost, ctx->stream);
    if (st == M3B_OK && e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    dev_free(d);
    if (st != M3B_OK) return st;
    CK(ctx, e);
    for (int q = 0; q < ctx->comm.rank; ++q) cell_offset += h[q];
    return M3B_OK;
  }
  int random_vector(double* X, long long ld, long long rows, int ncols, double* y, int chunk, int nblk, int nb,
                    bool sharded_rows = false) {
    long long total = rows, off = 0;
    if (sharded_rows && ctx->comm.world > 1) {
      RET(find_cell_offset());
      total = m->n_cells_total;
      off = cell_offset;
    }
    std::vector<double> h((size_t)std::max<long long>(total, 1));
    if (ctx->rnorm((int64_t)total, h.data(), ctx->rnorm_user) != 0) {
      m3b_set_error(ctx, "rnorm callback failed");
      return M3B_E_STATE;
    }
    if (rows) CK(ctx, cudaMemcpyAsync(y, h.data() + off, (size_t)rows * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    CK(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->stats.h2d_bytes += rows * 8;
    return orthog(X, ld, rows, ncols, y, chunk, nblk, nb, sharded_rows && ctx->comm.world > 1);
  }

  int finish_w(int jw, bool first, int np, bool force_zero = false) {
    if (ctx->comm.world == 1) {
      spmv_timer_begin(ctx, 5);
      k_scale_w_fused<<<nb_m, NT, 0, ctx->stream>>>(W + (long long)jw * mrows, mrows, pn, ps, np, scal, ctl, eps,
                                                    first ? 1 : 0, force_zero ? 1 : 0);
      count_launch(ctx);
      spmv_timer_end(ctx);
      CK(ctx, cudaGetLastError());
      return M3B_OK;
    }
    spmv_timer_begin(ctx, 5);
    k_reduce2<<<1, NT, 0, ctx->stream>>>(pn, ps, np, buf2);
    count_launch(ctx);
    RET(comm_allreduce_f64(ctx, buf2, 2));
    k_scale_w<<<nb_m, NT, 0, ctx->stream>>>(W + (long long)jw * mrows, mrows, buf2, scal, ctl, eps, first ? 1 : 0,
                                            force_zero ? 1 : 0);
    count_launch(ctx);
    spmv_timer_end(ctx);
    CK(ctx, cudaGetLastError());
    return M3B_OK;
  }

  int gemm_inplace(double* X, long long ld, long long rows, int kin, int kout, const double* Qd) {
    if (rows == 0 || kout == 0) return M3B_OK;
    const size_t budget = 224 * 1024;
    int kinp = ((kin + 3) & ~3);
    while ((kinp & 15) != 4) kinp += 4;  // kinp % 16 == 4: the 8 rows of a fragment hit disjoint banks
    {
      // second version (bulk-copied 64-row stages, 4 x 4 warp grid) whenever it applies
      static const bool no_v2 = getenv("M3B_GEMM_V1") != nullptr;
      const int k4 = (kin + 3) & ~3;
      const size_t smem2 = ((size_t)GEMM_CPP * kinp + 2 * (size_t)k4 * G2_RS) * sizeof(double);
      if (!no_v2 && kout <= GEMM_CPP && smem2 <= (size_t)226 * 1024 - 64 && rows >= G2_ROWS) {
        const long long n_stages = (rows + G2_ROWS - 1) / G2_ROWS;
        const long long per = (n_stages + ctx->sm_count - 1) / ctx->sm_count;
        const unsigned grid = (unsigned)((n_stages + per - 1) / per);
        k_restart_gemm2<<<grid, G2_THREADS, smem2, ctx->stream>>>(X, ld, rows, kin, kinp, kout, Qd, per);
        count_launch(ctx);
        CK(ctx, cudaGetLastError());
        return M3B_OK;
      }
    }
    int nbuf = 2;
    size_t smem = (size_t)(GEMM_CPP + nbuf * GEMM_RB) * kinp * sizeof(double);
    if (smem > budget) {
      nbuf = 1;
      smem = (size_t)(GEMM_CPP + GEMM_RB) * kinp * sizeof(double);
    }
    if (smem > budget) {
      m3b_set_error(ctx, "restart GEMM: work=%d too large for the shared-memory tiling (max ~160)", kin);
      return M3B_E_DIMS;
    }
    const long long nblk = (rows + GEMM_RB - 1) / GEMM_RB;
    const unsigned grid = (unsigned)std::min<long long>(nblk, ctx->sm_count);
    k_restart_gemm<<<grid, GEMM_THREADS, smem, ctx->stream>>>(X, ld, rows, kin, kinp, kout, Qd, nbuf);
    count_launch(ctx);
    CK(ctx, cudaGetLastError());
    return M3B_OK;
  }
};

// --------------------------------------------------------------------------
// K3: gene moments on layout B (R/utils.R:383,389).  Two passes, fp64:
//   mean_g = sum_c y / C ;  var_g = (sum_nz (y-mean)^2 + (C - nnz_g) mean^2) / (C-1)
// Pad entries are stored zeros, so pass 2 counts them as zeros: the closed-form
// term uses the padded run length instead of nnz_g.  Sharded: both sums are
// allreduced.  `center` holds the mean of the implicit dense matrix (mean + f0).
// --------------------------------------------------------------------------
__global__ void k_mean_from_sum(const double* __restrict__ sum, int n, double invC, double f0, double* __restrict__ mean_sparse,
                                double* __restrict__ center) {
  int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g < n) {
    double mu = sum[g] * invC;
    mean_sparse[g] = mu;
    center[g] = mu + f0;
  }
}
__global__ void k_ss_local(double* __restrict__ ss, const double* __restrict__ mean_sparse,
                           const long long* __restrict__ padlen, long long c_local, int n) {
  int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g < n) {
    double mu = mean_sparse[g];
    ss[g] += (double)(c_local - padlen[g]) * mu * mu;
  }
}
__global__ void k_sd_from_ss(const double* __restrict__ ss, int n, double denom, double* __restrict__ sd) {
  int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g < n) sd[g] = sqrt(ss[g] / denom);
}

int compute_moments(m3b_mat* m) {
  m3b_ctx* ctx = m->ctx;
  if (m->moments_done) return M3B_OK;
  const int n = m->n_kept;
  dev_free(m->center);
  dev_free(m->scale);
  dev_free(m->mean_sparse);
  RET(dev_alloc(ctx, &m->center, (size_t)std::max(n, 1)));
  RET(dev_alloc(ctx, &m->scale, (size_t)std::max(n, 1)));
  RET(dev_alloc(ctx, &m->mean_sparse, (size_t)std::max(n, 1)));
  double* tmp = nullptr;
  RET(dev_alloc(ctx, &tmp, (size_t)std::max(n, 1)));
  int st = M3B_OK;
  if (n) {
    const int nb = (n + 255) / 256;
    do {  // (tmp is released on every exit)
      if ((st = launch_tile_stream(ctx, m->B, OP_SUM, nullptr)) != M3B_OK) break;
      if ((st = launch_combine_rows(ctx, m->B, tmp)) != M3B_OK) break;
      if ((st = comm_allreduce_f64(ctx, tmp, n)) != M3B_OK) break;
      k_mean_from_sum<<<nb, 256, 0, ctx->stream>>>(tmp, n, 1.0 / (double)m->n_cells_total, m->f0, m->mean_sparse, m->center);
      if ((st = launch_tile_stream(ctx, m->B, OP_SQDEV, m->mean_sparse)) != M3B_OK) break;
      if ((st = launch_combine_rows(ctx, m->B, tmp)) != M3B_OK) break;
      k_ss_local<<<nb, 256, 0, ctx->stream>>>(tmp, m->mean_sparse, m->kept_padlen, m->n_cells, n);
      if ((st = comm_allreduce_f64(ctx, tmp, n)) != M3B_OK) break;
      k_sd_from_ss<<<nb, 256, 0, ctx->stream>>>(tmp, n, (double)(m->n_cells_total - 1), m->scale);
      count_launch(ctx, 3);
    } while (0);
  }
  cudaError_t e = cudaStreamSynchronize(ctx->stream);
  if (e == cudaSuccess) e = cudaGetLastError();
  dev_free(tmp);
  if (st != M3B_OK) return st;
  if (e != cudaSuccess) {
    m3b_set_error(ctx, "gene moments: %s", cudaGetErrorString(e));
    return M3B_E_CUDA;
  }
  m->moments_done = true;
  return M3B_OK;
}

// sparse_prcomp_irlba with center = FALSE, scale. = TRUE (R/utils.R:395-397): the scale is the root mean
// square about ZERO, sqrt(colSums(x^2) / max(1, nrow - 1)), not the standard deviation.  From the
// moments already on the device: sum x^2 = (C - 1) sd^2 + C mean^2.
__global__ void k_rms_scale(const double* __restrict__ mean, const double* __restrict__ sd, int n, double C,
                            double* __restrict__ out) {
  int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g < n) {
    const double d = fmax(1.0, C - 1.0);
    out[g] = sqrt(((C - 1.0) * sd[g] * sd[g] + C * mean[g] * mean[g]) / d);
  }
}

__global__ void k_fill(double* p, int n, double v) {
  int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g < n) p[g] = v;
}

int run_irlba(m3b_mat* m, int nv, int work, int maxit, double tol, double svtol, const double* v0, int center,
              int scale, double* d_out, double* V_out, double* X_out, long long ldx, double* center_out,
              double* scale_out, int* iter_out, int* mprod_out, bool copy_x) {
  m3b_ctx* ctx = m->ctx;
  const int n = m->n_kept;
  const long long mloc = m->n_cells, mtot = m->n_cells_total;
  const long long minmn = std::min<long long>(mtot, n);
  if (nv <= 0 || nv >= minmn) {
    m3b_set_error(ctx, "max(nu, nv) must be strictly less than min(nrow(A), ncol(A)) (nv=%d, min dim=%lld)", nv, minmn);
    return M3B_E_DIMS;
  }
  // scale without center: sparse_prcomp_irlba scales by the root mean square sqrt(colSums(x^2) / (n - 1))
  // (R/utils.R:396-400), not by the standard deviation -- see k_rms_scale.  preprocess_cds never asks
  // for that (it passes `scaling` for both); the dense tiny-problem branch does not implement it.
  if (scale && !center && minmn < 6) {
    m3b_set_error(ctx, "m3b_pca_irlba: scale without center is not supported on the dense branch (min dimension < 6)");
    return M3B_E_DIMS;
  }
  if (minmn < 6)  // irlba: "Tiny problem detected, using standard `svd` function" 
    return run_irlba_tiny(m, nv, center, scale, d_out, V_out, X_out, ldx, center_out, scale_out, iter_out, mprod_out, copy_x);
  if (work <= 0) work = nv + 7;
  if (work <= nv) work = nv + 1;
  if (work > minmn) work = (int)minmn;
  if (work < 4 || n < 4 || mtot < 4) {
    m3b_set_error(ctx, "irlba needs work, genes and cells >= 4 (work=%d, genes=%d, cells=%lld)", work, n, mtot);
    return M3B_E_DIMS;
  }
  if (copy_x && X_out && ldx < mloc) {
    m3b_set_error(ctx, "ldx (%lld) < local cells (%lld)", ldx, mloc);
    return M3B_E_DIMS;
  }
  if (svtol <= 0) svtol = tol;
  const bool use_scale = scale != 0;
  const bool use_center = center != 0;
  if (use_scale || use_center) RET(compute_moments(m));
  // an earlier run that ended in an error on one rank leaves the ranks' exchange counters apart
  RET(comm_xchg_resync(ctx));

  Irlba I;
  I.ctx = ctx;
  I.m = m;
  I.n = n;
  I.work = work;
  I.nu = nv;
  I.mrows = mloc;
  I.eps = 2.220446049250313e-16;
  I.nb_n = std::max(1, std::min((n + NT - 1) / NT, 64));
  I.nb_m = (int)std::max<long long>(1, std::min<long long>((mloc + NT - 1) / NT, MAX_PART));
  I.nb_ca = (int)std::max<long long>(1, std::min<long long>((mloc + 31) / 32, 8LL * ctx->sm_count));
  I.nbo_n = std::max(1, std::min((n + 63) / 64, MAX_PART));
  I.nbo_m = (int)std::max<long long>(1, std::min<long long>((mloc + 63) / 64, MAX_PART));
  I.chunk_n = Irlba::pick_chunk(n, ctx->sm_count);
  I.chunk_m = Irlba::pick_chunk(mloc, ctx->sm_count);
  I.nblk_n = std::max(1, (n + I.chunk_n - 1) / I.chunk_n);
  I.nblk_m = (int)std::max<long long>(1, (mloc + I.chunk_m - 1) / I.chunk_m);
  int status = M3B_OK;
  cudaEvent_t evA = nullptr, evB = nullptr, evF = nullptr, evK4 = nullptr;
  auto now = []() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
  const double t_enter = now();
  double t_loop0 = 0, t_loop1 = 0, t_fin = 0;
#define IR(expr)               \
  do {                         \
    status = (expr);           \
    if (status != M3B_OK) goto done; \
  } while (0)
#define ICK(call)                                                                               \
  do {                                                                                          \
    cudaError_t e__ = (call);                                                                   \
    if (e__ != cudaSuccess) {                                                                   \
      m3b_set_error(ctx, "%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
      status = (e__ == cudaErrorMemoryAllocation) ? M3B_E_NOMEM : M3B_E_CUDA;                   \
      goto done;                                                                                \
    }                                                                                           \
  } while (0)
  {
    std::vector<double> hv(n);
    std::vector<double> hB((size_t)work * work), hU((size_t)work * work), hS(work), hVt((size_t)work * work);
    int hctl[CT_COUNT];
    int it = 0, mprod = 0, k = 0;
    bool converged = false;
    // fused vector kernel: single rank, one CTA per SM, every CTA owns ceil(rows / grid) rows
    I.of_grid = ctx->sm_count;
    I.of_chunk_n = (n + I.of_grid - 1) / I.of_grid;
    I.of_chunk_m = (int)((mloc + I.of_grid - 1) / I.of_grid);
    {
      const size_t smem = ((size_t)std::max(I.of_chunk_n, I.of_chunk_m) + work + 2 + 1024) * sizeof(double);
      // sharded runs: only with the peer-memory exchange (checked against the oracle at world 2 and 8,
      // tests/test_gpu_multi.py; M3B_NO_MR_FUSED=1 falls back to the separate kernels with NCCL between them)
      const bool rank_ok = ctx->comm.world == 1 ||
                           (ctx->comm.p2p && n <= M3B_LL_CAPF && work <= M3B_LL_CAPW && I.of_grid <= M3B_LL_CAPG &&
                            !getenv("M3B_NO_MR_FUSED"));
      I.fused = rank_ok && work <= 1024 && I.of_grid <= 1024 && I.of_grid <= MAX_PART &&
                smem <= (size_t)M3B_OF_SMEM_MAX && !getenv("M3B_NO_FUSED_ORTH");
    }
    const size_t part_sz = (size_t)std::max(std::max(I.nblk_n, I.nblk_m), I.of_grid) * work;
    IR(I.alloc(&I.V, (size_t)n * work + 8));
    IR(I.alloc(&I.W, (size_t)std::max<long long>(mloc, 1) * work + 8));  // + slack: see k_restart_gemm2 (odd ld)
    IR(I.alloc(&I.F, (size_t)n));
    IR(I.alloc(&I.Fraw, (size_t)n + 2));
    IR(I.alloc(&I.xs, (size_t)n));
    IR(I.alloc(&I.part, part_sz));
    IR(I.alloc(&I.t, (size_t)work));
    IR(I.alloc(&I.pn, (size_t)MAX_PART));
    IR(I.alloc(&I.ps, (size_t)MAX_PART));
    IR(I.alloc(&I.pb, (size_t)1024));
    IR(I.alloc(&I.buf2, (size_t)4));
    IR(I.alloc(&I.scal, (size_t)SC_COUNT));
    IR(I.alloc(&I.Bm, (size_t)work * work));
    IR(I.alloc(&I.BU, (size_t)work * work));
    IR(I.alloc(&I.BV, (size_t)work * work));
    IR(I.alloc(&I.BS, (size_t)work));
    IR(I.alloc(&I.svr, (size_t)work));
    IR(I.alloc(&I.res, (size_t)work));
    IR(I.alloc(&I.Q, (size_t)work * work));
    IR(I.alloc(&I.svd_scratch, (size_t)4 * work * work + 10 * work));
    IR(I.alloc(&I.ctl, (size_t)CT_COUNT));
    IR(I.alloc(&I.xf_done, (size_t)4));
    ICK(cudaMemsetAsync(I.xf_done, 0, 4 * sizeof(unsigned int), ctx->stream));
    ICK(cudaMemsetAsync(I.Bm, 0, (size_t)work * work * sizeof(double), ctx->stream));
    ICK(cudaMemsetAsync(I.svr, 0, work * sizeof(double), ctx->stream));
    ICK(cudaMemsetAsync(I.scal, 0, SC_COUNT * sizeof(double), ctx->stream));
    ICK(cudaMemsetAsync(I.ctl, 0, CT_COUNT * sizeof(int), ctx->stream));
    ICK(cudaMemsetAsync(I.pb, 0, 1024 * sizeof(double), ctx->stream));
    IR(I.alloc(&I.bar, (size_t)2));
    ICK(cudaMemsetAsync(I.bar, 0, 2 * sizeof(unsigned long long), ctx->stream));
    IR(I.alloc(&I.xdiag, (size_t)XD_COUNT));
    ICK(cudaMemsetAsync(I.xdiag, 0, XD_COUNT * sizeof(unsigned long long), ctx->stream));
    // operator vectors
    I.s = use_scale ? m->scale : nullptr;
    if (use_scale && !use_center) {
      IR(I.alloc(&I.s_rms, (size_t)n));
      k_rms_scale<<<(n + 255) / 256, 256, 0, ctx->stream>>>(m->center, m->scale, n, (double)m->n_cells_total, I.s_rms);
      count_launch(ctx);
      I.s = I.s_rms;
    }
    if (use_center) {
      I.ce = m->mean_sparse;  // mean(D) - f0 = mean of the stored sparse part
    } else if (m->f0 != 0.0) {
      IR(I.alloc(&I.ce, (size_t)n));
      k_fill<<<(n + 255) / 256, 256, 0, ctx->stream>>>(I.ce, n, -m->f0);
      count_launch(ctx);
    } else {
      I.ce = nullptr;
    }
    // start vector: V[:,0] = v0 / ||v0||   (irlb.c: dnrm2 + dscal)
    {
      long double nn = 0;
      for (int g = 0; g < n; ++g) nn += (long double)v0[g] * v0[g];
      double dn = std::sqrt((double)nn);
      if (!(dn >= I.eps)) {
        m3b_set_error(ctx, "start vector has zero norm");
        status = M3B_E_DIMS;
        goto done;
      }
      double inv = 1.0 / dn;
      for (int g = 0; g < n; ++g) hv[g] = v0[g] * inv;
      ICK(cudaMemcpyAsync(I.V, hv.data(), n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
      ctx->stats.h2d_bytes += n * 8;
    }
    ICK(cudaEventCreate(&evA));
    ICK(cudaEventCreate(&evB));
    ICK(cudaEventCreateWithFlags(&evF, cudaEventDisableTiming));
    ICK(cudaEventCreateWithFlags(&evK4, cudaEventDisableTiming));
    ICK(cudaEventRecord(evA, ctx->stream));
    t_loop0 = now();

    bool pre_k4 = false;  // next cycle's first product already running on the side stream
    bool careful = false; // this cycle is the re-run of one that hit an invariant subspace
    while (it < maxit) {
      int j = (it == 0) ? 0 : k;
      // W[:,j] = Ax(V[:,j])
      if (pre_k4) {
        ICK(cudaStreamWaitEvent(ctx->stream, evK4, 0));
      } else {
        k_make_xs<<<I.nb_n, NT, 0, ctx->stream>>>(I.V + (long long)j * n, I.s, I.ce, n, I.xs, I.pb);
        count_launch(ctx);
      }
      I.pb_n = I.nb_n;
      const bool fz = I.fused && !careful;
      const int mprod_cycle0 = mprod;
      IR(I.product_A(j, false, pre_k4, !(fz && it > 0)));
      pre_k4 = false;
      mprod++;
      if (fz && it > 0) {
        IR(I.orth_fused(OF_W, j, j, false));
      } else {
        if (it > 0) {
          IR(I.orthog(I.W, mloc, mloc, j, I.W + (long long)j * mloc, I.chunk_m, I.nblk_m, I.nbo_m, true));
        } else {
          k_norm_partials<<<I.nb_m, NT, 0, ctx->stream>>>(I.W + (long long)j * mloc, mloc, I.pn, I.ps);
          count_launch(ctx);
        }
        IR(I.finish_w(j, it == 0 && j == 0, it > 0 ? I.nbo_m : I.nb_m));
      }
      // Lanczos process
      while (j < work) {
        if (fz) {
          IR(I.product_B(j, false));
          mprod++;
          if (j + 1 < work) {
            IR(I.orth_fused(OF_F, j, j + 1, false));
            IR(I.product_A(j + 1, true, false, false));
            mprod++;
            IR(I.orth_fused(OF_W, j + 1, j + 1, true));
          } else {
            IR(I.orth_fused(OF_FLAST, j, j + 1, false));
          }
          j++;
          continue;
        }
        IR(I.product_B(j));
        mprod++;
        IR(I.orthog(I.V, n, n, j + 1, I.F, I.chunk_n, I.nblk_n, I.nbo_n, false));
        if (j + 1 < work) {
          // careful mode (re-run of a cycle that broke down): look at every norm on the host and
          // replace a vanished vector by a random one, as irlb.c does
          bool rnd = false;
          if (careful) {
            double n2 = 0.0;
            IR(I.read_norm2(I.nbo_n, &n2));
            if (!(std::sqrt(n2) >= I.eps)) {
              rnd = true;
              IR(I.random_vector(I.V, n, n, j + 1, I.F, I.chunk_n, I.nblk_n, I.nbo_n));
            }
          }
          k_finish_F<<<I.nb_n, NT, 0, ctx->stream>>>(I.F, I.pn, I.nbo_n, I.s, I.ce, n, I.V + (long long)(j + 1) * n, I.xs,
                                                     I.pb, I.Bm, work, j, I.scal, I.ctl, I.eps, rnd ? 1 : 0);
          count_launch(ctx);
          IR(I.product_A(j + 1, true));
          mprod++;
          IR(I.orthog(I.W, mloc, mloc, j + 1, I.W + (long long)(j + 1) * mloc, I.chunk_m, I.nblk_m, I.nbo_m, true));
          rnd = false;
          if (careful) {
            double n2 = 0.0;
            IR(I.read_norm2(I.nbo_m, &n2, true));
            if (!(std::sqrt(n2) >= I.eps)) {
              rnd = true;
              IR(I.random_vector(I.W, mloc, mloc, j + 1, I.W + (long long)(j + 1) * mloc, I.chunk_m, I.nblk_m, I.nbo_m, true));
            }
          }
          IR(I.finish_w(j + 1, false, I.nbo_m, rnd));
        } else {
          k_finish_last<<<I.nb_n, NT, 0, ctx->stream>>>(I.F, I.pn, I.nbo_n, n, I.Bm, work, j, I.scal, I.eps);
          count_launch(ctx);
        }
        j++;
      }
      ICK(cudaGetLastError());
      // The first product of the NEXT cycle is A * (F / s) with F the restart vector, which is
      // final before the SVD: run it on the side stream underneath the (latency-bound, few-SM)
      // restart SVD.  It takes all but 28 SMs so that the SVD kernels find free ones.  If this
      // restart turns out to be the last one the product is simply not used.
      if (!ctx->small_svd && flat_has_plan(m->A, ctx->sm_count - M3B_SIDE_CTA_RESERVE) && ctx->sm_count > 56 && it + 1 < maxit) {
        ICK(cudaEventRecord(evF, ctx->stream));
        ICK(cudaStreamWaitEvent(ctx->stream2, evF, 0));
        k_make_xs<<<I.nb_n, NT, 0, ctx->stream2>>>(I.F, I.s, I.ce, n, I.xs, I.pb);
        count_launch(ctx);
        spmv_timer_begin(ctx, 0, ctx->stream2);
        IR(launch_tile_stream(ctx, m->A, OP_DOT, I.xs, ctx->stream2, ctx->sm_count - M3B_SIDE_CTA_RESERVE));
        spmv_timer_end(ctx, ctx->stream2);
        ICK(cudaEventRecord(evK4, ctx->stream2));
        pre_k4 = true;
      }
      // small SVD of B
      spmv_timer_begin(ctx, 3);
      if (ctx->small_svd) {
        ICK(cudaMemcpyAsync(hB.data(), I.Bm, hB.size() * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        ICK(cudaStreamSynchronize(ctx->stream));
        if (ctx->small_svd(work, hB.data(), hU.data(), hS.data(), hVt.data(), ctx->small_svd_user) != 0) {
          m3b_set_error(ctx, "small_svd callback failed");
          status = M3B_E_STATE;
          goto done;
        }
        ICK(cudaMemcpyAsync(I.BU, hU.data(), hU.size() * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
        ICK(cudaMemcpyAsync(I.BS, hS.data(), hS.size() * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
        ICK(cudaMemcpyAsync(I.Q, hVt.data(), hVt.size() * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
        k_transpose_small<<<(work * work + 255) / 256, 256, 0, ctx->stream>>>(I.Q, work, I.BV);
        count_launch(ctx);
      } else if (it > 0 && k >= 1 && !(ctx->flags & M3B_FLAG_JACOBI_ONLY)) {
        // structured fast path; the Jacobi kernel runs only if it raised its flag
        IR(launch_arrow_svd(ctx, work, k, I.ctl + CT_K, I.Bm, I.BU, I.BS, I.BV, I.svd_scratch, I.ctl + CT_SVDFLAG));
        IR(launch_jacobi_svd(ctx, work, I.Bm, I.BU, I.BS, I.BV, I.svd_scratch, I.ctl + CT_SWEEPS, I.ctl + CT_SVDFLAG));
      } else {
        IR(launch_jacobi_svd(ctx, work, I.Bm, I.BU, I.BS, I.BV, I.svd_scratch, I.ctl + CT_SWEEPS, nullptr));
      }
      k_restart_logic<<<1, 128, 0, ctx->stream>>>(work, nv, tol, svtol, I.BU, I.BS, I.svr, I.res, I.Bm, I.scal, I.ctl);
      count_launch(ctx);
      spmv_timer_end(ctx);
      ICK(cudaMemcpyAsync(hctl, I.ctl, sizeof(hctl), cudaMemcpyDeviceToHost, ctx->stream));
      ICK(cudaStreamSynchronize(ctx->stream));
      ctx->stats.d2h_bytes += sizeof(hctl);
      if (hctl[CT_FLAG] & 1) {
        m3b_set_error(ctx, "starting vector near the null space");
        status = M3B_E_NULLSPACE;
        goto done;
      }
      if (hctl[CT_FLAG] & 4) {
        m3b_set_error(ctx, "peer exchange timed out: a rank did not reach Lanczos restart %d", it);
        status = M3B_E_NCCL;
        goto done;
      }
      if (hctl[CT_FLAG] & 8) {
        m3b_set_error(ctx, "internal: grid barrier of the fused vector kernel timed out at restart %d", it);
        status = M3B_E_CUDA;
        goto done;
      }
      if ((hctl[CT_FLAG] & 2) && !careful && ctx->rnorm) {  // (the flag derives from replicated scalars: every rank takes this branch)
        // Invariant subspace (irlb.c: R_F < eps or S < eps).  The cycle only appended columns
        // behind the restart state (V[:, :k+1], W[:, :k], the arrow part of B) and the restart
        // logic left svratio / B / k alone, so the cycle is simply run again, this time looking
        // at every norm on the host and drawing replacement vectors from the caller's stream.
        ICK(cudaStreamSynchronize(ctx->stream2));
        const int cleared = hctl[CT_FLAG] & ~2;
        ICK(cudaMemcpyAsync(I.ctl + CT_FLAG, &cleared, sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
        ICK(cudaStreamSynchronize(ctx->stream));
        mprod = mprod_cycle0;
        pre_k4 = false;
        careful = true;
        continue;
      }
      careful = false;
      if (hctl[CT_FLAG] & 2) {
        m3b_set_error(ctx, "linear dependency encountered (invariant subspace) at restart %d", it);
        status = M3B_E_LINDEP;
        goto done;
      }
      k = hctl[CT_K];
      if (it == 0 || hctl[CT_SVDFLAG] || (ctx->flags & M3B_FLAG_JACOBI_ONLY)) ctx->stats.svd_sweeps += hctl[CT_SWEEPS];
      if (it > 0 && hctl[CT_SVDFLAG]) {
        ctx->stats.svd_fallbacks++;
        ctx->stats.svd_fallback_reasons |= hctl[CT_SVDFLAG];
      }
      if (hctl[CT_CONV]) {
        converged = true;
        it++;
        break;
      }
      if (ctx->should_abort && ctx->should_abort(ctx->should_abort_user)) {
        m3b_set_error(ctx, "aborted by caller");
        status = M3B_E_ABORTED;
        goto done;
      }
      // thick restart: V[:, :k] = V BV[:, :k]; V[:,k] = F; W[:, :k] = W BU[:, :k]
      spmv_timer_begin(ctx, 4);
      IR(I.gemm_inplace(I.V, n, n, work, k, I.BV));
      ICK(cudaMemcpyAsync(I.V + (long long)k * n, I.F, n * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
      IR(I.gemm_inplace(I.W, mloc, mloc, work, k, I.BU));
      spmv_timer_end(ctx);
      it++;
    }
    t_loop1 = now();
    // results: d = BS[:nu]; U*diag(d) = W (BU[:, :nu] diag(d)); V = V BV[:, :nu]
    k_scale_cols<<<(work * nv + 255) / 256, 256, 0, ctx->stream>>>(I.BU, I.BS, work, nv, I.Q);
    count_launch(ctx);
    IR(I.gemm_inplace(I.W, mloc, mloc, work, nv, I.Q));
    IR(I.gemm_inplace(I.V, n, n, work, nv, I.BV));
    ICK(cudaEventRecord(evB, ctx->stream));
    ICK(cudaMemcpyAsync(d_out, I.BS, nv * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    if (V_out) ICK(cudaMemcpyAsync(V_out, I.V, (size_t)n * nv * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    if (copy_x && X_out) {
      ICK(cudaMemcpy2DAsync(X_out, ldx * sizeof(double), I.W, mloc * sizeof(double), mloc * sizeof(double), nv,
                            cudaMemcpyDeviceToHost, ctx->stream));
      ctx->stats.d2h_bytes += mloc * nv * 8;
    }
    if (center_out && use_center)
      ICK(cudaMemcpyAsync(center_out, m->center, n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    if (scale_out && use_scale)
      ICK(cudaMemcpyAsync(scale_out, I.s, n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    {
      unsigned long long hd[XD_COUNT];
      ICK(cudaMemcpyAsync(hd, I.xdiag, sizeof(hd), cudaMemcpyDeviceToHost, ctx->stream));
      ICK(cudaStreamSynchronize(ctx->stream));
      const int khz = std::max(1, ctx->clock_khz);
      for (int q = 0; q < 5; ++q) ctx->stats.xchg_diag[q] = (double)hd[q] / ((double)khz * 1e-3);  // cycles -> us (nominal clock)
      ctx->stats.xchg_diag[5] = (double)hd[XD_N_F];
      ctx->stats.xchg_diag[6] = (double)hd[XD_N_W];
      ctx->stats.xchg_diag[7] = (double)khz;
    }
    spmv_timer_collect(ctx);
    ctx->stats.d2h_bytes += nv * 8 + (V_out ? (long long)n * nv * 8 : 0);
    {
      float ms = 0;
      cudaEventElapsedTime(&ms, evA, evB);
      ctx->stats.irlba_ms = ms;
    }
    // keep the embedding on the device for callers that want it there
    dev_free(m->Xdev);
    m->Xdev = nullptr;
    if (!copy_x) {
      // hand W's first nv columns over without a copy: steal the buffer
      for (auto& p : I.owned)
        if (p == I.W) p = nullptr;
      m->Xdev = I.W;
      m->Xdev_nv = nv;
    }
    t_fin = now();
    if (iter_out) *iter_out = it;
    if (mprod_out) *mprod_out = mprod;
    status = converged ? M3B_OK : M3B_W_NOT_CONVERGED;
    if (!converged) m3b_set_error(ctx, "did not converge--results might be invalid!; try increasing work or maxit");
  }
done:
  cudaStreamSynchronize(ctx->stream2);
  cudaStreamSynchronize(ctx->stream);
  if (evA) cudaEventDestroy(evA);
  if (evB) cudaEventDestroy(evB);
  if (evF) cudaEventDestroy(evF);
  if (evK4) cudaEventDestroy(evK4);
  I.release();
  ctx->stats.host_ms[0] = t_loop0 - t_enter;  // moments + allocation + setup
  ctx->stats.host_ms[1] = t_loop1 - t_loop0;  // Lanczos loop (host wall clock)
  ctx->stats.host_ms[2] = t_fin - t_loop1;    // final GEMMs + copies
  ctx->stats.host_ms[3] = now() - t_fin;      // release
  return status;
#undef IR
#undef ICK
}
