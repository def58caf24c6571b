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

// ---------------------------------------------------------------------------
// K10, tiled: the scaled loadings are staged in SHARED memory, gene tile by gene tile.
//
// The warp-per-cell kernel above pulls one 800-byte row of Vs out of L2 per non-zero: at the cfg5
// shape that is 600 GB of L2 gathers (L2 serves ~12 TB/s chip-wide => >= 50 ms on one GPU) against
// 1.5e11 fp64 flops (~4 ms of DFMA).  Here a CTA owns a block of 128 query cells (16 warps x 8
// cells, accumulators in registers: 4 doubles per lane and cell) and walks the kept genes in tiles
// of ~128 rows of Vs (100 KB), double-buffered in shared memory with one bulk async copy
// (cp.async.bulk + mbarrier; a tile of the row-major Vs is ONE contiguous range) per tile, so the
// per-entry gather becomes two LDS.128 per lane.  The raw CSC is sorted by gene inside a cell, so a
// cell's entries that fall into the current tile are a prefix of what is left of the cell: each
// warp keeps one cursor per cell and fetches 4 entries of each of its 8 cells per round.
// The bound is now the shared-memory pipe (800 B per non-zero at 128 B/clk/SM); the Vs tiles cost
// n_blocks x 24 MB of L2 -> SM traffic, overlapped with the arithmetic.
// Requires: kept index monotone in the gene id (no use_genes re-ordering) and values taken from
// the normalised CSC (not after TF-IDF, which rewrites the layouts only) -- otherwise k_project.
// ---------------------------------------------------------------------------
#define PT_WARPS 16
#define PT_CPW 8                      // cells per warp
#define PT_CELLS (PT_WARPS * PT_CPW)  // cells per CTA
#define PT_SMEM_TILE (110 * 1024)     // bytes per Vs tile buffer (two buffers)

__device__ __forceinline__ void pt_mbar_init(unsigned bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void pt_mbar_expect_tx(unsigned bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void pt_mbar_wait(unsigned bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "PT_WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra PT_WAIT_DONE;\n"
      "bra PT_WAIT_LOOP;\n"
      "PT_WAIT_DONE:\n"
      "}\n" ::"r"(bar),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void pt_bulk_g2s(unsigned dst, const void* src, unsigned bytes, unsigned bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
               "r"(bytes), "r"(bar)
               : "memory");
}

template <bool TMA>
__global__ void __launch_bounds__(PT_WARPS * 32, 1)
k_project_tiled(const long long* __restrict__ p, const int* __restrict__ gi, const double* __restrict__ y,
                const int* __restrict__ kept_of_gene, long long n_cells, int n_kept, const double* __restrict__ Vs, int nvp,
                int ncols, int tg, const double* __restrict__ b, double* __restrict__ Y, long long ldy) {
  extern __shared__ __align__(128) unsigned char pt_smem[];
  double* buf0 = reinterpret_cast<double*>(pt_smem);
  double* buf1 = reinterpret_cast<double*>(pt_smem + PT_SMEM_TILE);
  __shared__ __align__(8) unsigned long long mbar[2];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned bar0 = (unsigned)__cvta_generic_to_shared(&mbar[0]), bar1 = (unsigned)__cvta_generic_to_shared(&mbar[1]);
  if (TMA) {
    if (threadIdx.x == 0) {
      pt_mbar_init(bar0, 1);
      pt_mbar_init(bar1, 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
  }
  unsigned ph0 = 0, ph1 = 0;  // parity of the next completion of each buffer's barrier
  const int n_gt = (n_kept + tg - 1) / tg;
  const size_t row_bytes = (size_t)nvp * sizeof(double);
  const long long n_blocks = (n_cells + PT_CELLS - 1) / PT_CELLS;
  const bool second = 64 + 2 * lane < nvp;  // this lane also owns columns 64 + 2 lane, + 1
  const bool first = 2 * lane < nvp;
  auto load_tile = [&](int gt, int which) {
    const int g0 = gt * tg, rows = min(tg, n_kept - g0);
    const unsigned bytes = (unsigned)(rows * row_bytes);
    double* dst = which ? buf1 : buf0;
    const double* src = Vs + (size_t)g0 * nvp;
    if (TMA) {
      if (threadIdx.x == 0) {
        const unsigned bar = which ? bar1 : bar0;
        pt_mbar_expect_tx(bar, bytes);
        pt_bulk_g2s((unsigned)__cvta_generic_to_shared(dst), src, bytes, bar);
      }
    } else {
      const double2* s2 = reinterpret_cast<const double2*>(src);
      double2* d2 = reinterpret_cast<double2*>(dst);
      for (int k = threadIdx.x; k < (int)(bytes / 16); k += blockDim.x) d2[k] = s2[k];
    }
  };
  for (long long blk = blockIdx.x; blk < n_blocks; blk += gridDim.x) {
    const long long c0 = blk * PT_CELLS + (long long)warp * PT_CPW;
    long long cur = 0, end = 0;  // lanes 0..7: cursor / end of cell c0 + lane
    if (lane < PT_CPW && c0 + lane < n_cells) {
      cur = p[c0 + lane];
      end = p[c0 + lane + 1];
    }
    double acc[PT_CPW][4];
#pragma unroll
    for (int j = 0; j < PT_CPW; ++j) acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.0;
    load_tile(0, 0);
    for (int gt = 0; gt < n_gt; ++gt) {
      const int which = gt & 1;
      if (gt + 1 < n_gt) load_tile(gt + 1, which ^ 1);  // that buffer was released by the barrier below
      if (TMA) {
        if (which) {
          pt_mbar_wait(bar1, ph1);
          ph1 ^= 1u;
        } else {
          pt_mbar_wait(bar0, ph0);
          ph0 ^= 1u;
        }
      } else {
        __syncthreads();
      }
      const double* tile = which ? buf1 : buf0;
      const int g0 = gt * tg, g1 = min(n_kept, g0 + tg);
      const int j_of = lane & 7, sub = lane >> 3;
      for (;;) {
        const long long cj = __shfl_sync(0xffffffffu, cur, j_of), ej = __shfl_sync(0xffffffffu, end, j_of);
        const long long k = cj + sub;
        const bool have = k < ej;
        int kq = -1;
        double val = 0.0;
        if (have) {
          kq = kept_of_gene[gi[k]];
          val = y[k];
        }
        // consumable: a dropped gene (kq < 0) or a kept gene of this or an earlier tile; a cell's
        // consumable entries must form a prefix of the four fetched ones
        const unsigned m = __ballot_sync(0xffffffffu, have && kq < g1);
        const unsigned need = ((1u << (8 * sub + 1)) - 1u) & 0x01010101u;
        const bool cons = (((m >> j_of) & 0x01010101u) & need) == need;
        const unsigned mc = __ballot_sync(0xffffffffu, cons);
        if (mc == 0u) break;
        // (a branch-free variant -- slots without a contribution read row 0 with the value 0 so that the
        // eight chains of a group overlap -- was SLOWER, 89 vs 73 ms at the cfg5 shape: the kernel is
        // bound by the shared-memory pipe, and the dummy slots add LDS traffic)
        const unsigned mv = __ballot_sync(0xffffffffu, cons && kq >= g0);
        const int local = kq - g0;
#pragma unroll
        for (int j = 0; j < PT_CPW; ++j) {
#pragma unroll
          for (int s4 = 0; s4 < 4; ++s4) {
            const int ql = s4 * 8 + j;
            if ((mv >> ql) & 1u) {  // warp-uniform
              const int g = __shfl_sync(0xffffffffu, local, ql);
              const double v = __shfl_sync(0xffffffffu, val, ql);
              const double* row = tile + (size_t)g * nvp;
              if (first) {
                const double2 a = *reinterpret_cast<const double2*>(row + 2 * lane);
                acc[j][0] = fma(v, a.x, acc[j][0]);
                acc[j][1] = fma(v, a.y, acc[j][1]);
              }
              if (second) {
                const double2 c = *reinterpret_cast<const double2*>(row + 64 + 2 * lane);
                acc[j][2] = fma(v, c.x, acc[j][2]);
                acc[j][3] = fma(v, c.y, acc[j][3]);
              }
            }
          }
        }
        if (lane < PT_CPW) cur += __popc((mc >> lane) & 0x01010101u);
      }
      __syncthreads();  // every warp is done with this tile: its buffer may be refilled
    }
    // Y[c, col] = acc - b[col]
#pragma unroll
    for (int j = 0; j < PT_CPW; ++j) {
      const long long c = c0 + j;
      if (c < n_cells) {
        if (first) {
          const int col = 2 * lane;
          if (col < ncols) Y[(long long)col * ldy + c] = acc[j][0] - b[col];
          if (col + 1 < ncols) Y[(long long)(col + 1) * ldy + c] = acc[j][1] - b[col + 1];
        }
        if (second) {
          const int col = 64 + 2 * lane;
          if (col < ncols) Y[(long long)col * ldy + c] = acc[j][2] - b[col];
          if (col + 1 < ncols) Y[(long long)(col + 1) * ldy + c] = acc[j][3] - b[col + 1];
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------
// K10, tiled, second version.  What ncu showed on the first one (profiles/r02e_k_project_tiled*):
// 25 % of the stall samples sit in the block barrier at the end of every gene tile (235 tiles per
// cell block, 16 warps in lockstep, a cell has ~6 entries per tile: whoever has a heavy tile makes
// everybody wait), 17 % wait for the two dependent loads kept_of_gene[gi[k]] that EVERY round
// repeats for its four entries per cell, consumed or not; shared-memory pipe 32 % busy.
//   * a 17th warp is the producer: it alone issues the bulk copies of the Vs tiles and waits for the
//     "empty" mbarrier of a buffer (one arrival per consumer warp) before refilling it; consumer
//     warps wait for the "full" barrier of the tile they need and arrive on "empty" when done with
//     it.  No block barrier: a warp may run one tile ahead of the slowest one;
//   * the four fetched entries of a cell live in registers ACROSS rounds and tiles: after a round
//     that consumed c entries of a cell, the remaining ones move up by c lanes (three shuffles) and
//     only the c vacated slots are loaded -- every entry is loaded exactly once.
// Same arithmetic and summation order per cell as the first version (entries in CSC order).
// ---------------------------------------------------------------------------
// <CPW cells per warp, CW consumer warps> (+ 1 producer warp; CW + 1 a multiple of 4 = the register
// allocation unit): <8, 15> keeps 4 queued entries per cell and 64 accumulator registers, <4, 23>
// 8 queued entries per cell (a tile's ~6 entries of a cell in ONE round) and 32 accumulator
// registers, which lets 24 warps fit.
__device__ __forceinline__ void pt_mbar_arrive(unsigned bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

template <int CPW, int CW>
__global__ void __launch_bounds__((CW + 1) * 32, 1)
k_project_tiled2(const long long* __restrict__ p, const int* __restrict__ gi, const double* __restrict__ y,
                 const int* __restrict__ kept_of_gene, long long n_cells, int n_kept, const double* __restrict__ Vs, int nvp,
                 int ncols, int tg, const double* __restrict__ b, double* __restrict__ Y, long long ldy) {
  extern __shared__ __align__(128) unsigned char pt_smem[];
  double* buf0 = reinterpret_cast<double*>(pt_smem);
  double* buf1 = reinterpret_cast<double*>(pt_smem + PT_SMEM_TILE);
  __shared__ __align__(8) unsigned long long mbar[4];  // full[0], full[1], empty[0], empty[1]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned bar_full = (unsigned)__cvta_generic_to_shared(&mbar[0]);
  const unsigned bar_empty = (unsigned)__cvta_generic_to_shared(&mbar[2]);
  if (threadIdx.x == 0) {
    pt_mbar_init(bar_full, 1);
    pt_mbar_init(bar_full + 8, 1);
    pt_mbar_init(bar_empty, CW);
    pt_mbar_init(bar_empty + 8, CW);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int n_gt = (n_kept + tg - 1) / tg;
  const size_t row_bytes = (size_t)nvp * sizeof(double);
  constexpr int CELLS = CPW * CW;            // cells per CTA
  constexpr int NSUB = 32 / CPW;             // queued entries per cell
  constexpr unsigned CMASK = (CPW == 8) ? 0x01010101u : (CPW == 4 ? 0x11111111u : 0x55555555u);  // one bit per queue slot of a cell
  const long long n_blocks = (n_cells + CELLS - 1) / CELLS;
  if (warp == CW) {
    // ---- producer: one lane streams the tiles of every cell block this CTA owns
    if (lane == 0) {
      long long it = 0;
      for (long long blk = blockIdx.x; blk < n_blocks; blk += gridDim.x)
        for (int gt = 0; gt < n_gt; ++gt, ++it) {
          const int sb = (int)(it & 1);
          if (it >= 2) pt_mbar_wait(bar_empty + 8u * sb, (unsigned)(((it >> 1) - 1) & 1));
          const int g0 = gt * tg, rows = min(tg, n_kept - g0);
          const unsigned bytes = (unsigned)(rows * row_bytes);
          pt_mbar_expect_tx(bar_full + 8u * sb, bytes);
          pt_bulk_g2s((unsigned)__cvta_generic_to_shared(sb ? buf1 : buf0), Vs + (size_t)g0 * nvp, bytes, bar_full + 8u * sb);
        }
    }
    return;
  }
  // ---- consumers
  const bool second = 64 + 2 * lane < nvp;  // this lane also owns columns 64 + 2 lane, + 1
  const bool first = 2 * lane < nvp;
  const int j_of = lane & (CPW - 1), sub = lane / CPW;
  const int NONE = 0x7fffffff;
  long long it = 0;
  for (long long blk = blockIdx.x; blk < n_blocks; blk += gridDim.x) {
    const long long c0 = blk * CELLS + (long long)warp * CPW;
    long long cur = 0, end = 0;  // lanes 0..7: cursor / end of cell c0 + lane
    if (lane < CPW && c0 + lane < n_cells) {
      cur = p[c0 + lane];
      end = p[c0 + lane + 1];
    }
    // this lane's slot of the register queue: entry number `sub` of cell j_of behind its cursor
    int kq = NONE;
    double val = 0.0;
    {
      const long long k = __shfl_sync(0xffffffffu, cur, j_of) + sub, ej = __shfl_sync(0xffffffffu, end, j_of);
      if (k < ej) {
        kq = kept_of_gene[gi[k]];
        val = y[k];
      }
    }
    double acc[CPW][4];
#pragma unroll
    for (int j = 0; j < CPW; ++j) acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.0;
    for (int gt = 0; gt < n_gt; ++gt, ++it) {
      const int sb = (int)(it & 1);
      pt_mbar_wait(bar_full + 8u * sb, (unsigned)((it >> 1) & 1));
      const double* tile = sb ? buf1 : buf0;
      const int g0 = gt * tg, g1 = min(n_kept, g0 + tg);
      for (;;) {
        // consumable: a dropped gene (kq < 0) or a kept gene of this or an earlier tile; a cell's
        // consumable entries must form a prefix of the four queued ones
        const unsigned m = __ballot_sync(0xffffffffu, kq < g1);
        const unsigned need = ((2u << (CPW * sub)) - 1u) & CMASK;
        const bool cons = (((m >> j_of) & CMASK) & need) == need;
        const unsigned mc = __ballot_sync(0xffffffffu, cons);
        if (mc == 0u) break;
        const unsigned mv = __ballot_sync(0xffffffffu, cons && kq >= g0);
        const int local = kq - g0;
#pragma unroll
        for (int j = 0; j < CPW; ++j) {
#pragma unroll
          for (int s4 = 0; s4 < NSUB; ++s4) {
            const int ql = s4 * CPW + j;
            if ((mv >> ql) & 1u) {  // warp-uniform
              const int g = __shfl_sync(0xffffffffu, local, ql);
              const double v = __shfl_sync(0xffffffffu, val, ql);
              const double* row = tile + (size_t)g * nvp;
              if (first) {
                const double2 a = *reinterpret_cast<const double2*>(row + 2 * lane);
                acc[j][0] = fma(v, a.x, acc[j][0]);
                acc[j][1] = fma(v, a.y, acc[j][1]);
              }
              if (second) {
                const double2 c = *reinterpret_cast<const double2*>(row + 64 + 2 * lane);
                acc[j][2] = fma(v, c.x, acc[j][2]);
                acc[j][3] = fma(v, c.y, acc[j][3]);
              }
            }
          }
        }
        // advance the cursors and shift the queue: slot (cell, sub) takes what slot (cell, sub + c)
        // held, c = entries of the cell consumed in this round; vacated slots are loaded
        if (lane < CPW) cur += __popc((mc >> lane) & CMASK);
        const int c = __popc((mc >> j_of) & CMASK);
        const int src = sub + c;
        const int kq_s = __shfl_sync(0xffffffffu, kq, (src & (NSUB - 1)) * CPW + j_of);
        const double val_s = __shfl_sync(0xffffffffu, val, (src & (NSUB - 1)) * CPW + j_of);
        const long long k = __shfl_sync(0xffffffffu, cur, j_of) + sub, ej = __shfl_sync(0xffffffffu, end, j_of);
        if (src < NSUB) {
          kq = kq_s;
          val = val_s;
        } else if (k < ej) {
          kq = kept_of_gene[gi[k]];
          val = y[k];
        } else {
          kq = NONE;
          val = 0.0;
        }
      }
      __syncwarp();
      if (lane == 0) pt_mbar_arrive(bar_empty + 8u * sb);  // this warp is done with the tile
    }
    // Y[c, col] = acc - b[col]
#pragma unroll
    for (int j = 0; j < CPW; ++j) {
      const long long c = c0 + j;
      if (c < n_cells) {
        if (first) {
          const int col = 2 * lane;
          if (col < ncols) Y[(long long)col * ldy + c] = acc[j][0] - b[col];
          if (col + 1 < ncols) Y[(long long)(col + 1) * ldy + c] = acc[j][1] - b[col + 1];
        }
        if (second) {
          const int col = 64 + 2 * lane;
          if (col < ncols) Y[(long long)col * ldy + c] = acc[j][2] - b[col];
          if (col + 1 < ncols) Y[(long long)(col + 1) * ldy + c] = acc[j][3] - b[col + 1];
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------
// Gene selection for the projection path (m3b_select_genes_for_transform): the zero-gene filter of
// R/projection.R:117-118 WITHOUT the two product layouts m3b_select_genes builds for IRLBA -- the
// tiled projection kernel reads the normalised CSC directly.  One pass over the stored entries
// with integer atomics (order-free, exact): per selected gene the number of positive entries, and
// flags for negative / non-finite values.  For non-negative data (every size-factor based
// norm_method on counts) "row sum finite and != 0" is exactly "no non-finite entry and at least
// one positive entry"; anything else falls back to the exact row sums of m3b_select_genes.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_gene_flags(const long long* __restrict__ p, const int* __restrict__ gi,
                                                    const double* __restrict__ y, const int* __restrict__ sel_of_gene,
                                                    long long n_cells, long long* __restrict__ npos, long long* __restrict__ flags) {
  const int lane = threadIdx.x & 31;
  long long warp = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  int bad = 0;
  for (long long c = warp; c < n_cells; c += nwarps) {
    const long long lo = p[c], hi = p[c + 1];
    for (long long k = lo + lane; k < hi; k += 32) {
      const int s = sel_of_gene[gi[k]];
      if (s < 0) continue;
      const double v = y[k];
      if (!isfinite(v)) {
        bad |= 2;
        atomicAdd(reinterpret_cast<unsigned long long*>(flags + 2 + s), 1ULL);
      } else if (v > 0.0) {
        atomicAdd(reinterpret_cast<unsigned long long*>(npos + s), 1ULL);
      } else if (v < 0.0) {
        bad |= 1;
      }
    }
  }
  if (bad & 1) atomicOr(reinterpret_cast<unsigned long long*>(flags), 1ULL);
  if (bad & 2) atomicOr(reinterpret_cast<unsigned long long*>(flags + 1), 1ULL);
}

__global__ void k_compose_kept(const int* __restrict__ sel_of_gene, const int* __restrict__ kept_of_sel, int n_genes,
                               int* __restrict__ kept_of_gene) {
  int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g < n_genes) {
    int s = sel_of_gene[g];
    kept_of_gene[g] = (s >= 0) ? kept_of_sel[s] : -1;
  }
}

// returns M3B_OK with *fallback = 1 when the exact path is needed (negative values, dense pseudo-count path)
int run_select_for_transform(m3b_mat* m, const int* gene_order, int n_sel, unsigned char* keep, int* n_kept, int* fallback) {
  m3b_ctx* ctx = m->ctx;
  *fallback = 0;
  if (m->f0 != 0.0) {
    *fallback = 1;
    return M3B_OK;
  }
  if (!gene_order) n_sel = m->n_genes;
  std::vector<int> sel_of_gene(std::max(m->n_genes, 1), -1);
  bool mono = true;
  for (int s = 0; s < n_sel; ++s) {
    const int g = gene_order ? gene_order[s] : s;
    if (g < 0 || g >= m->n_genes) {
      m3b_set_error(ctx, "use_genes index %d out of range", g);
      return M3B_E_DIMS;
    }
    if (sel_of_gene[g] != -1) {
      m3b_set_error(ctx, "use_genes lists gene %d twice", g);
      return M3B_E_DIMS;
    }
    sel_of_gene[g] = s;
    if (s > 0 && gene_order && gene_order[s] <= gene_order[s - 1]) mono = false;
  }
  long long *d_npos = nullptr, *d_flags = nullptr;
  int status = M3B_OK;
  std::vector<long long> npos(std::max(n_sel, 1), 0), flags(2 + std::max(n_sel, 1), 0);
  std::vector<int> kept_of_sel(std::max(n_sel, 1), -1);
  cudaError_t e = cudaSuccess;
  dev_free(m->sel_of_gene);
  m->sel_of_gene = nullptr;
  RET(dev_alloc(ctx, &m->sel_of_gene, sel_of_gene.size()));
  RET(dev_alloc(ctx, &d_npos, npos.size()));
  if (dev_alloc(ctx, &d_flags, flags.size()) != M3B_OK) {
    dev_free(d_npos);
    return M3B_E_NOMEM;
  }
  do {
    if ((e = cudaMemcpyAsync(m->sel_of_gene, sel_of_gene.data(), sel_of_gene.size() * sizeof(int), cudaMemcpyHostToDevice, ctx->stream)) != cudaSuccess) break;
    if ((e = cudaMemsetAsync(d_npos, 0, npos.size() * sizeof(long long), ctx->stream)) != cudaSuccess) break;
    if ((e = cudaMemsetAsync(d_flags, 0, flags.size() * sizeof(long long), ctx->stream)) != cudaSuccess) break;
    if (m->n_cells && n_sel) {
      const int blocks = (int)std::min<long long>((m->n_cells + 7) / 8, (long long)ctx->sm_count * 8);
      k_gene_flags<<<blocks, 256, 0, ctx->stream>>>(m->p, m->i, m->y, m->sel_of_gene, m->n_cells, d_npos, d_flags);
      count_launch(ctx);
    }
    if ((status = comm_allreduce_i64(ctx, d_npos, npos.size())) != M3B_OK) break;
    if ((status = comm_allreduce_i64(ctx, d_flags, flags.size())) != M3B_OK) break;
    if ((e = cudaMemcpyAsync(npos.data(), d_npos, npos.size() * sizeof(long long), cudaMemcpyDeviceToHost, ctx->stream)) != cudaSuccess) break;
    if ((e = cudaMemcpyAsync(flags.data(), d_flags, flags.size() * sizeof(long long), cudaMemcpyDeviceToHost, ctx->stream)) != cudaSuccess) break;
    if ((e = cudaStreamSynchronize(ctx->stream)) != cudaSuccess) break;
    if (flags[0]) {  // negative values somewhere: the sign of a row sum needs the exact sums
      *fallback = 1;
      break;
    }
    int nk = 0;
    for (int s = 0; s < n_sel; ++s) {
      const bool kp = npos[s] > 0 && flags[2 + s] == 0;  // finite and non-zero row sum (R/projection.R:117-118)
      if (keep) keep[s] = kp ? 1 : 0;
      kept_of_sel[s] = kp ? nk++ : -1;
    }
    m->n_sel = n_sel;
    m->n_kept = nk;
    m->keep_mask.assign(std::max(n_sel, 1), 0);
    for (int s = 0; s < n_sel; ++s) m->keep_mask[s] = kept_of_sel[s] >= 0;
    dev_free(m->kept_of_sel);
    dev_free(m->kept_of_gene);
    m->kept_of_sel = m->kept_of_gene = nullptr;
    if ((status = dev_alloc(ctx, &m->kept_of_sel, kept_of_sel.size())) != M3B_OK) break;
    if ((status = dev_alloc(ctx, &m->kept_of_gene, (size_t)std::max(m->n_genes, 1))) != M3B_OK) break;
    if ((e = cudaMemcpyAsync(m->kept_of_sel, kept_of_sel.data(), kept_of_sel.size() * sizeof(int), cudaMemcpyHostToDevice, ctx->stream)) != cudaSuccess) break;
    if (m->n_genes) {
      k_compose_kept<<<(m->n_genes + 255) / 256, 256, 0, ctx->stream>>>(m->sel_of_gene, m->kept_of_sel, m->n_genes, m->kept_of_gene);
      count_launch(ctx);
    }
    if ((e = cudaStreamSynchronize(ctx->stream)) != cudaSuccess) break;
    // no product layouts: only m3b_project (tiled kernel) may follow
    m->A.free_all();
    m->B.free_all();
    m->layouts_built = false;
    m->sel_monotone = mono;
    m->selected = true;
    m->moments_done = false;
    m->tfidf_applied = false;
    m->split = false;
    if (n_kept) *n_kept = nk;
  } while (0);
  dev_free(d_npos);
  dev_free(d_flags);
  if (e != cudaSuccess) {
    m3b_set_error(ctx, "m3b_select_genes_for_transform: %s", cudaGetErrorString(e));
    return M3B_E_CUDA;
  }
  return status;
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
  // The tiled kernel reads the normalised CSC (sorted by gene inside a cell): it needs the kept index
  // to grow with the gene id and the values of m3b_normalize (TF-IDF rewrites the layouts only).
  const bool tiled = q->sel_monotone && !q->tfidf_applied && (!q->layouts_built || !getenv("M3B_PROJECT_OLD")) && n > 0;
  if (!tiled && !q->layouts_built) {
    m3b_set_error(ctx, "m3b_project: this selection (re-ordered genes) needs the product layouts: call m3b_select_genes, "
                       "not m3b_select_genes_for_transform");
    return M3B_E_STATE;
  }
  // scaled loadings (row-major, rows padded to an even number of columns for the tiled kernel) and
  // the constant row b, built on the host: G'q x nv is tiny
  const int pass_cols = tiled ? 128 : 32 * PJ_CPL;
  const int nvp = tiled ? ((std::min(nv, pass_cols) + 1) & ~1) : nv;
  const int n_pass = (nv + pass_cols - 1) / pass_cols;
  std::vector<double> Vs((size_t)std::max(n, 1) * nvp * (tiled ? n_pass : 1), 0.0), b(nv, 0.0);
  std::vector<long double> bacc(nv, 0.0L);
  for (int g = 0; g < n; ++g) {
    const int r = model_row[g];
    const double s = scale ? scale[r] : 1.0;
    const double ce = (center ? center[r] : 0.0) - q->f0;
    for (int c = 0; c < nv; ++c) {
      const double v = V[(size_t)c * n_model + r];
      if (tiled) Vs[((size_t)(c / pass_cols) * n + g) * nvp + (c % pass_cols)] = v / s;  // one [n][nvp] block per pass
      else Vs[(size_t)g * nv + c] = v / s;
      bacc[c] += (long double)(ce / s) * v;
    }
  }
  for (int c = 0; c < nv; ++c) b[c] = (double)bacc[c];
  double *dVs = nullptr, *db = nullptr, *dY = nullptr;
  DevScope guard;
  guard.own(&dVs);
  guard.own(&db);
  guard.own(&dY);
  RET(dev_alloc(ctx, &dVs, Vs.size()));
  RET(dev_alloc(ctx, &db, (size_t)nv));
  RET(dev_alloc(ctx, &dY, (size_t)std::max<long long>(rows, 1) * nv));
  int status = M3B_OK;
  cudaError_t e;
  do {
    if ((e = cudaMemcpyAsync(dVs, Vs.data(), Vs.size() * sizeof(double), cudaMemcpyHostToDevice, ctx->stream)) != cudaSuccess) break;
    if ((e = cudaMemcpyAsync(db, b.data(), nv * sizeof(double), cudaMemcpyHostToDevice, ctx->stream)) != cudaSuccess) break;
    ctx->stats.h2d_bytes += (long long)Vs.size() * 8 + nv * 8;
    if (rows && tiled) {
      static bool attr = false;
      if (!attr) {
        if ((e = cudaFuncSetAttribute(k_project_tiled<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * PT_SMEM_TILE)) != cudaSuccess) break;
        if ((e = cudaFuncSetAttribute(k_project_tiled<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * PT_SMEM_TILE)) != cudaSuccess) break;
        if ((e = cudaFuncSetAttribute(k_project_tiled2<8, 15>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * PT_SMEM_TILE)) != cudaSuccess) break;
        if ((e = cudaFuncSetAttribute(k_project_tiled2<4, 23>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * PT_SMEM_TILE)) != cudaSuccess) break;
        if ((e = cudaFuncSetAttribute(k_project_tiled2<4, 27>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * PT_SMEM_TILE)) != cudaSuccess) break;
        if ((e = cudaFuncSetAttribute(k_project_tiled2<2, 31>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * PT_SMEM_TILE)) != cudaSuccess) break;
        attr = true;
      }
      const int tg = std::max(8, (int)(PT_SMEM_TILE / (nvp * sizeof(double))));
      const bool tma = !getenv("M3B_PROJECT_NO_TMA");
      const bool v1 = getenv("M3B_PROJECT_V1") != nullptr;  // first tiled version (block barrier per tile), for A/B runs
      const bool wide = getenv("M3B_PROJECT_CPW8") != nullptr;  // 8 cells per warp, 16 warps (A/B runs); default 4 cells, 24 warps
      // default: 4 cells per warp, 27 consumer warps + the producer = 28 warps (cfg5: 52.0 ms; 24 warps 55.6, 32 warps with
      // 2 cells each 55.8, 16 warps with 8 cells each 68.3, first version 72.7); the others stay selectable for A/B runs
      const bool cw27 = getenv("M3B_PROJECT_CW23") == nullptr;
      const bool cpw2 = getenv("M3B_PROJECT_CPW2") != nullptr;  // 2 cells per warp, 32 warps (A/B runs)
      const int cells_per_cta = (tma && !v1) ? (wide ? 8 * 15 : (cpw2 ? 2 * 31 : (cw27 ? 4 * 27 : 4 * 23))) : PT_CELLS;
      const long long n_blocks = (rows + cells_per_cta - 1) / cells_per_cta;
      const int grid = (int)std::min<long long>(n_blocks, ctx->sm_count);
      cudaEventRecord(ctx->evk0, ctx->stream);
      for (int ps = 0; ps < n_pass; ++ps) {
        const int col0 = ps * pass_cols, ncols = std::min(pass_cols, nv - col0);
        const double* vs = dVs + (size_t)ps * n * nvp;
        if (tma && !v1 && wide)
          k_project_tiled2<8, 15><<<grid, 16 * 32, 2 * PT_SMEM_TILE, ctx->stream>>>(q->p, q->i, q->y, q->kept_of_gene, rows, n, vs, nvp,
                                                                                   ncols, tg, db + col0, dY + (size_t)col0 * rows, rows);
        else if (tma && !v1 && cpw2)
          k_project_tiled2<2, 31><<<grid, 32 * 32, 2 * PT_SMEM_TILE, ctx->stream>>>(q->p, q->i, q->y, q->kept_of_gene, rows, n, vs, nvp,
                                                                                   ncols, tg, db + col0, dY + (size_t)col0 * rows, rows);
        else if (tma && !v1 && cw27)
          k_project_tiled2<4, 27><<<grid, 28 * 32, 2 * PT_SMEM_TILE, ctx->stream>>>(q->p, q->i, q->y, q->kept_of_gene, rows, n, vs, nvp,
                                                                                   ncols, tg, db + col0, dY + (size_t)col0 * rows, rows);
        else if (tma && !v1)
          k_project_tiled2<4, 23><<<grid, 24 * 32, 2 * PT_SMEM_TILE, ctx->stream>>>(q->p, q->i, q->y, q->kept_of_gene, rows, n, vs, nvp,
                                                                                   ncols, tg, db + col0, dY + (size_t)col0 * rows, rows);
        else if (tma)
          k_project_tiled<true><<<grid, PT_WARPS * 32, 2 * PT_SMEM_TILE, ctx->stream>>>(q->p, q->i, q->y, q->kept_of_gene, rows, n, vs, nvp,
                                                                                  ncols, tg, db + col0, dY + (size_t)col0 * rows, rows);
        else
          k_project_tiled<false><<<grid, PT_WARPS * 32, 2 * PT_SMEM_TILE, ctx->stream>>>(q->p, q->i, q->y, q->kept_of_gene, rows, n, vs,
                                                                                   nvp, ncols, tg, db + col0, dY + (size_t)col0 * rows, rows);
        count_launch(ctx);
      }
      cudaEventRecord(ctx->evk1, ctx->stream);
      ctx->k10_timed = true;
    } else if (rows) {
      int blocks = (int)std::min<long long>((rows + 7) / 8, (long long)ctx->sm_count * 8);
      cudaEventRecord(ctx->evk0, ctx->stream);
      for (int col0 = 0; col0 < nv; col0 += 32 * PJ_CPL) {
        const TiledLayout* U = q->A.unit;
        k_project<<<blocks, 256, 0, ctx->stream>>>(q->A.idx, q->A.val, q->A.items, q->A.row_item_ptr,
                                                   U ? U->idx : nullptr, U ? U->items : nullptr,
                                                   U ? U->row_item_ptr : nullptr, U ? U->rowscale : nullptr, rows,
                                                   q->A.n_tiles, q->A.tile_w, dVs, db, nv, col0, dY, rows);
        count_launch(ctx);
      }
      cudaEventRecord(ctx->evk1, ctx->stream);
      ctx->k10_timed = true;
    }
    if (rows) {
      if ((e = cudaGetLastError()) != cudaSuccess) break;
      if ((e = cudaMemcpy2DAsync(Y, ldy * sizeof(double), dY, rows * sizeof(double), rows * sizeof(double), nv,
                                 cudaMemcpyDeviceToHost, ctx->stream)) != cudaSuccess)
        break;
      ctx->stats.d2h_bytes += rows * nv * 8;
    }
    e = cudaStreamSynchronize(ctx->stream);
    if (e == cudaSuccess && ctx->k10_timed) {
      float ms = 0.f;
      if (cudaEventElapsedTime(&ms, ctx->evk0, ctx->evk1) == cudaSuccess) ctx->stats.k10_ms = ms;
      ctx->k10_timed = false;
    }
  } while (0);
  if (e != cudaSuccess) {
    m3b_set_error(ctx, "m3b_project: %s", cudaGetErrorString(e));
    status = M3B_E_CUDA;
  }
  return status;
}
