// Per-item stream kernels over the tiled layout: the K3 row statistics (row sums, squared
// deviations, positive counts), and the fallback for the two Lanczos products when a layout has no
// flat plan (flat.cu holds the product kernel proper; M3B_NO_FLAT or > 2^31 groups per layout).
//
// Replaces the DelayedMatrixStats passes of R/utils.R:383,389 (and, as the fallback, CHOLMOD's
// cholmod_sdmult inside irlba, the two products per Lanczos step behind R/utils.R:417).
//
// Shape of the kernel: persistent CTAs (one per SM, 512 threads).  The dense
// operand of a CTA's tile (<= 28672 doubles = 224 KB) is staged in shared memory
// once; each warp then claims CHUNKS of ITEMS (an item = a padded row run of
// <= 2048 entries) and streams them with 64-/128-bit L1-bypassing loads, gathers
// from shared memory, reduces in a fixed shuffle tree and writes ONE partial per
// item.  A second tiny kernel adds the partials of a row in a fixed order, so
// results are bit-reproducible run to run and independent of the CTA schedule.
// k_unit_stream is the same for the index-only unit layouts (count-1 entries).
//
// HBM traffic per launch is the 10 B/entry stream (8 B value + 2 B tile-local
// index; the ALGORITHMIC figure of SURVEY.md section 8d stays 12 B/nnz for an
// int32-indexed CSR) plus 16 B/item descriptors and 8 B/item partials; the
// gathers never leave the SM.
#include <algorithm>

#include "m3b_internal.h"

__device__ __forceinline__ unsigned int ld_stream_u32(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ double2 ld_stream_double2(const double2* p) {
  double2 v;
  asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
  return v;
}
__device__ __forceinline__ int4 ld_stream_int4(const int4* p) {
  int4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0,%1,%2,%3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
               : "l"(p));
  return v;
}

template <int OP>
__device__ __forceinline__ double op_apply(double v, int id, const double* sx, double mu) {
  if (OP == OP_DOT) return v * sx[id];
  if (OP == OP_SUM) return v;
  if (OP == OP_COUNTPOS) return (v > 0.0) ? 1.0 : 0.0;  // Matrix::rowSums(count_matrix > 0)
  double d = v - mu;  // OP_SQDEV (pad entries are corrected by the caller)
  return d * d;
}

// Scheduling.  Every CTA starts on a tile (L.cta_tile, CTAs are split over the tiles by entry
// count), stages that tile's slice of the dense operand once, and then its 32 warps claim
// CHUNKS (short item ranges) of the tile independently with one atomic per chunk on the
// tile's cursor -- no block barrier between chunks, so a slow warp never stalls the other 31
// (the first cut synchronised the block after every ~35 items and lost ~45% of its issue
// slots to barrier stalls; profiles/).  The next claim is issued before the current chunk is
// streamed, so the atomic's latency is hidden.  When a tile runs dry the CTA moves to the
// next tile that still has a worthwhile amount of work (one barrier + one re-stage).
#define TS_THREADS 512  // 16 warps x <=128 registers: room for two register-resident load batches
#define TS_BP 6         // pairs (of entries) per lane per batch

struct Batch {
  unsigned int i[TS_BP];  // two 16-bit tile-local indices
  double2 v[TS_BP];
};

// one batch = 32 lanes x TS_BP pairs = 256 entries (3 KB) of the item's stream, predicated at the tail
__device__ __forceinline__ void batch_load(Batch& b, const unsigned int* __restrict__ ip,
                                           const double2* __restrict__ vp, int q0, int n2) {
#pragma unroll
  for (int u = 0; u < TS_BP; ++u) {
    const int q = q0 + 32 * u;
    if (q < n2) {
      b.i[u] = ld_stream_u32(ip + q);
      b.v[u] = ld_stream_double2(vp + q);
    } else {
      b.i[u] = 0u;
      b.v[u] = make_double2(0.0, 0.0);
    }
  }
}

template <int OP>
__device__ __forceinline__ void batch_apply(const Batch& b, const double* sx, double mu, double (&a)[TS_BP], int q0,
                                            int n2) {
#pragma unroll
  for (int u = 0; u < TS_BP; ++u) {
    if (OP == OP_SQDEV) {
      // predicated-off lanes hold zeros, which OP_SQDEV would count as (0-mu)^2
      if (q0 + 32 * u < n2) {
        a[u] += op_apply<OP>(b.v[u].x, (int)(b.i[u] & 0x7fffu), sx, mu);
        a[u] += op_apply<OP>(b.v[u].y, (int)(b.i[u] >> 16), sx, mu);
      }
    } else {
      a[u] += op_apply<OP>(b.v[u].x, (int)(b.i[u] & 0x7fffu), sx, mu);
      a[u] += op_apply<OP>(b.v[u].y, (int)(b.i[u] >> 16), sx, mu);
    }
  }
}

// Scheduling.  Every CTA starts on a tile (L.cta_tile, CTAs are split over the tiles by entry
// count), stages that tile's slice of the dense operand once, and then its warps claim
// CHUNKS (short item ranges) of the tile independently with one atomic per chunk on the
// tile's cursor -- no block barrier between chunks, so a slow warp never stalls the others
// (the first cut synchronised the block after every ~35 items and lost ~45% of its issue
// slots to barrier stalls; profiles/).  The next claim is issued before the current chunk is
// streamed, so the atomic's latency is hidden.  Inside an item the stream is software-
// pipelined over two register-resident batches: the loads of batch k+1 are in flight while
// batch k is gathered and accumulated (the kernel is latency-bound, not issue-bound).
// When a tile runs dry the CTA moves to the next tile that still has a worthwhile amount of
// work (one barrier + one re-stage).
template <int OP>
__global__ void __launch_bounds__(TS_THREADS, 1)
k_tile_stream(const unsigned short* __restrict__ idx, const double* __restrict__ val, const ItemDesc* __restrict__ items,
              const int* __restrict__ item_slot, const int2* __restrict__ chunks,
              const int* __restrict__ tile_chunk_ptr, const int* __restrict__ cta_tile, int n_tiles, int tile_w,
              long long cols, const double* __restrict__ x, double* __restrict__ partial,
              unsigned int* __restrict__ counters) {
  extern __shared__ double sx[];
  __shared__ int s_tile;
  const int lane = threadIdx.x & 31;
  int tile = cta_tile[blockIdx.x];
  for (int visited = 0; visited < n_tiles && tile >= 0; ++visited) {
    const int cbeg = tile_chunk_ptr[tile];
    const int nch = tile_chunk_ptr[tile + 1] - cbeg;
    unsigned int* cursor = counters + tile;
    // claim the first chunk before staging: the atomic overlaps the fill
    unsigned int c_next = 0;
    if (lane == 0) c_next = atomicAdd(cursor, 1u);
    if (OP == OP_DOT) {
      // stage the tile of the dense operand (coalesced 16-byte loads; x is L2-resident)
      const long long off = (long long)tile * tile_w;
      const int len = (int)min((long long)tile_w, cols - off);
      const double* xs0 = x + off;
      if ((reinterpret_cast<unsigned long long>(xs0) & 15ull) == 0ull) {
        const double2* src = reinterpret_cast<const double2*>(xs0);
        double2* dst = reinterpret_cast<double2*>(sx);
        const int n2 = len >> 1;
        for (int k = threadIdx.x; k < n2; k += blockDim.x) dst[k] = src[k];
        if ((len & 1) && threadIdx.x == 0) sx[len - 1] = xs0[len - 1];
      } else {  // odd leading dimension (e.g. an odd number of local cells): 8-byte loads
        for (int k = threadIdx.x; k < len; k += blockDim.x) sx[k] = xs0[k];
      }
      __syncthreads();
    }
    for (;;) {
      const int c = (int)__shfl_sync(0xffffffffu, c_next, 0);
      if (c >= nch) break;
      if (lane == 0) c_next = atomicAdd(cursor, 1u);  // prefetch the next claim
      const int2 ch = chunks[cbeg + c];
      for (int ib = ch.x; ib < ch.y; ib += 32) {
        // the descriptors of up to 32 items in one coalesced load
        const int nit = min(32, ch.y - ib);
        int4 mydesc = make_int4(0, 0, 0, 0);
        int myslot = 0;
        if (lane < nit) {
          mydesc = ld_stream_int4(reinterpret_cast<const int4*>(items + ib + lane));
          myslot = item_slot[ib + lane];
        }
        for (int k = 0; k < nit; ++k) {
          const int rx = __shfl_sync(0xffffffffu, mydesc.x, k), ry = __shfl_sync(0xffffffffu, mydesc.y, k);
          const int len = __shfl_sync(0xffffffffu, mydesc.z, k);
          const int row = __shfl_sync(0xffffffffu, mydesc.w, k);
          const long long start = ((long long)(unsigned)rx) | ((long long)ry << 32);
          const double mu = (OP == OP_SQDEV) ? x[row] : 0.0;
          const unsigned int* ip = reinterpret_cast<const unsigned int*>(idx + start);
          const double2* vp = reinterpret_cast<const double2*>(val + start);
          const int n2 = len >> 1;
          double a[TS_BP];
#pragma unroll
          for (int u = 0; u < TS_BP; ++u) a[u] = 0.0;
          Batch b0, b1;
          const int step = 32 * TS_BP;
          int q = lane;
          batch_load(b0, ip, vp, q, n2);
          for (; q < n2 + lane; q += 2 * step) {  // q - lane is warp-uniform
            const bool more1 = (q - lane + step) < n2;
            if (more1) batch_load(b1, ip, vp, q + step, n2);
            batch_apply<OP>(b0, sx, mu, a, q, n2);
            if (!more1) break;
            const bool more2 = (q - lane + 2 * step) < n2;
            if (more2) batch_load(b0, ip, vp, q + 2 * step, n2);
            batch_apply<OP>(b1, sx, mu, a, q + step, n2);
            if (!more2) break;
          }
          double acc = 0.0;
#pragma unroll
          for (int u = 0; u < TS_BP; ++u) acc += a[u];
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
          const int slot = __shfl_sync(0xffffffffu, myslot, k);
          if (lane == 0) partial[slot] = acc;
        }
      }
    }
    if (n_tiles == 1) break;
    // this tile is dry: look for another one that still has a worthwhile amount of work
    __syncthreads();
    if (threadIdx.x == 0) {
      int found = -1, t2 = tile;
      for (int k = 1; k < n_tiles; ++k) {
        t2 = (t2 + 1 == n_tiles) ? 0 : t2 + 1;
        const int n2c = tile_chunk_ptr[t2 + 1] - tile_chunk_ptr[t2];
        const unsigned int used = *((volatile unsigned int*)(counters + t2));
        if ((long long)n2c - (long long)used >= 32) {
          found = t2;
          break;
        }
      }
      s_tile = found;
    }
    __syncthreads();
    tile = s_tile;
  }
  // the last CTA to leave re-arms the cursors for the next launch
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    unsigned int done = atomicInc(&counters[n_tiles], gridDim.x - 1);
    if (done == gridDim.x - 1)
      for (int t = 0; t < n_tiles; ++t) counters[t] = 0u;
  }
}

// ---------------------------------------------------------------------------
// The UNIT stream (count-1 entries; m3b_internal.h "UNIT ENTRIES"): same scheduling, but an
// entry is just a 16-bit tile-local index -- 2 B instead of 10 B of HBM traffic -- and its value
// is applied on the dense side: the staged tile holds x[c] * colscale[c] and the item's sum is
// multiplied by rowscale[row].  A lane loads 8 indices with one 16-byte load; runs are padded
// to 8 with idx == tile_w, a slot behind the tile that holds 0.
//   OP_DOT   partial = rowscale * sum (x * colscale)[idx]
//   OP_SUM   the same with x == 1 (row sums of the values)
//   OP_SQDEV partial = sum (colscale[idx] - mu_row)^2      (pads skipped; rowscale unsupported)
// ---------------------------------------------------------------------------
#define TU_BV 2  // 16-byte loads per lane per batch => 512 entries per warp batch

__device__ __forceinline__ uint4 ld_stream_uint4(const uint4* p) {
  uint4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
               : "l"(p));
  return v;
}

struct UBatch {
  uint4 i[TU_BV];
};

__device__ __forceinline__ void ubatch_load(UBatch& b, const uint4* __restrict__ ip, int q0, int n8, unsigned int padpair) {
#pragma unroll
  for (int u = 0; u < TU_BV; ++u) {
    const int q = q0 + 32 * u;
    if (q < n8) b.i[u] = ld_stream_uint4(ip + q);
    else b.i[u] = make_uint4(padpair, padpair, padpair, padpair);
  }
}

template <int OP>
__device__ __forceinline__ double uop(unsigned int id, const double* sx, double mu, int tile_w) {
  if (OP == OP_SQDEV) {
    const double d = sx[id] - mu;
    return ((int)id < tile_w) ? d * d : 0.0;
  }
  return sx[id];
}

template <int OP>
__device__ __forceinline__ void ubatch_apply(const UBatch& b, const double* sx, double mu, int tile_w, double (&a)[4]) {
#pragma unroll
  for (int u = 0; u < TU_BV; ++u) {
    a[0] += uop<OP>(b.i[u].x & 0x7fffu, sx, mu, tile_w);
    a[1] += uop<OP>(b.i[u].x >> 16, sx, mu, tile_w);
    a[2] += uop<OP>(b.i[u].y & 0x7fffu, sx, mu, tile_w);
    a[3] += uop<OP>(b.i[u].y >> 16, sx, mu, tile_w);
    a[0] += uop<OP>(b.i[u].z & 0x7fffu, sx, mu, tile_w);
    a[1] += uop<OP>(b.i[u].z >> 16, sx, mu, tile_w);
    a[2] += uop<OP>(b.i[u].w & 0x7fffu, sx, mu, tile_w);
    a[3] += uop<OP>(b.i[u].w >> 16, sx, mu, tile_w);
  }
}

template <int OP>
__global__ void __launch_bounds__(TS_THREADS, 1)
k_unit_stream(const unsigned short* __restrict__ idx, const ItemDesc* __restrict__ items,
              const int* __restrict__ item_slot, const int2* __restrict__ chunks,
              const int* __restrict__ tile_chunk_ptr, const int* __restrict__ cta_tile, int n_tiles, int tile_w,
              long long cols, const double* __restrict__ x, const double* __restrict__ colscale,
              const double* __restrict__ rowscale, double* __restrict__ partial, unsigned int* __restrict__ counters) {
  extern __shared__ double sx[];
  __shared__ int s_tile;
  const int lane = threadIdx.x & 31;
  const unsigned int padpair = (unsigned)tile_w | ((unsigned)tile_w << 16);
  int tile = cta_tile[blockIdx.x];
  for (int visited = 0; visited < n_tiles && tile >= 0; ++visited) {
    const int cbeg = tile_chunk_ptr[tile];
    const int nch = tile_chunk_ptr[tile + 1] - cbeg;
    unsigned int* cursor = counters + tile;
    unsigned int c_next = 0;
    if (lane == 0) c_next = atomicAdd(cursor, 1u);
    {
      // stage x * colscale for the tile's columns; slot tile_w (the pads) and the tail hold 0
      const long long off = (long long)tile * tile_w;
      const int len = (int)min((long long)tile_w, cols - off);
      for (int k = threadIdx.x; k <= tile_w; k += blockDim.x) {
        double v = 0.0;
        if (k < len) {
          v = (OP == OP_DOT && x) ? x[off + k] : 1.0;
          if (colscale) v *= colscale[off + k];
        }
        sx[k] = v;
      }
      __syncthreads();
    }
    for (;;) {
      const int c = (int)__shfl_sync(0xffffffffu, c_next, 0);
      if (c >= nch) break;
      if (lane == 0) c_next = atomicAdd(cursor, 1u);  // prefetch the next claim
      const int2 ch = chunks[cbeg + c];
      for (int ib = ch.x; ib < ch.y; ib += 32) {
        const int nit = min(32, ch.y - ib);
        int4 mydesc = make_int4(0, 0, 0, 0);
        int myslot = 0;
        double myrs = 1.0;
        if (lane < nit) {
          mydesc = ld_stream_int4(reinterpret_cast<const int4*>(items + ib + lane));
          myslot = item_slot[ib + lane];
          if (rowscale) myrs = rowscale[mydesc.w];
        }
        for (int k = 0; k < nit; ++k) {
          const int rx = __shfl_sync(0xffffffffu, mydesc.x, k), ry = __shfl_sync(0xffffffffu, mydesc.y, k);
          const int len = __shfl_sync(0xffffffffu, mydesc.z, k);
          const int row = __shfl_sync(0xffffffffu, mydesc.w, k);
          const long long start = ((long long)(unsigned)rx) | ((long long)ry << 32);
          const double mu = (OP == OP_SQDEV) ? x[row] : 0.0;
          const uint4* ip = reinterpret_cast<const uint4*>(idx + start);
          const int n8 = len >> 3;
          double a[4] = {0.0, 0.0, 0.0, 0.0};
          UBatch b0, b1;
          const int step = 32 * TU_BV;
          int q = lane;
          ubatch_load(b0, ip, q, n8, padpair);
          for (; q < n8 + lane; q += 2 * step) {  // q - lane is warp-uniform
            const bool more1 = (q - lane + step) < n8;
            if (more1) ubatch_load(b1, ip, q + step, n8, padpair);
            ubatch_apply<OP>(b0, sx, mu, tile_w, a);
            if (!more1) break;
            const bool more2 = (q - lane + 2 * step) < n8;
            if (more2) ubatch_load(b0, ip, q + 2 * step, n8, padpair);
            ubatch_apply<OP>(b1, sx, mu, tile_w, a);
            if (!more2) break;
          }
          double acc = (a[0] + a[1]) + (a[2] + a[3]);
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
          const int slot = __shfl_sync(0xffffffffu, myslot, k);
          const double rs = __shfl_sync(0xffffffffu, myrs, k);
          if (lane == 0) partial[slot] = rs * acc;
        }
      }
    }
    if (n_tiles == 1) break;
    __syncthreads();
    if (threadIdx.x == 0) {
      int found = -1, t2 = tile;
      for (int k = 1; k < n_tiles; ++k) {
        t2 = (t2 + 1 == n_tiles) ? 0 : t2 + 1;
        const int n2c = tile_chunk_ptr[t2 + 1] - tile_chunk_ptr[t2];
        const unsigned int used = *((volatile unsigned int*)(counters + t2));
        if ((long long)n2c - (long long)used >= 32) {
          found = t2;
          break;
        }
      }
      s_tile = found;
    }
    __syncthreads();
    tile = s_tile;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    unsigned int done = atomicInc(&counters[n_tiles], gridDim.x - 1);
    if (done == gridDim.x - 1)
      for (int t = 0; t < n_tiles; ++t) counters[t] = 0u;
  }
}

static size_t tile_smem_bytes(const TiledLayout& L, int op) {
  if (L.is_unit) return ((size_t)L.tile_w + 2) * sizeof(double);
  if (op != OP_DOT) return 16;
  return (size_t)L.tile_w * sizeof(double);
}

int spmv_init(m3b_ctx* ctx) {
  const int bytes = (int)((M3B_TILE_W_MAX + 2) * sizeof(double));
  CK(ctx, cudaFuncSetAttribute(k_tile_stream<OP_DOT>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  CK(ctx, cudaFuncSetAttribute(k_tile_stream<OP_SUM>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  CK(ctx, cudaFuncSetAttribute(k_tile_stream<OP_SQDEV>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  CK(ctx, cudaFuncSetAttribute(k_tile_stream<OP_COUNTPOS>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  CK(ctx, cudaFuncSetAttribute(k_unit_stream<OP_DOT>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  CK(ctx, cudaFuncSetAttribute(k_unit_stream<OP_SUM>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  CK(ctx, cudaFuncSetAttribute(k_unit_stream<OP_SQDEV>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  RET(flat_init(ctx));
  return M3B_OK;
}

static int launch_unit_stream(m3b_ctx* ctx, const TiledLayout& L, int op, const double* x, cudaStream_t stream, int grid) {
  if (L.n_chunks == 0) return M3B_OK;
  if (op == OP_COUNTPOS || (op == OP_SQDEV && L.rowscale)) {
    m3b_set_error(ctx, "internal: operation %d is not defined on this unit layout", op);
    return M3B_E_STATE;
  }
  const size_t smem = tile_smem_bytes(L, op);
#define M3B_ULAUNCH(OPX)                                                                                              \
  k_unit_stream<OPX><<<grid, TS_THREADS, smem, stream>>>(L.idx, L.items, L.item_slot, L.chunks, L.tile_chunk_ptr,    \
                                                         L.cta_tile, L.n_tiles, L.tile_w, L.cols, x, L.colscale,     \
                                                         L.rowscale, L.partial, L.counters)
  switch (op) {
    case OP_DOT: M3B_ULAUNCH(OP_DOT); break;
    case OP_SUM: M3B_ULAUNCH(OP_SUM); break;
    default: M3B_ULAUNCH(OP_SQDEV); break;
  }
#undef M3B_ULAUNCH
  count_launch(ctx);
  CK(ctx, cudaGetLastError());
  return M3B_OK;
}

int launch_tile_stream(m3b_ctx* ctx, const TiledLayout& L, int op, const double* x, cudaStream_t stream, int max_cta) {
  if (!stream) stream = ctx->stream;
  if (op == OP_DOT && L.plan[0]) return flat_launch_dot(ctx, L, x, stream, max_cta);  // the product proper (flat.cu)
  // fewer CTAs than SMs is allowed for a single-tile layout (chunks are claimed dynamically)
  const int grid = (max_cta > 0 && L.n_tiles == 1) ? std::min(L.n_cta, max_cta) : L.n_cta;
  if (L.unit) RET(launch_unit_stream(ctx, *L.unit, op, x, stream, grid));
  if (L.n_chunks == 0) return M3B_OK;
  const size_t smem = tile_smem_bytes(L, op);
  switch (op) {
    case OP_DOT:
      k_tile_stream<OP_DOT><<<grid, TS_THREADS, smem, stream>>>(L.idx, L.val, L.items, L.item_slot, L.chunks, L.tile_chunk_ptr,
                                                               L.cta_tile, L.n_tiles, L.tile_w, L.cols, x, L.partial,
                                                               L.counters);
      break;
    case OP_SUM:
      k_tile_stream<OP_SUM><<<grid, TS_THREADS, smem, stream>>>(L.idx, L.val, L.items, L.item_slot, L.chunks, L.tile_chunk_ptr,
                                                               L.cta_tile, L.n_tiles, L.tile_w, L.cols, x, L.partial,
                                                               L.counters);
      break;
    case OP_COUNTPOS:
      k_tile_stream<OP_COUNTPOS><<<grid, TS_THREADS, smem, stream>>>(L.idx, L.val, L.items, L.item_slot, L.chunks, L.tile_chunk_ptr,
                                                                    L.cta_tile, L.n_tiles, L.tile_w, L.cols, x, L.partial,
                                                                    L.counters);
      break;
    default:
      k_tile_stream<OP_SQDEV><<<grid, TS_THREADS, smem, stream>>>(L.idx, L.val, L.items, L.item_slot, L.chunks, L.tile_chunk_ptr,
                                                                 L.cta_tile, L.n_tiles, L.tile_w, L.cols, x, L.partial,
                                                                 L.counters);
      break;
  }
  count_launch(ctx);
  CK(ctx, cudaGetLastError());
  return M3B_OK;
}

// out[row] = sum of the row's partials (contiguous slots, tile-then-segment order): 8 lanes per
// row stride over the slots, fixed 3-step shuffle tree => deterministic
__global__ void __launch_bounds__(256) k_combine_rows(const double* __restrict__ partial, const int* __restrict__ rsp,
                                                      long long rows, double* __restrict__ out) {
  const int sub = threadIdx.x & 7;
  const long long rpb = blockDim.x >> 3;
  for (long long r0 = blockIdx.x * rpb; r0 < rows; r0 += gridDim.x * rpb) {
    const long long r = r0 + (threadIdx.x >> 3);
    double a = 0.0;
    if (r < rows) {
      const int b = rsp[r], e = rsp[r + 1];
      for (int k = b + sub; k < e; k += 8) a += partial[k];
    }
    a += __shfl_xor_sync(0xffffffffu, a, 4);
    a += __shfl_xor_sync(0xffffffffu, a, 2);
    a += __shfl_xor_sync(0xffffffffu, a, 1);
    if (r < rows && sub == 0) out[r] = a;
  }
}

int launch_combine_rows(m3b_ctx* ctx, const TiledLayout& L, double* out) {
  if (L.rows == 0) return M3B_OK;
  const unsigned grid = (unsigned)std::max<long long>(1, std::min<long long>((L.rows + 31) / 32, 8LL * ctx->sm_count));
  k_combine_rows<<<grid, 256, 0, ctx->stream>>>(L.partial, L.row_slot_ptr, L.rows, out);
  count_launch(ctx);
  CK(ctx, cudaGetLastError());
  return M3B_OK;
}

// B_spmv = 12*nnz + 8*(rows + cols) + 8*(rows+1)
long long spmv_algorithmic_bytes(const TiledLayout& L) {
  const long long nnz = L.nnz + (L.unit ? L.unit->nnz : 0);
  return 12LL * nnz + 8LL * (L.rows + L.cols) + 8LL * (L.rows + 1);
}

// EXACT DRAM bytes of one product launch over this layout (the physical counterpart of the
// figure above): the two entry streams as stored (pads included), the plan descriptors, the
// per-item slot / scale words, the dense operand (and the column scale of layout B) once --
// re-reads by the other CTAs of a tile are served by L2 -- and one 8-byte partial per item.
long long flat_plan_pieces(const struct FlatPlan* p);
long long spmv_stream_bytes(const TiledLayout& L) {
  const TiledLayout* U = L.unit;
  long long b = 10LL * L.n_entries + 4LL * L.n_items + 8LL * L.n_items;
  if (U) {
    b += 2LL * U->n_entries + 4LL * U->n_items + 8LL * U->n_items;
    if (U->item_scale) b += 8LL * U->n_items;
    if (U->colscale) b += 8LL * L.cols;
  }
  b += 8LL * L.cols;
  if (L.plan[0]) b += 16LL * flat_plan_pieces(L.plan[0]);
  return b;
}
