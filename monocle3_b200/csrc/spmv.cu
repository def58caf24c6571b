// K4 / K5 (and the K3 row statistics): one kernel family over the tiled layout.
//
// Replaces CHOLMOD's cholmod_sdmult inside irlba (the two products per Lanczos
// step behind R/utils.R:417) and the DelayedMatrixStats passes of R/utils.R:383,389.
//
// Shape of the kernel: persistent CTAs (one per SM, 1024 threads) pull WORK
// UNITS (a contiguous item range of one tile) from a global counter.  The
// dense operand of the unit's tile (<= 28672 doubles = 224 KB) is staged in
// shared memory once per tile change; each warp then streams one ITEM (a padded
// row run, <= 2048 entries) with 64-/128-bit L1-bypassing loads, gathers from
// shared memory, reduces in a fixed shuffle tree and writes ONE partial per
// item.  A second tiny kernel adds the partials of a row in a fixed order, so
// results are bit-reproducible run to run and independent of the CTA schedule.
//
// HBM traffic per launch is the 12 B/entry stream (8 B value + 4 B index) plus
// 16 B/item descriptors and 8 B/item partials; the gathers never leave the SM.
#include <algorithm>

#include "m3b_internal.h"

__device__ __forceinline__ int2 ld_stream_int2(const int2* p) {
  int2 v;
  asm volatile("ld.global.nc.L1::no_allocate.v2.s32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p));
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
  double d = v - mu;  // OP_SQDEV (pad entries are corrected by the caller)
  return d * d;
}

template <int OP>
__global__ void __launch_bounds__(1024, 1)
k_tile_stream(const int* __restrict__ idx, const double* __restrict__ val, const ItemDesc* __restrict__ items,
              const int4* __restrict__ units, int n_units, int tile_w, long long cols,
              const double* __restrict__ x, double* __restrict__ partial, unsigned int* __restrict__ sched) {
  extern __shared__ double sx[];
  __shared__ int s_unit;
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int nwarps = blockDim.x >> 5;
  int cur_tile = -1;
  for (;;) {
    if (threadIdx.x == 0) s_unit = (int)atomicAdd(&sched[0], 1u);
    __syncthreads();
    const int u = s_unit;
    __syncthreads();
    if (u >= n_units) break;
    const int4 un = units[u];
    if (OP == OP_DOT && un.x != cur_tile) {
      // stage the tile of the dense operand (coalesced 16-byte loads; x is L2-resident)
      const long long off = (long long)un.x * tile_w;
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
      cur_tile = un.x;
      __syncthreads();
    }
    for (int it = un.y + warp; it < un.z; it += nwarps) {
      const int4 raw = ld_stream_int4(reinterpret_cast<const int4*>(items + it));
      const long long start = ((long long)(unsigned)raw.x) | ((long long)raw.y << 32);
      const int len = raw.z;
      const double mu = (OP == OP_SQDEV) ? x[raw.w] : 0.0;
      const int2* ip = reinterpret_cast<const int2*>(idx + start);
      const double2* vp = reinterpret_cast<const double2*>(val + start);
      const int n2 = len >> 1;
      double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
      int q = lane;
      for (; q + 96 < n2; q += 128) {  // 4 independent pairs in flight per lane
        const int2 i0 = ld_stream_int2(ip + q), i1 = ld_stream_int2(ip + q + 32);
        const int2 i2 = ld_stream_int2(ip + q + 64), i3 = ld_stream_int2(ip + q + 96);
        const double2 v0 = ld_stream_double2(vp + q), v1 = ld_stream_double2(vp + q + 32);
        const double2 v2 = ld_stream_double2(vp + q + 64), v3 = ld_stream_double2(vp + q + 96);
        a0 += op_apply<OP>(v0.x, i0.x, sx, mu);
        a0 += op_apply<OP>(v0.y, i0.y, sx, mu);
        a1 += op_apply<OP>(v1.x, i1.x, sx, mu);
        a1 += op_apply<OP>(v1.y, i1.y, sx, mu);
        a2 += op_apply<OP>(v2.x, i2.x, sx, mu);
        a2 += op_apply<OP>(v2.y, i2.y, sx, mu);
        a3 += op_apply<OP>(v3.x, i3.x, sx, mu);
        a3 += op_apply<OP>(v3.y, i3.y, sx, mu);
      }
      for (; q < n2; q += 32) {
        const int2 i0 = ld_stream_int2(ip + q);
        const double2 v0 = ld_stream_double2(vp + q);
        a0 += op_apply<OP>(v0.x, i0.x, sx, mu);
        a0 += op_apply<OP>(v0.y, i0.y, sx, mu);
      }
      double acc = (a0 + a1) + (a2 + a3);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
      if (lane == 0) partial[it] = acc;
    }
  }
  // the last CTA to leave re-arms the scheduler for the next launch
  if (threadIdx.x == 0) {
    __threadfence();
    unsigned int done = atomicInc(&sched[1], gridDim.x - 1);
    if (done == gridDim.x - 1) sched[0] = 0u;
  }
}

static size_t tile_smem_bytes(const TiledLayout& L, int op) {
  if (op != OP_DOT) return 16;
  return (size_t)L.tile_w * sizeof(double);
}

int launch_tile_stream(m3b_ctx* ctx, const TiledLayout& L, int op, const double* x) {
  if (L.n_units == 0) return M3B_OK;
  const size_t smem = tile_smem_bytes(L, op);
  const int grid = std::min(ctx->sm_count, L.n_units);
  static bool attr_set[3] = {false, false, false};
  if (!attr_set[op]) {
    const void* fn = (op == OP_DOT)   ? (const void*)k_tile_stream<OP_DOT>
                     : (op == OP_SUM) ? (const void*)k_tile_stream<OP_SUM>
                                      : (const void*)k_tile_stream<OP_SQDEV>;
    CK(ctx, cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(M3B_TILE_W_MAX * sizeof(double))));
    attr_set[op] = true;
  }
  switch (op) {
    case OP_DOT:
      k_tile_stream<OP_DOT><<<grid, 1024, smem, ctx->stream>>>(L.idx, L.val, L.items, L.units, L.n_units, L.tile_w,
                                                               L.cols, x, L.partial, ctx->sched);
      break;
    case OP_SUM:
      k_tile_stream<OP_SUM><<<grid, 1024, smem, ctx->stream>>>(L.idx, L.val, L.items, L.units, L.n_units, L.tile_w,
                                                               L.cols, x, L.partial, ctx->sched);
      break;
    default:
      k_tile_stream<OP_SQDEV><<<grid, 1024, smem, ctx->stream>>>(L.idx, L.val, L.items, L.units, L.n_units, L.tile_w,
                                                                 L.cols, x, L.partial, ctx->sched);
      break;
  }
  count_launch(ctx);
  CK(ctx, cudaGetLastError());
  return M3B_OK;
}

// out[row] = sum over tiles (ascending) and the row's items (ascending) of partial[]
__global__ void k_combine_rows(const double* __restrict__ partial, const int* __restrict__ rip, long long rows,
                               int n_tiles, double* __restrict__ out) {
  long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (r >= rows) return;
  double acc = 0.0;
  for (int t = 0; t < n_tiles; ++t) {
    const int* p = rip + (long long)t * (rows + 1) + r;
    const int b = p[0], e = p[1];
    for (int k = b; k < e; ++k) acc += partial[k];
  }
  out[r] = acc;
}

int launch_combine_rows(m3b_ctx* ctx, const TiledLayout& L, double* out) {
  if (L.rows == 0) return M3B_OK;
  k_combine_rows<<<(unsigned)((L.rows + 255) / 256), 256, 0, ctx->stream>>>(L.partial, L.row_item_ptr, L.rows, L.n_tiles, out);
  count_launch(ctx);
  CK(ctx, cudaGetLastError());
  return M3B_OK;
}

// B_spmv = 12*nnz + 8*(rows + cols) + 8*(rows+1)
long long spmv_algorithmic_bytes(const TiledLayout& L) {
  return 12LL * L.nnz + 8LL * (L.rows + L.cols) + 8LL * (L.rows + 1);
}
