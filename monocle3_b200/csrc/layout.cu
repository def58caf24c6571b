// K0 (ingest + dual tiled layout build), K1 (cell totals), K2 (fused normalise+log).
//
// Replaces, for dgCMatrix input: round()+Matrix::colSums (R/utils.R:57-61),
// t(t(FM)/sf) + log2(FM@x+1) (R/preprocess_cds.R:272-283), FM[use_genes,],
// Matrix::rowSums filter (R/preprocess_cds.R:117-122) and the Matrix::t()
// materialisation handed to irlba (R/preprocess_cds.R:139).
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <cstring>

#include "m3b_internal.h"

// ---------------------------------------------------------------------------
// Device memory: a small caching allocator in front of cudaMalloc/cudaFree.
// A preprocess_cds call allocates and releases ~40 buffers (several GB); going to
// the driver for each one costs milliseconds and a device synchronisation per
// call, which showed up as 50-150 ms of jitter per step.  Freed blocks are kept
// in a size-ordered free list (per device) and handed out again when a request
// fits within 12.5% slack.  Why reuse needs no extra synchronisation: the pool is per
// device and process-wide, but a process drives one context per device, every kernel that
// touches a pooled block is ordered on that context's main stream, and the one user of the
// side stream (the pre-launched product inside run_irlba) works on blocks run_irlba owns
// and is joined (both streams synchronised) before run_irlba gives them back.
// m3b_destroy() releases the cache of its device.
// ---------------------------------------------------------------------------
#include <map>
#include <mutex>
#include <unordered_map>
namespace {
struct DevPool {
  std::mutex mu;
  std::multimap<size_t, void*> free_blocks[16];  // per device
  std::unordered_map<void*, std::pair<size_t, int>> live;  // ptr -> (bytes, device)
  size_t cached_bytes[16] = {0};
};
DevPool& pool() {
  static DevPool p;
  return p;
}
}  // namespace

static cudaError_t pool_alloc(void** out, size_t bytes) {
  int dev = 0;
  cudaGetDevice(&dev);
  dev &= 15;
  // size classes: small requests (the ~100 tables of a layout build differ by a few bytes from call to
  // call) are rounded up to a power of two so that they find a cached block instead of going to
  // cudaMalloc (~0.1-1 ms each: 40 ms per m3b_select_genes on a small matrix before this)
  if (bytes <= 65536) {
    bytes = 65536;
  } else if (bytes <= (size_t)64 << 20) {
    size_t p2 = 65536;
    while (p2 < bytes) p2 <<= 1;
    bytes = p2;
  } else {
    bytes = (bytes + 511) & ~(size_t)511;
  }
  DevPool& P = pool();
  {
    std::lock_guard<std::mutex> g(P.mu);
    auto it = P.free_blocks[dev].lower_bound(bytes);
    if (it != P.free_blocks[dev].end() && it->first <= bytes + bytes / 8 + 4096) {
      *out = it->second;
      P.live[*out] = {it->first, dev};
      P.cached_bytes[dev] -= it->first;
      P.free_blocks[dev].erase(it);
      return cudaSuccess;
    }
  }
  cudaError_t e = cudaMalloc(out, bytes);
  if (e == cudaErrorMemoryAllocation) {  // give the cache back to the driver and retry once
    cudaGetLastError();
    dev_pool_release(dev);
    e = cudaMalloc(out, bytes);
  }
  if (e == cudaSuccess) {
    std::lock_guard<std::mutex> g(P.mu);
    P.live[*out] = {bytes, dev};
  }
  return e;
}

void dev_pool_release(int dev) {
  DevPool& P = pool();
  std::vector<void*> blocks;
  {
    std::lock_guard<std::mutex> g(P.mu);
    for (auto& kv : P.free_blocks[dev & 15]) blocks.push_back(kv.second);
    P.free_blocks[dev & 15].clear();
    P.cached_bytes[dev & 15] = 0;
  }
  for (void* b : blocks) cudaFree(b);
}

template <typename T>
int dev_alloc(m3b_ctx* ctx, T** p, size_t n) {
  *p = nullptr;
  if (n == 0) n = 1;
  CK(ctx, pool_alloc((void**)p, n * sizeof(T)));
  return M3B_OK;
}
template int dev_alloc<double>(m3b_ctx*, double**, size_t);
template int dev_alloc<int>(m3b_ctx*, int**, size_t);
template int dev_alloc<long long>(m3b_ctx*, long long**, size_t);
template int dev_alloc<ItemDesc>(m3b_ctx*, ItemDesc**, size_t);
template int dev_alloc<int4>(m3b_ctx*, int4**, size_t);
template int dev_alloc<int2>(m3b_ctx*, int2**, size_t);
template int dev_alloc<unsigned int>(m3b_ctx*, unsigned int**, size_t);
template int dev_alloc<unsigned char>(m3b_ctx*, unsigned char**, size_t);
template int dev_alloc<unsigned short>(m3b_ctx*, unsigned short**, size_t);
template int dev_alloc<unsigned long long>(m3b_ctx*, unsigned long long**, size_t);

void dev_free(void* p) {
  if (!p) return;
  DevPool& P = pool();
  std::lock_guard<std::mutex> g(P.mu);
  auto it = P.live.find(p);
  if (it == P.live.end()) {  // not ours (should not happen): hand it to the driver
    cudaFree(p);
    return;
  }
  const size_t bytes = it->second.first;
  const int dev = it->second.second;
  P.live.erase(it);
  P.free_blocks[dev].insert({bytes, p});
  P.cached_bytes[dev] += bytes;
}

void TiledLayout::free_all() {
  if (unit) {
    unit->partial = nullptr;  // shared with this layout
    unit->row_slot_ptr = nullptr;
    unit->free_all();
    delete unit;
  }
  flat_plan_free(plan[0]);
  flat_plan_free(plan[1]);
  dev_free(item_scale);
  dev_free(tile_len);
  dev_free(item_tile);
  dev_free(item_cum);
  dev_free(idx);
  dev_free(val);
  dev_free(items);
  dev_free(chunks);
  dev_free(tile_chunk_ptr);
  dev_free(cta_tile);
  dev_free(counters);
  dev_free(row_item_ptr);
  dev_free(item_slot);
  dev_free(row_slot_ptr);
  dev_free(partial);
  *this = TiledLayout();
}

// ---------------------------------------------------------------------------
// streaming loads: the CSC arrays are read exactly once per pass, so keep them
// out of L1 (ld.global.nc.L1::no_allocate) and leave L1/L2 to the gathers.
// ---------------------------------------------------------------------------
__device__ __forceinline__ double ld_stream_f64(const double* p) {
  double v;
  asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ int ld_stream_s32(const int* p) {
  int v;
  asm volatile("ld.global.nc.L1::no_allocate.s32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}

// ---------------------------------------------------------------------------
// ingest helpers
// ---------------------------------------------------------------------------
__global__ void k_offset_colptr(const int* __restrict__ p_chunk, long long n, long long base,
                                long long* __restrict__ p_out, int* __restrict__ flag) {
  long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (k < n) {
    p_out[k] = base + (long long)p_chunk[k];
    if (k + 1 < n && p_chunk[k] > p_chunk[k + 1]) atomicOr(flag, 1);  // column pointers must not decrease
  }
}

// Validation of an appended chunk (the kernels downstream index per-gene tables with these values
// and rely on one entry per (gene, cell)): every row index in [0, n_genes), strictly increasing
// inside a column -- what a dgCMatrix guarantees.  One warp per cell; offsets are clamped to the
// chunk so that a bad column pointer cannot send the check itself out of bounds.
__global__ void __launch_bounds__(256) k_validate_rows(const long long* __restrict__ p, const int* __restrict__ gi,
                                                       long long n_cells, long long lo_all, long long hi_all, int n_genes,
                                                       int* __restrict__ flag) {
  const int lane = threadIdx.x & 31;
  long long warp = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  int bad = 0;
  for (long long c = warp; c < n_cells; c += nwarps) {
    const long long lo = min(max(p[c], lo_all), hi_all), hi = min(max(p[c + 1], lo_all), hi_all);
    for (long long k = lo + lane; k < hi; k += 32) {
      const int g = gi[k];
      if (g < 0 || g >= n_genes) bad |= 2;
      if (k + 1 < hi && gi[k + 1] <= g) bad |= 4;
    }
  }
  if (bad) atomicOr(flag, bad);
}

int ingest_append(m3b_mat* m, long long n_cells_chunk, const int* p, const int* i, const double* x) {
  m3b_ctx* ctx = m->ctx;
  if (m->ended) {
    m3b_set_error(ctx, "m3b_matrix_append_csc after m3b_matrix_end");
    return M3B_E_STATE;
  }
  if (n_cells_chunk < 0 || m->cells_appended + n_cells_chunk > m->n_cells || p[0] != 0) {
    m3b_set_error(ctx, "append: chunk of %lld cells does not fit (have %lld of %lld) or p[0] != 0",
                  n_cells_chunk, m->cells_appended, m->n_cells);
    return M3B_E_DIMS;
  }
  long long nnz_chunk = p[n_cells_chunk];
  if (nnz_chunk < 0) {
    m3b_set_error(ctx, "append: negative nnz");
    return M3B_E_DIMS;
  }
  if (m->nnz + nnz_chunk > m->nnz_cap) {  // grow
    long long cap = std::max(m->nnz + nnz_chunk, (long long)(m->nnz_cap * 1.5) + 1024);
    int* ni = nullptr;
    double* nx = nullptr;
    DevScope grow;  // the new arrays, until they replace the old ones
    grow.own(&ni);
    grow.own(&nx);
    RET(dev_alloc(ctx, &ni, (size_t)cap));
    RET(dev_alloc(ctx, &nx, (size_t)cap));
    if (m->nnz) {
      CK(ctx, cudaMemcpyAsync(ni, m->i, m->nnz * sizeof(int), cudaMemcpyDeviceToDevice, ctx->stream));
      CK(ctx, cudaMemcpyAsync(nx, m->x, m->nnz * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
      CK(ctx, cudaStreamSynchronize(ctx->stream));
    }
    dev_free(m->i);
    dev_free(m->x);
    m->i = ni;
    m->x = nx;
    ni = nullptr;  // handed over
    nx = nullptr;
    m->nnz_cap = cap;
  }
  int* d_pchunk = nullptr;
  DevScope guard;  // d_pchunk goes back to the pool on every exit
  guard.own(&d_pchunk);
  RET(dev_alloc(ctx, &d_pchunk, (size_t)n_cells_chunk + 1));
  CK(ctx, cudaMemcpyAsync(d_pchunk, p, (n_cells_chunk + 1) * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
  CK(ctx, cudaMemcpyAsync(m->i + m->nnz, i, nnz_chunk * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
  CK(ctx, cudaMemcpyAsync(m->x + m->nnz, x, nnz_chunk * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  long long n = n_cells_chunk + 1;
  if (!ctx->ingest_flag) {
    CK(ctx, cudaMalloc((void**)&ctx->ingest_flag, sizeof(int)));
    CK(ctx, cudaMemsetAsync(ctx->ingest_flag, 0, sizeof(int), ctx->stream));
  }
  k_offset_colptr<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(d_pchunk, n, m->nnz, m->p + m->cells_appended,
                                                                        ctx->ingest_flag);
  count_launch(ctx);
  if (n_cells_chunk > 0 && nnz_chunk > 0) {
    const int vblocks = (int)std::min<long long>((n_cells_chunk + 7) / 8, (long long)ctx->sm_count * 8);
    k_validate_rows<<<vblocks, 256, 0, ctx->stream>>>(m->p + m->cells_appended, m->i, n_cells_chunk, m->nnz,
                                                      m->nnz + nnz_chunk, m->n_genes, ctx->ingest_flag);
    count_launch(ctx);
  }
  CK(ctx, cudaGetLastError());
  int bad = 0;
  CK(ctx, cudaMemcpyAsync(&bad, ctx->ingest_flag, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  CK(ctx, cudaStreamSynchronize(ctx->stream));
  if (bad) {
    cudaMemsetAsync(ctx->ingest_flag, 0, sizeof(int), ctx->stream);
    m3b_set_error(ctx, "append: malformed CSC chunk (%s%s%s): the input must be a valid dgCMatrix slice -- non-decreasing "
                       "column pointers, row indices in [0, n_genes) and strictly increasing inside every column",
                  (bad & 1) ? "column pointers decrease; " : "", (bad & 2) ? "row index out of range; " : "",
                  (bad & 4) ? "row indices not strictly increasing within a column" : "");
    return M3B_E_DIMS;
  }
  ctx->stats.h2d_bytes += (n_cells_chunk + 1) * 4 + nnz_chunk * 12;
  m->nnz += nnz_chunk;
  m->cells_appended += n_cells_chunk;
  return M3B_OK;
}

// ---------------------------------------------------------------------------
// K1: per-cell totals.  One warp per cell (column of the CSC): coalesced 8-byte
// loads, fixed-order shuffle reduction.  Sums of integers < 2^53 are exact in
// fp64 regardless of order, so the totals are bit-identical to colSums().
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_cell_totals(const long long* __restrict__ p, const double* __restrict__ x,
                                                     long long n_cells, int round_exprs, double* __restrict__ tot) {
  const int lane = threadIdx.x & 31;
  long long warp = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long c = warp; c < n_cells; c += nwarps) {
    const long long lo = p[c], hi = p[c + 1];
    double acc = 0.0;
    long long k = lo + lane;
    for (; k + 96 < hi; k += 128) {  // four independent 8-byte loads in flight per lane
      const double v0 = ld_stream_f64(x + k), v1 = ld_stream_f64(x + k + 32);
      const double v2 = ld_stream_f64(x + k + 64), v3 = ld_stream_f64(x + k + 96);
      if (round_exprs) acc += (rint(v0) + rint(v1)) + (rint(v2) + rint(v3));  // R's round(): half-to-even
      else acc += (v0 + v1) + (v2 + v3);
    }
    for (; k < hi; k += 32) {
      const double v = ld_stream_f64(x + k);
      acc += round_exprs ? rint(v) : v;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) tot[c] = acc;
  }
}

int k1_cell_totals(m3b_mat* m, int round_exprs, double* d_tot) {
  m3b_ctx* ctx = m->ctx;
  if (m->n_cells == 0) return M3B_OK;
  int blocks = (int)std::min<long long>((m->n_cells + 7) / 8, (long long)ctx->sm_count * 8);
  cudaEventRecord(ctx->evk0, ctx->stream);
  k_cell_totals<<<blocks, 256, 0, ctx->stream>>>(m->p, m->x, m->n_cells, round_exprs, d_tot);
  cudaEventRecord(ctx->evk1, ctx->stream);
  count_launch(ctx);
  CK(ctx, cudaGetLastError());
  ctx->k1_timed = true;
  return M3B_OK;
}

// ---------------------------------------------------------------------------
// K2: fused size-factor divide + log.  y = f(x / sf_c) - f(0); f(0) is non-zero
// only on the reference's dense paths (pseudo_count != 1 for log, != 0 for
// size_only, R/preprocess_cds.R:274-276,283) and is carried as m->f0.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_normalize(const long long* __restrict__ p, const double* __restrict__ x,
                                                   const double* __restrict__ sf, long long n_cells, int method,
                                                   double pc, double f0, double* __restrict__ y,
                                                   double* __restrict__ unit_val) {
  const int lane = threadIdx.x & 31;
  long long warp = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long c = warp; c < n_cells; c += nwarps) {
    const long long lo = p[c], hi = p[c + 1];
    const double s = (method == M3B_NORM_NONE || method == M3B_NORM_BINARY) ? 1.0 : sf[c];
    auto f = [&](double v) -> double {
      switch (method) {
        case M3B_NORM_LOG2: return log2(v / s + pc) - f0;
        case M3B_NORM_LOG10: return log10(v / s + pc);
        case M3B_NORM_SIZE_ONLY: return v / s;  // + pc is the implicit constant f0
        case M3B_NORM_BINARY: return (v > 0.0) ? 1.0 : 0.0;
        default: return v;
      }
    };
    // Counts are small integers (51-79% ones in the reference's fixtures): lane l keeps f(l+1)
    // for this cell and entries with an integer value 1..32 are served by a shuffle instead of
    // an fp64 log (which made the pass instruction-bound).  Same function, same argument =>
    // bit-identical to evaluating f per entry.
    const bool use_tbl = (method == M3B_NORM_LOG2 || method == M3B_NORM_LOG10);
    const double tbl = use_tbl ? f((double)(lane + 1)) : 0.0;
    if (lane == 0) unit_val[c] = f(1.0);  // value of the cell's count-1 entries (the UNIT layouts)
    auto g = [&](double v) -> double {
      if (!use_tbl) return f(v);
      const int iv = (int)v;
      const bool hit = (v >= 1.0) && (v <= 32.0) && ((double)iv == v);
      const double t = __shfl_sync(0xffffffffu, tbl, hit ? iv - 1 : 0);
      return hit ? t : f(v);
    };
    long long k = lo + lane;
    // the trip count is made warp-uniform (shuffles inside): lanes past the end compute on 0
    const long long span = hi - lo;
    const long long iters = (span + 31) / 32;
    long long it = 0;
    for (; it + 3 < iters; it += 4, k += 128) {  // four independent loads in flight per lane
      const bool b0 = k < hi, b1 = k + 32 < hi, b2 = k + 64 < hi, b3 = k + 96 < hi;
      const double v0 = b0 ? ld_stream_f64(x + k) : 0.0, v1 = b1 ? ld_stream_f64(x + k + 32) : 0.0;
      const double v2 = b2 ? ld_stream_f64(x + k + 64) : 0.0, v3 = b3 ? ld_stream_f64(x + k + 96) : 0.0;
      const double r0 = g(v0), r1 = g(v1), r2 = g(v2), r3 = g(v3);
      if (b0) y[k] = r0;
      if (b1) y[k + 32] = r1;
      if (b2) y[k + 64] = r2;
      if (b3) y[k + 96] = r3;
    }
    for (; it < iters; ++it, k += 32) {
      const bool b0 = k < hi;
      const double r0 = g(b0 ? ld_stream_f64(x + k) : 0.0);
      if (b0) y[k] = r0;
    }
  }
}

int k2_normalize(m3b_mat* m, const double* d_sf, int method, double pc) {
  m3b_ctx* ctx = m->ctx;
  if (!m->y) RET(dev_alloc(ctx, &m->y, (size_t)std::max<long long>(m->nnz, 1)));
  if (!m->unit_val) RET(dev_alloc(ctx, &m->unit_val, (size_t)std::max<long long>(m->n_cells, 1)));
  double f0 = 0.0;
  if (method == M3B_NORM_LOG2) f0 = (pc == 1.0) ? 0.0 : std::log2(pc);
  if (method == M3B_NORM_SIZE_ONLY) f0 = pc;
  if (method == M3B_NORM_LOG10) f0 = 0.0;
  m->f0 = f0;
  if (m->n_cells) {
    int blocks = (int)std::min<long long>((m->n_cells + 7) / 8, (long long)ctx->sm_count * 8);
    cudaEventRecord(ctx->evk0, ctx->stream);
    k_normalize<<<blocks, 256, 0, ctx->stream>>>(m->p, m->x, d_sf, m->n_cells, method, pc, f0, m->y, m->unit_val);
    cudaEventRecord(ctx->evk1, ctx->stream);
    ctx->k2_timed = true;
    count_launch(ctx);
    CK(ctx, cudaGetLastError());
  }
  return M3B_OK;
}

// ---------------------------------------------------------------------------
// host-side construction of the item / unit / row-pointer tables from the
// per-(tile,row) entry counts.  cnt[t*rows_in + r] = true entries of input row r
// in tile t; row_map (nullable) maps input rows to output rows (-1 = dropped).
// base_out[t*rows_in + r] = entry offset of the run (only meaningful for runs
// that get storage; storage is laid out for ALL input rows so that a later
// re-map needs no data movement).
// ---------------------------------------------------------------------------

static inline long long pad4(long long n) { return (n + 3) & ~3LL; }
static inline long long padn(long long n, int pad) { return (n + pad - 1) / pad * pad; }

// shape of a layout's runs / items: general layouts pad runs to 4 entries, unit layouts to 8
struct ItemShape {
  int pad, item_max, item_tail;
};
static const ItemShape SHAPE_GENERAL = {4, M3B_ITEM_MAX, M3B_ITEM_TAIL};
static const ItemShape SHAPE_UNIT = {8, M3B_UITEM_MAX, M3B_UITEM_TAIL};

static void compute_bases(const std::vector<int>& cnt, int n_tiles, long long rows_in, HostTables& T,
                          const ItemShape& sh) {
  T.base.resize((size_t)n_tiles * rows_in);
  long long off = 0;
  for (size_t k = 0; k < T.base.size(); ++k) {
    T.base[k] = off;
    off += padn(cnt[k], sh.pad);
  }
  T.n_entries = off;
}

// partial slots: contiguous per output row -- the row's general items (tile, segment order)
// followed by its unit items
static void assign_slots(HostTables& G, HostTables* U, long long rows_out) {
  std::vector<int>& rsp = G.row_slot_ptr;
  rsp.assign((size_t)rows_out + 1, 0);
  for (const ItemDesc& d : G.items) rsp[(size_t)d.row + 1]++;
  if (U)
    for (const ItemDesc& d : U->items) rsp[(size_t)d.row + 1]++;
  for (long long r = 0; r < rows_out; ++r) rsp[r + 1] += rsp[r];
  std::vector<int> fill(rsp.begin(), rsp.end() - 1);
  G.item_slot.resize(G.items.size());
  for (size_t i = 0; i < G.items.size(); ++i) G.item_slot[i] = fill[G.items[i].row]++;
  if (U) {
    U->item_slot.resize(U->items.size());
    for (size_t i = 0; i < U->items.size(); ++i) U->item_slot[i] = fill[U->items[i].row]++;
    U->row_slot_ptr.clear();
  }
}

static void build_items(const std::vector<int>& cnt, int n_tiles, long long rows_in, const int* row_map,
                        long long rows_out, int sm_count, HostTables& T, const ItemShape& sh) {
  T.items.clear();
  T.chunks.clear();
  T.tile_chunk_ptr.assign(n_tiles + 1, 0);
  T.row_item_ptr.assign((size_t)n_tiles * (rows_out + 1), 0);
  T.nnz = 0;
  // per-tile: rows are visited in OUTPUT-row order so that row_item_ptr is monotone
  std::vector<long long> in_of_out;
  if (row_map) {
    in_of_out.assign(rows_out, -1);
    for (long long r = 0; r < rows_in; ++r)
      if (row_map[r] >= 0) in_of_out[row_map[r]] = r;
  }
  std::vector<long long> tile_len(n_tiles, 0);
  for (int t = 0; t < n_tiles; ++t)
    for (long long ro = 0; ro < rows_out; ++ro) {
      long long ri = row_map ? in_of_out[ro] : ro;
      if (ri >= 0) tile_len[t] += padn(cnt[(size_t)t * rows_in + ri], sh.pad);
    }
  for (int t = 0; t < n_tiles; ++t) {
    int* rip = &T.row_item_ptr[(size_t)t * (rows_out + 1)];
    // the last ~20% of a tile's entries are cut into short items/chunks so that the
    // warps of a launch run dry together (tail = one short chunk, not one long one)
    const long long tail_from = tile_len[t] - tile_len[t] / 5;
    long long seen = 0;
    T.tile_chunk_ptr[t] = (int)T.chunks.size();
    int chunk_begin = (int)T.items.size();
    long long chunk_acc = 0;
    for (long long ro = 0; ro < rows_out; ++ro) {
      rip[ro] = (int)T.items.size();
      long long ri = row_map ? in_of_out[ro] : ro;
      if (ri < 0) continue;
      long long c = cnt[(size_t)t * rows_in + ri];
      if (c == 0) continue;
      T.nnz += c;
      long long len = padn(c, sh.pad), start = T.base[(size_t)t * rows_in + ri];
      while (len > 0) {
        const int lim = (seen >= tail_from) ? sh.item_tail : sh.item_max;
        int l = (int)std::min<long long>(len, lim);
        T.items.push_back(ItemDesc{start, l, (int)ro});
        start += l;
        len -= l;
        seen += l;
        chunk_acc += l + 32;  // +32: per-item overhead, so that runs of tiny rows are split too
        if (chunk_acc >= lim) {
          T.chunks.push_back(make_int2(chunk_begin, (int)T.items.size()));
          chunk_begin = (int)T.items.size();
          chunk_acc = 0;
        }
      }
    }
    if (chunk_begin < (int)T.items.size()) T.chunks.push_back(make_int2(chunk_begin, (int)T.items.size()));
    rip[rows_out] = (int)T.items.size();
  }
  T.tile_chunk_ptr[n_tiles] = (int)T.chunks.size();
  // CTA -> starting tile, proportional to the tiles' entry counts (every non-empty tile gets >= 1)
  T.cta_tile.assign(sm_count, 0);
  long long total = 0;
  for (int t = 0; t < n_tiles; ++t) total += tile_len[t];
  if (total > 0) {
    long long acc = 0;
    int b = 0;
    for (int t = 0; t < n_tiles && b < sm_count; ++t) {
      acc += tile_len[t];
      int upto = (int)((double)acc / (double)total * sm_count + 0.5);
      if (tile_len[t] > 0 && upto <= b) upto = b + 1;
      upto = std::min(upto, sm_count);
      for (; b < upto; ++b) T.cta_tile[b] = t;
    }
    for (; b < sm_count; ++b) T.cta_tile[b] = n_tiles - 1;
  }
}

// n_partial: slots of the shared partial array (general + unit items); a unit layout borrows
// its parent's partial and row_slot_ptr (upload the parent first)
static int upload_tables(m3b_ctx* ctx, TiledLayout& L, const HostTables& T, size_t n_partial,
                         const TiledLayout* parent = nullptr) {
  dev_free(L.items);
  dev_free(L.chunks);
  dev_free(L.tile_chunk_ptr);
  dev_free(L.cta_tile);
  dev_free(L.counters);
  dev_free(L.row_item_ptr);
  dev_free(L.item_slot);
  if (!L.is_unit) {
    dev_free(L.row_slot_ptr);
    dev_free(L.partial);
  }
  L.item_slot = nullptr;
  L.row_slot_ptr = nullptr;
  L.items = nullptr;
  L.chunks = nullptr;
  L.tile_chunk_ptr = nullptr;
  L.cta_tile = nullptr;
  L.counters = nullptr;
  L.row_item_ptr = nullptr;
  L.partial = nullptr;
  L.n_items = (long long)T.items.size();
  L.n_chunks = (int)T.chunks.size();
  L.n_cta = (int)T.cta_tile.size();
  L.nnz = T.nnz;
  RET(dev_alloc(ctx, &L.items, T.items.size()));
  RET(dev_alloc(ctx, &L.chunks, T.chunks.size()));
  RET(dev_alloc(ctx, &L.tile_chunk_ptr, T.tile_chunk_ptr.size()));
  RET(dev_alloc(ctx, &L.cta_tile, T.cta_tile.size()));
  RET(dev_alloc(ctx, &L.counters, (size_t)L.n_tiles + 1));
  RET(dev_alloc(ctx, &L.row_item_ptr, T.row_item_ptr.size()));
  RET(dev_alloc(ctx, &L.item_slot, T.item_slot.size()));
  if (!T.item_slot.empty())
    CK(ctx, cudaMemcpyAsync(L.item_slot, T.item_slot.data(), T.item_slot.size() * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
  if (L.is_unit) {
    L.partial = parent->partial;
    L.row_slot_ptr = parent->row_slot_ptr;
  } else {
    RET(dev_alloc(ctx, &L.partial, n_partial));
    RET(dev_alloc(ctx, &L.row_slot_ptr, T.row_slot_ptr.size()));
    CK(ctx, cudaMemcpyAsync(L.row_slot_ptr, T.row_slot_ptr.data(), T.row_slot_ptr.size() * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
  }
  if (!T.items.empty())
    CK(ctx, cudaMemcpyAsync(L.items, T.items.data(), T.items.size() * sizeof(ItemDesc), cudaMemcpyHostToDevice, ctx->stream));
  if (!T.chunks.empty())
    CK(ctx, cudaMemcpyAsync(L.chunks, T.chunks.data(), T.chunks.size() * sizeof(int2), cudaMemcpyHostToDevice, ctx->stream));
  CK(ctx, cudaMemcpyAsync(L.tile_chunk_ptr, T.tile_chunk_ptr.data(), T.tile_chunk_ptr.size() * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
  if (!T.cta_tile.empty())
    CK(ctx, cudaMemcpyAsync(L.cta_tile, T.cta_tile.data(), T.cta_tile.size() * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
  CK(ctx, cudaMemsetAsync(L.counters, 0, ((size_t)L.n_tiles + 1) * sizeof(unsigned int), ctx->stream));
  CK(ctx, cudaMemcpyAsync(L.row_item_ptr, T.row_item_ptr.data(), T.row_item_ptr.size() * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
  CK(ctx, cudaStreamSynchronize(ctx->stream));
  return M3B_OK;
}

// The product kernel gives every tile a whole number of CTAs (or every CTA a whole number of tiles),
// so the tile count decides the balance: 98 tiles on 148 SMs leave 50 tiles with two CTAs and 48
// with one -- the launch then lasts as long as a one-CTA tile, 1.5x the balanced time (measured at
// 2M cells on one GPU: genes-out 2.72 ms against 2.17 ms for cells-out on the same entries).
// So: the smallest tile count >= cols / tile_w_max whose whole-number split uses >= 95 % of the
// CTAs (or, failing that, the best one up to twice the minimum).
static double tiling_efficiency(int nt, int P) {
  if (nt <= P) return (double)nt * (double)(P / nt) / (double)P;
  return (double)nt / ((double)((nt + P - 1) / P) * (double)P);
}
static void choose_tiling(long long cols, int tile_w_max, int P, int* tile_w, int* n_tiles) {
  if (cols <= 0) {
    *tile_w = 32;
    *n_tiles = 1;
    return;
  }
  int nt = (int)((cols + tile_w_max - 1) / tile_w_max);
  if (nt > 1 && P > 1 && !getenv("M3B_PLAIN_TILING")) {
    const int hi = (int)std::min<long long>(std::max(2 * nt, nt + P), std::max<long long>(nt, cols / 2048));
    int best = nt;
    double best_e = tiling_efficiency(nt, P);
    for (int c = nt; c <= hi && best_e < 0.95; ++c) {
      const double e = tiling_efficiency(c, P);
      if (e > best_e + 1e-9) {
        best = c;
        best_e = e;
      }
    }
    nt = best;
  }
  long long tw = (cols + nt - 1) / nt;
  tw = (tw + 31) & ~31LL;
  if (tw > tile_w_max) tw = tile_w_max;
  *tile_w = (int)tw;
  *n_tiles = (int)((cols + tw - 1) / tw);
}

// ---------------------------------------------------------------------------
// Layout B (gene-major, tiled along cells): stable counting-sort transpose.
// A CHUNK is a run of consecutive cells of one tile owned by ONE warp, which
// walks its cells in order; within a cell the row indices are distinct, so the
// per-(chunk,row) cursor hands out slots in cell order => the transpose is
// deterministic and every gene row is sorted by cell.
// ---------------------------------------------------------------------------
// Entry classes of a build pass: 0 = every stored entry (unsplit layouts), 1 = raw count != 1
// (general layout of a split matrix), 2 = raw count == 1 (unit layout).
__device__ __forceinline__ bool in_class(int cls, const double* __restrict__ xraw, long long k) {
  if (cls == 0) return true;
  const bool one = ld_stream_f64(xraw + k) == 1.0;
  return (cls == 2) == one;
}

#define TB_THREADS 512
#define TB_WARPS (TB_THREADS / 32)
#define TB_GSEG_MAX 24576  // most selected rows per pass: 96 KB of cursors + 96 KB of cell masks in shared memory
// Rows per pass (M3B_GSEG overrides).  Smaller passes were tried to keep the partially written
// output sectors L2-resident; the extra passes over the index array cost more than they saved
// (cfg3 shape: 1024/2048/4096/8192/24576 rows => 405/212/122/75/56 ms), so one pass takes as many
// rows as shared memory holds.
#define TB_GSEG_DEFAULT 24576

// A CHUNK is a run of consecutive cells of one tile, owned by ONE CTA.  Pass 1 counts the
// chunk's entries per selected row in a shared-memory histogram (rows [g0, g1) of this pass).
__global__ void __launch_bounds__(TB_THREADS) k_count_B(const long long* __restrict__ p, const int* __restrict__ gi,
                                                        const double* __restrict__ xraw, int cls,
                                                        const int* __restrict__ sel_of_gene, int n_sel, int g0, int g1,
                                                        int tile_w, int chunks_per_tile, int cells_per_chunk,
                                                        long long n_cells, int* __restrict__ H) {
  extern __shared__ int hist[];
  const long long chunk = blockIdx.x;
  const long long t = chunk / chunks_per_tile, s = chunk % chunks_per_tile;
  const long long c0 = min(t * tile_w + s * (long long)cells_per_chunk, n_cells);
  const long long c1 = min(min(c0 + cells_per_chunk, (t + 1) * (long long)tile_w), n_cells);
  const int ng = g1 - g0;
  for (int r = threadIdx.x; r < ng; r += blockDim.x) hist[r] = 0;
  __syncthreads();
  if (c0 < c1) {
    const long long lo = p[c0], hi = p[c1];  // the chunk's entries are one contiguous range
    for (long long k = lo + threadIdx.x; k < hi; k += blockDim.x) {
      const int r = sel_of_gene[ld_stream_s32(gi + k)];
      if (r >= g0 && r < g1 && in_class(cls, xraw, k)) atomicAdd(&hist[r - g0], 1);
    }
  }
  __syncthreads();
  int* Hc = H + chunk * n_sel + g0;
  for (int r = threadIdx.x; r < ng; r += blockDim.x) Hc[r] = hist[r];
}

// exclusive scan over the chunks of a tile, per selected row; cnt = tile total
__global__ void k_scan_chunks(int* __restrict__ H, int n_sel, int chunks_per_tile, int n_tiles, int* __restrict__ cnt) {
  long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (k >= (long long)n_tiles * n_sel) return;
  long long t = k / n_sel, r = k % n_sel;
  int run = 0;
  for (int s = 0; s < chunks_per_tile; ++s) {
    int* h = H + ((t * chunks_per_tile + s) * (long long)n_sel + r);
    int v = *h;
    *h = run;
    run += v;
  }
  cnt[k] = run;
}

// pads of a run: general layouts idx 0 / val 0.0; unit layouts (val == nullptr) idx == tile_w,
// the zero slot behind the staged tile
__global__ void k_fill_pads(const int* __restrict__ cnt, const long long* __restrict__ base, long long n_runs, int pad,
                            int pad_idx, unsigned short* __restrict__ idx, double* __restrict__ val) {
  long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (k >= n_runs) return;
  int c = cnt[k];
  long long b = base[k];
  for (int e = c; e < (c + pad - 1) / pad * pad; ++e) {
    idx[b + e] = (unsigned short)pad_idx;
    if (val) val[b + e] = 0.0;
  }
}

// Pass 2: stable scatter.  The chunk's write cursors (one per selected row of this pass) live
// in shared memory.  The CTA walks its cells 16 at a time, one warp per cell:
//   A  every warp sets its bit in mask[row] for the rows its cell contains,
//   B  an entry's slot is cursor[row] + (number of EARLIER cells of the group that contain the
//      row) -- so every gene row ends up sorted by cell, deterministically, without a global
//      atomic and without serialising the cells,
//   C  the first cell of the group that contains a row advances its cursor and clears the mask.
// (The first version used one warp per chunk with the cursors in a 0.5 GB global table: the
// random read-modify-writes on it made the transpose DRAM-latency bound, 62 ms at 474 M nnz.)
__global__ void __launch_bounds__(TB_THREADS) k_scatter_B(const long long* __restrict__ p, const int* __restrict__ gi,
                                                          const double* __restrict__ y, const double* __restrict__ xraw,
                                                          int cls, const int* __restrict__ sel_of_gene,
                                                          int n_sel, int g0, int g1, int tile_w, int chunks_per_tile,
                                                          int cells_per_chunk, long long n_cells, const int* __restrict__ H,
                                                          const long long* __restrict__ base,
                                                          unsigned short* __restrict__ idx, double* __restrict__ val) {
  extern __shared__ int sm_tb[];
  const int ng = g1 - g0;
  int* cursor = sm_tb;                                         // [ng] slot inside the (tile,row) run
  unsigned int* mask = reinterpret_cast<unsigned int*>(sm_tb + ng);  // [ng] cells of the current group
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const long long chunk = blockIdx.x;
  const long long t = chunk / chunks_per_tile, s = chunk % chunks_per_tile;
  const long long c0 = min(t * tile_w + s * (long long)cells_per_chunk, n_cells);
  const long long c1 = min(min(c0 + cells_per_chunk, (t + 1) * (long long)tile_w), n_cells);
  const int* Hc = H + chunk * n_sel + g0;
  const long long* bt = base + t * n_sel + g0;
  for (int r = threadIdx.x; r < ng; r += blockDim.x) {
    cursor[r] = Hc[r];
    mask[r] = 0u;
  }
  __syncthreads();
  const unsigned int below = (1u << w) - 1u;
  for (long long cg = c0; cg < c1; cg += TB_WARPS) {
    const long long c = cg + w;
    const bool have = c < c1;
    const long long lo = have ? p[c] : 0, hi = have ? p[c + 1] : 0;
    // every phase walks the cell's entries 4 x 32 at a time so that four index loads and four
    // map look-ups are in flight per lane (one at a time made the kernel latency-bound)
#define TB_FOR_ROWS(BODY)                                                       \
    for (long long k0 = lo; k0 < hi; k0 += 128) {                                \
      int rr[4];                                                                 \
      long long kk[4];                                                           \
      _Pragma("unroll") for (int u = 0; u < 4; ++u) {                            \
        kk[u] = k0 + lane + 32 * u;                                              \
        rr[u] = (kk[u] < hi && in_class(cls, xraw, kk[u])) ? gi[kk[u]] : -1;     \
      }                                                                          \
      _Pragma("unroll") for (int u = 0; u < 4; ++u) rr[u] = (rr[u] >= 0) ? sel_of_gene[rr[u]] : -1; \
      _Pragma("unroll") for (int u = 0; u < 4; ++u) {                            \
        const int r = rr[u];                                                     \
        const long long k = kk[u];                                               \
        if (r >= g0 && r < g1) { BODY }                                          \
      }                                                                          \
    }
    TB_FOR_ROWS(atomicOr(&mask[r - g0], 1u << w); (void)k;)  // A
    __syncthreads();
    const int local = (int)(c - t * tile_w);
    TB_FOR_ROWS(  // B
        const int q = r - g0;
        const long long dst = bt[q] + cursor[q] + __popc(mask[q] & below);
        idx[dst] = (unsigned short)local;
        if (cls != 2) val[dst] = ld_stream_f64(y + k);)
    __syncthreads();
    TB_FOR_ROWS(  // C
        const int q = r - g0; (void)k;
        const unsigned int mq = mask[q];
        if ((mq & below) == 0u) {  // this warp's cell is the first of the group with this row
          cursor[q] += __popc(mq);
          mask[q] = 0u;
        })
#undef TB_FOR_ROWS
    __syncthreads();
  }
}

// ---- split matrices: both entry classes (general = raw count != 1, unit = raw count == 1) in ONE
// pass.  The per-row state of the two classes shares a 32-bit word (general in the low half, unit
// in the high half: counts inside a tile are <= 28672 < 2^16, the group mask has 16 bits), so the
// shared-memory footprint is that of the single-class kernels and every entry is visited once per
// phase instead of twice.
__global__ void __launch_bounds__(TB_THREADS) k_count_B2(const long long* __restrict__ p, const int* __restrict__ gi,
                                                         const double* __restrict__ xraw,
                                                         const int* __restrict__ sel_of_gene, int n_sel, int g0, int g1,
                                                         int tile_w, int chunks_per_tile, int cells_per_chunk,
                                                         long long n_cells, int* __restrict__ Hg, int* __restrict__ Hu) {
  extern __shared__ int hist[];  // [2][ng]
  const long long chunk = blockIdx.x;
  const long long t = chunk / chunks_per_tile, s = chunk % chunks_per_tile;
  const long long c0 = min(t * tile_w + s * (long long)cells_per_chunk, n_cells);
  const long long c1 = min(min(c0 + cells_per_chunk, (t + 1) * (long long)tile_w), n_cells);
  const int ng = g1 - g0;
  for (int r = threadIdx.x; r < 2 * ng; r += blockDim.x) hist[r] = 0;
  __syncthreads();
  if (c0 < c1) {
    const long long lo = p[c0], hi = p[c1];
    for (long long k = lo + threadIdx.x; k < hi; k += blockDim.x) {
      const int r = sel_of_gene[ld_stream_s32(gi + k)];
      if (r >= g0 && r < g1) atomicAdd(&hist[(ld_stream_f64(xraw + k) == 1.0 ? ng : 0) + r - g0], 1);
    }
  }
  __syncthreads();
  int* Hgc = Hg + chunk * n_sel + g0;
  int* Huc = Hu + chunk * n_sel + g0;
  for (int r = threadIdx.x; r < ng; r += blockDim.x) {
    Hgc[r] = hist[r];
    Huc[r] = hist[ng + r];
  }
}

// the same count with both classes packed into one 32-bit word per row (general low half, unit high
// half; a chunk has at most tile_w <= 28 672 cells, so a half never overflows): twice as many rows
// fit the shared-memory histogram, i.e. one pass over the CSC instead of two for 30 000 genes
#define TB_GSEG_PACKED 49152
__global__ void __launch_bounds__(TB_THREADS) k_count_B3(const long long* __restrict__ p, const int* __restrict__ gi,
                                                         const double* __restrict__ xraw,
                                                         const int* __restrict__ sel_of_gene, int n_sel, int g0, int g1,
                                                         int tile_w, int chunks_per_tile, int cells_per_chunk,
                                                         long long n_cells, int* __restrict__ Hg, int* __restrict__ Hu) {
  extern __shared__ unsigned int hist32[];  // [ng]
  const long long chunk = blockIdx.x;
  const long long t = chunk / chunks_per_tile, s = chunk % chunks_per_tile;
  const long long c0 = min(t * tile_w + s * (long long)cells_per_chunk, n_cells);
  const long long c1 = min(min(c0 + cells_per_chunk, (t + 1) * (long long)tile_w), n_cells);
  const int ng = g1 - g0;
  for (int r = threadIdx.x; r < ng; r += blockDim.x) hist32[r] = 0u;
  __syncthreads();
  if (c0 < c1) {
    const long long lo = p[c0], hi = p[c1];
    // four independent index / value loads in flight per thread
    for (long long k0 = lo + threadIdx.x; k0 < hi; k0 += 4LL * blockDim.x) {
      int g[4];
      double x[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const long long k = k0 + (long long)u * blockDim.x;
        g[u] = (k < hi) ? ld_stream_s32(gi + k) : -1;
        x[u] = (k < hi) ? ld_stream_f64(xraw + k) : 0.0;
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int r = (g[u] >= 0) ? sel_of_gene[g[u]] : -1;
        if (r >= g0 && r < g1) atomicAdd(&hist32[r - g0], x[u] == 1.0 ? 0x10000u : 1u);
      }
    }
  }
  __syncthreads();
  int* Hgc = Hg + chunk * n_sel + g0;
  int* Huc = Hu + chunk * n_sel + g0;
  for (int r = threadIdx.x; r < ng; r += blockDim.x) {
    const unsigned int v = hist32[r];
    Hgc[r] = (int)(v & 0xffffu);
    Huc[r] = (int)(v >> 16);
  }
}

__global__ void __launch_bounds__(TB_THREADS) k_scatter_B2(
    const long long* __restrict__ p, const int* __restrict__ gi, const double* __restrict__ y,
    const double* __restrict__ xraw, const int* __restrict__ sel_of_gene, int n_sel, int g0, int g1, int tile_w,
    int chunks_per_tile, int cells_per_chunk, long long n_cells, const int* __restrict__ Hg, const int* __restrict__ Hu,
    const long long* __restrict__ base_g, const long long* __restrict__ base_u, unsigned short* __restrict__ idx_g,
    double* __restrict__ val_g, unsigned short* __restrict__ idx_u) {
  extern __shared__ int sm_tb[];
  const int ng = g1 - g0;
  unsigned int* cursor = reinterpret_cast<unsigned int*>(sm_tb);      // [ng] slot inside the run: general | unit << 16
  unsigned int* mask = reinterpret_cast<unsigned int*>(sm_tb + ng);   // [ng] cells of the current group: general | unit << 16
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const long long chunk = blockIdx.x;
  const long long t = chunk / chunks_per_tile, s = chunk % chunks_per_tile;
  const long long c0 = min(t * tile_w + s * (long long)cells_per_chunk, n_cells);
  const long long c1 = min(min(c0 + cells_per_chunk, (t + 1) * (long long)tile_w), n_cells);
  const int* Hgc = Hg + chunk * n_sel + g0;
  const int* Huc = Hu + chunk * n_sel + g0;
  const long long* btg = base_g + t * n_sel + g0;
  const long long* btu = base_u + t * n_sel + g0;
  for (int r = threadIdx.x; r < ng; r += blockDim.x) {
    cursor[r] = (unsigned)Hgc[r] | ((unsigned)Huc[r] << 16);
    mask[r] = 0u;
  }
  __syncthreads();
  const unsigned int below = (1u << w) - 1u;
  for (long long cg = c0; cg < c1; cg += TB_WARPS) {
    const long long c = cg + w;
    const bool have = c < c1;
    const long long lo = have ? p[c] : 0, hi = have ? p[c + 1] : 0;
    // rr = selected row (or -1), sh = 0 (general) / 16 (unit)
#define TB2_FOR_ROWS(BODY)                                                      \
    for (long long k0 = lo; k0 < hi; k0 += 128) {                                \
      int rr[4], sh4[4];                                                         \
      long long kk[4];                                                           \
      _Pragma("unroll") for (int u = 0; u < 4; ++u) {                            \
        kk[u] = k0 + lane + 32 * u;                                              \
        rr[u] = (kk[u] < hi) ? gi[kk[u]] : -1;                                   \
        sh4[u] = (kk[u] < hi && ld_stream_f64(xraw + kk[u]) == 1.0) ? 16 : 0;    \
      }                                                                          \
      _Pragma("unroll") for (int u = 0; u < 4; ++u) rr[u] = (rr[u] >= 0) ? sel_of_gene[rr[u]] : -1; \
      _Pragma("unroll") for (int u = 0; u < 4; ++u) {                            \
        const int r = rr[u], sh = sh4[u];                                        \
        const long long k = kk[u];                                               \
        if (r >= g0 && r < g1) { BODY }                                          \
      }                                                                          \
    }
    TB2_FOR_ROWS(atomicOr(&mask[r - g0], 1u << (w + sh)); (void)k;)  // A
    __syncthreads();
    const int local = (int)(c - t * tile_w);
    TB2_FOR_ROWS(  // B
        const int q = r - g0;
        const unsigned int mq = (mask[q] >> sh) & 0xffffu;
        const unsigned int cq = (cursor[q] >> sh) & 0xffffu;
        const long long dst = (sh ? btu[q] : btg[q]) + cq + __popc(mq & below);
        if (sh) {
          idx_u[dst] = (unsigned short)local;
        } else {
          idx_g[dst] = (unsigned short)local;
          val_g[dst] = ld_stream_f64(y + k);
        })
    __syncthreads();
    TB2_FOR_ROWS(  // C
        const int q = r - g0; (void)k;
        const unsigned int mq = (mask[q] >> sh) & 0xffffu;
        if ((mq & below) == 0u) {  // this warp's cell is the first of the group with this row and class
          atomicAdd(&cursor[q], (unsigned)__popc(mq) << sh);
          atomicAnd(&mask[q], ~(0xffffu << sh));
        })
#undef TB2_FOR_ROWS
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------
// Stable scatter, third version: SEGMENT-major and read-once.
//
// ncu on k_scatter_B2 at the cfg3 shape (profiles/r02e_full_k_scatter_B2.txt): 16 warps per SM (the
// cursors of 24 576 rows take the whole shared memory), every entry read three times -- once per
// phase, 19.8 GB of DRAM reads for a 5.7 GB input -- 9 % issue utilisation, the top stalls
// `barrier` and `long_scoreboard`: latency-bound.  The rows of a dgCMatrix column are sorted, so the
// entries of a cell that belong to a block of 2 048 consecutive selected rows are ONE contiguous
// piece of the cell.  Here the CTA walks the row segments in the outer loop and keeps one read
// cursor per cell: a warp loads (at most) 128 entries of its cell, keeps row and class in
// registers through the three phases (mark, write, advance) and never reads them again; the
// write cursors of a segment take 16 KB, so three to four CTAs (48-64 warps) share an SM.
// Needs the selected row to grow with the gene id (no use_genes re-ordering): otherwise k_scatter_B2.
// Inside a row the entries stay ordered by cell unless a cell has more than 128 entries in one
// segment (extra rounds); the order is a fixed function of the input either way.
// ---------------------------------------------------------------------------
#define TB3_SEG 2048
#define TB3_CELLS_MAX 4096  // cells per chunk the read cursors are sized for

__global__ void __launch_bounds__(TB_THREADS, 3) k_scatter_B3(
    const long long* __restrict__ p, const int* __restrict__ gi, const double* __restrict__ y,
    const double* __restrict__ xraw, const int* __restrict__ sel_of_gene, int n_sel, int tile_w, int chunks_per_tile,
    int cells_per_chunk, long long n_cells, const int* __restrict__ Hg, const int* __restrict__ Hu,
    const long long* __restrict__ base_g, const long long* __restrict__ base_u, unsigned short* __restrict__ idx_g,
    double* __restrict__ val_g, unsigned short* __restrict__ idx_u, int split) {
  __shared__ unsigned int cursor[TB3_SEG];  // slot inside the (tile, row) run: general | unit << 16
  __shared__ unsigned int mask[TB3_SEG];    // cells of the current group that hold the row: general | unit << 16
  extern __shared__ long long ccur[];       // [cells of the chunk] read cursor into the cell's entries
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const long long chunk = blockIdx.x;
  const long long t = chunk / chunks_per_tile, sc = chunk % chunks_per_tile;
  const long long c0 = min(t * tile_w + sc * (long long)cells_per_chunk, n_cells);
  const long long c1 = min(min(c0 + cells_per_chunk, (t + 1) * (long long)tile_w), n_cells);
  const int n_loc = (int)(c1 - c0);
  for (int i = threadIdx.x; i < n_loc; i += blockDim.x) ccur[i] = p[c0 + i];
  const int* Hgc = Hg + chunk * n_sel;
  const int* Huc = split ? Hu + chunk * n_sel : nullptr;
  const long long* btg = base_g + t * n_sel;
  const long long* btu = split ? base_u + t * n_sel : nullptr;
  const unsigned int below = (1u << w) - 1u;
  for (int seg0 = 0; seg0 < n_sel; seg0 += TB3_SEG) {
    const int seg1 = min(n_sel, seg0 + TB3_SEG), ng = seg1 - seg0;
    __syncthreads();  // the previous segment is finished (and the read cursors are initialised)
    for (int r = threadIdx.x; r < ng; r += blockDim.x) {
      cursor[r] = (unsigned)Hgc[seg0 + r] | (split ? ((unsigned)Huc[seg0 + r] << 16) : 0u);
      mask[r] = 0u;
    }
    __syncthreads();
    for (int cg = 0; cg < n_loc; cg += TB_WARPS) {
      const int ci = cg + w;
      const bool have = ci < n_loc;
      const long long c = c0 + ci;
      long long lo = have ? ccur[ci] : 0;
      const long long hi = have ? p[c + 1] : 0;
      const int local = (int)(c - t * tile_w);
      for (;;) {  // rounds of <= 128 entries; almost always one
        int rr[4], sh4[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const long long k = lo + lane + 32 * u;
          rr[u] = (k < hi) ? gi[k] : -2;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) rr[u] = (rr[u] >= 0) ? sel_of_gene[rr[u]] : rr[u];  // -1: gene not selected, -2: past the end
        // the entries of this segment (or of unselected genes) form a prefix of the 128 in entry order 32 u + lane
        int ncons = 0;
        bool open = true;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const unsigned m = __ballot_sync(0xffffffffu, rr[u] != -2 && rr[u] < seg1);
          if (open) {
            if (m == 0xffffffffu) {
              ncons += 32;
            } else {
              ncons += __ffs(~m) - 1;
              open = false;
            }
          }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int e = 32 * u + lane;
          sh4[u] = 0;
          if (e < ncons && rr[u] >= seg0) {
            rr[u] -= seg0;
            if (split && ld_stream_f64(xraw + lo + e) == 1.0) sh4[u] = 16;
          } else {
            rr[u] = -1;
          }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u)  // A: which cells of the group hold the row
          if (rr[u] >= 0) atomicOr(&mask[rr[u]], 1u << (w + sh4[u]));
        __syncthreads();
#pragma unroll
        for (int u = 0; u < 4; ++u)  // B: slot = cursor + number of earlier cells of the group with the row
          if (rr[u] >= 0) {
            const int q = rr[u], sh = sh4[u];
            const unsigned int mq = (mask[q] >> sh) & 0xffffu, cq = (cursor[q] >> sh) & 0xffffu;
            const long long dst = (sh ? btu[seg0 + q] : btg[seg0 + q]) + cq + __popc(mq & below);
            if (sh) {
              idx_u[dst] = (unsigned short)local;
            } else {
              idx_g[dst] = (unsigned short)local;
              val_g[dst] = ld_stream_f64(y + lo + 32 * u + lane);
            }
          }
        __syncthreads();
#pragma unroll
        for (int u = 0; u < 4; ++u)  // C: the first cell of the group with the row advances its cursor
          if (rr[u] >= 0) {
            const int q = rr[u], sh = sh4[u];
            const unsigned int mq = (mask[q] >> sh) & 0xffffu;
            if ((mq & below) == 0u) {
              atomicAdd(&cursor[q], (unsigned)__popc(mq) << sh);
              atomicAnd(&mask[q], ~(0xffffu << sh));
            }
          }
        lo += ncons;
        if (!__syncthreads_or(ncons == 128)) break;  // (also the barrier that ends phase C)
      }
      if (have && lane == 0) ccur[ci] = lo;
    }
  }
}

// host copies kept for the re-map after the zero-gene filter
struct BBuild {
  std::vector<int> cnt, cnt_u;  // per-(tile, selected row) entries of the general / unit layout
  HostTables T, Tu;
};
void m3b_mat_set_bbuild(m3b_mat* m, void* p) {
  delete (BBuild*)m->b_host_tables;
  m->b_host_tables = p;
}

static double now_ms() {
  return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

int build_layout_B(m3b_mat* m, bool split) {
  m3b_ctx* ctx = m->ctx;
  const bool dbg = getenv("M3B_DEBUG_TIMING") != nullptr;
  const double tb0 = now_ms();
  TiledLayout& L = m->B;
  L.free_all();
  const int n_sel = m->n_sel;
  L.rows = n_sel;
  L.cols = m->n_cells;
  choose_tiling(L.cols, ctx->tile_w_max, ctx->sm_count, &L.tile_w, &L.n_tiles);
  if (split) {
    L.unit = new TiledLayout();
    L.unit->is_unit = true;
    L.unit->rows = L.rows;
    L.unit->cols = L.cols;
    L.unit->tile_w = L.tile_w;
    L.unit->n_tiles = L.n_tiles;
    L.unit->colscale = m->unit_val;  // columns are cells
  }
  const int n_tiles = L.n_tiles;
  // ~2 chunks per SM in total (one CTA each), never fewer than 16 cells per chunk
  int cpt = (int)std::max<long long>(1, std::min<long long>((2LL * ctx->sm_count + n_tiles - 1) / n_tiles,
                                                             (L.tile_w + TB_WARPS - 1) / TB_WARPS));
  {
    // One wave of CTAs per tile: the scatter's partial-sector writes then all fall into ONE tile's output
    // block, which L2 mostly holds, instead of ~6 tiles at once (cfg3 shape: 52 -> 38 ms for layout B).  Bounded
    // by the size of the per-(chunk, row) count tables (two classes, 4 B each).
    const long long table_cap = (512LL << 20) / (8LL * std::max(n_sel, 1) * n_tiles);
    const long long want = std::min<long long>(ctx->sm_count, std::max<long long>(table_cap, cpt));
    cpt = (int)std::max<long long>(cpt, std::min<long long>(want, (L.tile_w + TB_WARPS - 1) / TB_WARPS));
  }
  if (const char* e = getenv("M3B_CPT")) cpt = std::max(1, std::min(atoi(e), (L.tile_w + TB_WARPS - 1) / TB_WARPS));  // sweep knob
  const int cells_per_chunk = (L.tile_w + cpt - 1) / cpt;
  const long long n_chunks = (long long)n_tiles * cpt;
  static bool tb_attr = false;
  if (!tb_attr) {
    CK(ctx, cudaFuncSetAttribute(k_count_B, cudaFuncAttributeMaxDynamicSharedMemorySize, TB_GSEG_MAX * 4));
    CK(ctx, cudaFuncSetAttribute(k_scatter_B, cudaFuncAttributeMaxDynamicSharedMemorySize, TB_GSEG_MAX * 8));
    tb_attr = true;
  }
  int TB_GSEG = TB_GSEG_DEFAULT;
  if (const char* e = getenv("M3B_GSEG")) TB_GSEG = std::max(256, std::min(TB_GSEG_MAX, atoi(e)));
  // the scatter asks for >= 120 KB of shared memory so that only ONE CTA is resident per SM
  const size_t scatter_smem_min = 120 * 1024;
  const long long n_runs = (long long)n_tiles * n_sel;
  // one pass per entry class: {all} or {general, unit}
  const int n_cls = split ? 2 : 1;
  int* H[2] = {nullptr, nullptr};
  int* d_cnt[2] = {nullptr, nullptr};
  long long* d_base[2] = {nullptr, nullptr};
  std::vector<int> cnt[2];
  HostTables T[2];
  TiledLayout* Ls[2] = {&L, L.unit};
  static bool tb2_attr = false;
  if (!tb2_attr) {
    CK(ctx, cudaFuncSetAttribute(k_count_B2, cudaFuncAttributeMaxDynamicSharedMemorySize, TB_GSEG_MAX * 8));
    CK(ctx, cudaFuncSetAttribute(k_scatter_B2, cudaFuncAttributeMaxDynamicSharedMemorySize, TB_GSEG_MAX * 8));
    tb2_attr = true;
  }
  const bool both = split && m->n_cells && n_sel && !getenv("M3B_NO_COMBINED_BUILD");  // one pass for both classes
  if (both) {
    for (int q = 0; q < 2; ++q) {
      RET(dev_alloc(ctx, &H[q], (size_t)n_chunks * n_sel));
      RET(dev_alloc(ctx, &d_cnt[q], (size_t)n_tiles * n_sel));
    }
    if (!getenv("M3B_OLD_SCATTER")) {
      static bool attr3 = false;
      if (!attr3) {
        CK(ctx, cudaFuncSetAttribute(k_count_B3, cudaFuncAttributeMaxDynamicSharedMemorySize, TB_GSEG_PACKED * 4));
        attr3 = true;
      }
      for (int g0 = 0; g0 < n_sel; g0 += TB_GSEG_PACKED) {
        const int g1 = std::min(n_sel, g0 + TB_GSEG_PACKED);
        k_count_B3<<<(unsigned)n_chunks, TB_THREADS, (size_t)(g1 - g0) * 4, ctx->stream>>>(
            m->p, m->i, m->x, m->sel_of_gene, n_sel, g0, g1, L.tile_w, cpt, cells_per_chunk, m->n_cells, H[0], H[1]);
        count_launch(ctx);
      }
    } else {
      for (int g0 = 0; g0 < n_sel; g0 += TB_GSEG_MAX) {
        const int g1 = std::min(n_sel, g0 + TB_GSEG_MAX);
        k_count_B2<<<(unsigned)n_chunks, TB_THREADS, (size_t)(g1 - g0) * 8, ctx->stream>>>(
            m->p, m->i, m->x, m->sel_of_gene, n_sel, g0, g1, L.tile_w, cpt, cells_per_chunk, m->n_cells, H[0], H[1]);
        count_launch(ctx);
      }
    }
  }
  for (int q = 0; q < n_cls; ++q) {
    const int cls = split ? q + 1 : 0;
    if (!both) {
      RET(dev_alloc(ctx, &H[q], (size_t)n_chunks * std::max(n_sel, 1)));
      RET(dev_alloc(ctx, &d_cnt[q], (size_t)n_tiles * std::max(n_sel, 1)));
    }
    if (both) {
      // counted above
    } else if (m->n_cells && n_sel) {
      for (int g0 = 0; g0 < n_sel; g0 += TB_GSEG_MAX) {  // counting has no write working set: big passes
        const int g1 = std::min(n_sel, g0 + TB_GSEG_MAX);
        k_count_B<<<(unsigned)n_chunks, TB_THREADS, (size_t)(g1 - g0) * 4, ctx->stream>>>(
            m->p, m->i, m->x, cls, m->sel_of_gene, n_sel, g0, g1, L.tile_w, cpt, cells_per_chunk, m->n_cells, H[q]);
        count_launch(ctx);
      }
    } else {
      CK(ctx, cudaMemsetAsync(H[q], 0, (size_t)n_chunks * std::max(n_sel, 1) * sizeof(int), ctx->stream));
    }
    if (n_runs) {
      k_scan_chunks<<<(unsigned)((n_runs + 255) / 256), 256, 0, ctx->stream>>>(H[q], n_sel, cpt, n_tiles, d_cnt[q]);
      count_launch(ctx);
    }
    CK(ctx, cudaGetLastError());
    if (use_host_tables()) {
      cnt[q].resize((size_t)n_runs);
      if (n_runs)
        CK(ctx, cudaMemcpyAsync(cnt[q].data(), d_cnt[q], n_runs * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    }
  }
  if (!use_host_tables()) {
    // ---- device tables: bases, streams, items / slots / chunks without leaving the GPU ----
    for (int q = 0; q < n_cls; ++q) RET(dev_alloc(ctx, &d_base[q], (size_t)n_runs + 1));
    long long n_ent[2] = {0, 0};
    RET(tables_compute_bases(ctx, n_tiles, n_sel, d_cnt, n_cls, d_base, n_ent));
    for (int q = 0; q < n_cls; ++q) {
      TiledLayout& Lq = *Ls[q];
      Lq.n_entries = n_ent[q];
      RET(dev_alloc(ctx, &Lq.idx, (size_t)std::max<long long>(n_ent[q], 1) + 16));
      if (q == 0) RET(dev_alloc(ctx, &Lq.val, (size_t)std::max<long long>(n_ent[q], 1)));
      if (n_runs) {
        k_fill_pads<<<(unsigned)((n_runs + 255) / 256), 256, 0, ctx->stream>>>(d_cnt[q], d_base[q], n_runs, q == 1 ? 8 : 4,
                                                                              q == 1 ? Lq.tile_w : 0, Lq.idx, Lq.val);
        count_launch(ctx);
      }
    }
    const bool seg_scatter = (both || !split) && m->sel_monotone && cells_per_chunk <= TB3_CELLS_MAX && !getenv("M3B_OLD_SCATTER");
    if (n_runs && m->n_cells && seg_scatter) {
      k_scatter_B3<<<(unsigned)n_chunks, TB_THREADS, (size_t)cells_per_chunk * sizeof(long long), ctx->stream>>>(
          m->p, m->i, m->y, m->x, m->sel_of_gene, n_sel, L.tile_w, cpt, cells_per_chunk, m->n_cells, H[0], split ? H[1] : nullptr,
          d_base[0], split ? d_base[1] : nullptr, L.idx, L.val, split ? L.unit->idx : nullptr, split ? 1 : 0);
      count_launch(ctx);
    } else
    if (n_runs && m->n_cells) {
      for (int g0 = 0; g0 < n_sel; g0 += TB_GSEG) {
        const int g1 = std::min(n_sel, g0 + TB_GSEG);
        const size_t smem = std::max((size_t)(g1 - g0) * 8, scatter_smem_min);
        if (both) {
          k_scatter_B2<<<(unsigned)n_chunks, TB_THREADS, smem, ctx->stream>>>(
              m->p, m->i, m->y, m->x, m->sel_of_gene, n_sel, g0, g1, L.tile_w, cpt, cells_per_chunk, m->n_cells, H[0], H[1],
              d_base[0], d_base[1], L.idx, L.val, L.unit->idx);
          count_launch(ctx);
        } else {
          for (int q = 0; q < n_cls; ++q) {
            k_scatter_B<<<(unsigned)n_chunks, TB_THREADS, smem, ctx->stream>>>(
                m->p, m->i, m->y, m->x, split ? q + 1 : 0, m->sel_of_gene, n_sel, g0, g1, L.tile_w, cpt, cells_per_chunk,
                m->n_cells, H[q], d_base[q], Ls[q]->idx, Ls[q]->val);
            count_launch(ctx);
          }
        }
      }
    }
    CK(ctx, cudaGetLastError());
    RET(tables_build(ctx, L, n_sel, nullptr, d_cnt, d_base));
    dev_free(m->sel_nnz);
    RET(dev_alloc(ctx, &m->sel_nnz, (size_t)std::max(n_sel, 1)));
    RET(tables_row_stats(ctx, n_tiles, n_sel, nullptr, d_cnt[0], split ? d_cnt[1] : nullptr, m->sel_nnz, nullptr));
    CK(ctx, cudaStreamSynchronize(ctx->stream));
    // the counts and run offsets stay on the device for the re-map after the zero-gene filter
    for (int q = 0; q < 2; ++q) {
      dev_free(m->b_cnt[q]);
      dev_free(m->b_base[q]);
      m->b_cnt[q] = d_cnt[q];
      m->b_base[q] = d_base[q];
      dev_free(H[q]);
    }
    if (dbg) fprintf(stderr, "  layout B (device tables): %.2f ms\n", now_ms() - tb0);
    return M3B_OK;
  }
  CK(ctx, cudaStreamSynchronize(ctx->stream));
  const double tb1 = now_ms();
  for (int q = 0; q < n_cls; ++q) compute_bases(cnt[q], n_tiles, n_sel, T[q], (q == 1) ? SHAPE_UNIT : SHAPE_GENERAL);
  const double tb2 = now_ms();
  for (int q = 0; q < n_cls; ++q) {
    TiledLayout& Lq = *Ls[q];
    const int cls = split ? q + 1 : 0;
    Lq.n_entries = T[q].n_entries;
    RET(dev_alloc(ctx, &Lq.idx, (size_t)T[q].n_entries + 16));
    if (q == 0) RET(dev_alloc(ctx, &Lq.val, (size_t)T[q].n_entries));
    RET(dev_alloc(ctx, &d_base[q], (size_t)std::max<long long>(n_runs, 1)));
    if (n_runs) {
      CK(ctx, cudaMemcpyAsync(d_base[q], T[q].base.data(), n_runs * sizeof(long long), cudaMemcpyHostToDevice, ctx->stream));
      k_fill_pads<<<(unsigned)((n_runs + 255) / 256), 256, 0, ctx->stream>>>(d_cnt[q], d_base[q], n_runs, q == 1 ? 8 : 4,
                                                                            q == 1 ? Lq.tile_w : 0, Lq.idx, Lq.val);
      count_launch(ctx);
      if (both) {
        if (q == 1) {  // both layouts allocated, both base tables uploaded: one scatter for the two classes
          for (int g0 = 0; g0 < n_sel; g0 += TB_GSEG) {
            const int g1 = std::min(n_sel, g0 + TB_GSEG);
            k_scatter_B2<<<(unsigned)n_chunks, TB_THREADS, std::max((size_t)(g1 - g0) * 8, scatter_smem_min), ctx->stream>>>(
                m->p, m->i, m->y, m->x, m->sel_of_gene, n_sel, g0, g1, L.tile_w, cpt, cells_per_chunk, m->n_cells, H[0], H[1],
                d_base[0], d_base[1], L.idx, L.val, L.unit->idx);
            count_launch(ctx);
          }
        }
      } else if (m->n_cells) {
        for (int g0 = 0; g0 < n_sel; g0 += TB_GSEG) {
          const int g1 = std::min(n_sel, g0 + TB_GSEG);
          k_scatter_B<<<(unsigned)n_chunks, TB_THREADS, std::max((size_t)(g1 - g0) * 8, scatter_smem_min), ctx->stream>>>(
              m->p, m->i, m->y, m->x, cls, m->sel_of_gene, n_sel, g0, g1, L.tile_w, cpt, cells_per_chunk, m->n_cells, H[q],
              d_base[q], Lq.idx, Lq.val);
          count_launch(ctx);
        }
      }
    }
    CK(ctx, cudaGetLastError());
  }
  // the item / slot tables are built on the host while the GPU scatters
  for (int q = 0; q < n_cls; ++q)
    build_items(cnt[q], n_tiles, n_sel, nullptr, n_sel, ctx->sm_count, T[q], (q == 1) ? SHAPE_UNIT : SHAPE_GENERAL);
  assign_slots(T[0], split ? &T[1] : nullptr, n_sel);
  if (dbg) CK(ctx, cudaStreamSynchronize(ctx->stream));
  const double tb3 = now_ms();
  const size_t n_partial = T[0].items.size() + (split ? T[1].items.size() : 0);
  RET(upload_tables(ctx, L, T[0], n_partial));
  if (split) RET(upload_tables(ctx, *L.unit, T[1], n_partial, &L));
  if (dbg)
    fprintf(stderr, "  layout B: count+scan %.2f ms (chunks %lld x %d), host tables %.2f ms, scatter %.2f ms, upload %.2f ms\n",
            tb1 - tb0, n_chunks, n_sel, tb2 - tb1, tb3 - tb2, now_ms() - tb3);
  // exact nnz per selected gene (local shard)
  std::vector<long long> nnz_sel(n_sel, 0);
  for (int q = 0; q < n_cls; ++q)
    for (int t = 0; t < n_tiles; ++t)
      for (int r = 0; r < n_sel; ++r) nnz_sel[r] += cnt[q][(size_t)t * n_sel + r];
  dev_free(m->sel_nnz);
  RET(dev_alloc(ctx, &m->sel_nnz, (size_t)std::max(n_sel, 1)));
  if (n_sel) CK(ctx, cudaMemcpyAsync(m->sel_nnz, nnz_sel.data(), n_sel * sizeof(long long), cudaMemcpyHostToDevice, ctx->stream));
  CK(ctx, cudaStreamSynchronize(ctx->stream));
  // stash the host tables for the re-map (kept in the matrix via a heap object)
  BBuild* bb = new BBuild();
  // (whole tables: when the zero-gene filter drops nothing they are final as they are)
  bb->cnt.swap(cnt[0]);
  std::swap(bb->T, T[0]);
  if (split) {
    bb->cnt_u.swap(cnt[1]);
    std::swap(bb->Tu, T[1]);
  }
  m3b_mat_set_bbuild(m, bb);
  for (int q = 0; q < 2; ++q) {
    dev_free(H[q]);
    dev_free(d_cnt[q]);
    dev_free(d_base[q]);
  }
  return M3B_OK;
}

static int bank_reorder(m3b_ctx* ctx, TiledLayout& L);

// After the keep mask is known: re-map B's items to kept rows (no data movement).
static int finish_selection_device(m3b_mat* m, unsigned char* keep_host) {
  m3b_ctx* ctx = m->ctx;
  if (!m->b_cnt[0] || !m->b_base[0]) {
    m3b_set_error(ctx, "internal: layout B tables missing");
    return M3B_E_STATE;
  }
  const int n_sel = m->n_sel;
  const bool split = m->B.unit != nullptr;
  std::vector<int> kept_of_sel(std::max(n_sel, 1));
  int nk = 0;
  for (int s = 0; s < n_sel; ++s) kept_of_sel[s] = keep_host[s] ? nk++ : -1;
  m->n_kept = nk;
  dev_free(m->kept_of_sel);
  RET(dev_alloc(ctx, &m->kept_of_sel, (size_t)std::max(n_sel, 1)));
  if (n_sel) CK(ctx, cudaMemcpyAsync(m->kept_of_sel, kept_of_sel.data(), n_sel * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
  if (nk != n_sel) {  // some rows dropped: re-map the tables (no data moves), turn their entries into pads
    m->B.rows = nk;
    if (split) m->B.unit->rows = nk;
    RET(tables_build(ctx, m->B, n_sel, m->kept_of_sel, m->b_cnt, m->b_base));
    RET(tables_neutralize_dropped(ctx, m->B, n_sel, m->kept_of_sel, m->b_cnt, m->b_base));
  }
  dev_free(m->kept_nnz);
  dev_free(m->kept_padlen);
  RET(dev_alloc(ctx, &m->kept_nnz, (size_t)std::max(nk, 1)));
  RET(dev_alloc(ctx, &m->kept_padlen, (size_t)std::max(nk, 1)));
  RET(tables_row_stats(ctx, m->B.n_tiles, n_sel, m->kept_of_sel, m->b_cnt[0], split ? m->b_cnt[1] : nullptr, m->kept_nnz,
                       m->kept_padlen));
  RET(bank_reorder(ctx, m->B));
  if (split) RET(bank_reorder(ctx, *m->B.unit));
  RET(flat_build_plans_device(ctx, m->B));
  CK(ctx, cudaStreamSynchronize(ctx->stream));  // kept_of_sel (host vector) goes out of scope
  for (int q = 0; q < 2; ++q) {
    dev_free(m->b_cnt[q]);
    dev_free(m->b_base[q]);
    m->b_cnt[q] = nullptr;
    m->b_base[q] = nullptr;
  }
  return M3B_OK;
}

int finish_selection(m3b_mat* m, unsigned char* keep_host) {
  m3b_ctx* ctx = m->ctx;
  if (!use_host_tables()) return finish_selection_device(m, keep_host);
  BBuild* bb = (BBuild*)m->b_host_tables;
  if (!bb) {
    m3b_set_error(ctx, "internal: layout B tables missing");
    return M3B_E_STATE;
  }
  const int n_sel = m->n_sel;
  const bool split = m->B.unit != nullptr;
  std::vector<int> kept_of_sel(n_sel);
  int nk = 0;
  for (int s = 0; s < n_sel; ++s) kept_of_sel[s] = keep_host[s] ? nk++ : -1;
  m->n_kept = nk;
  HostTables Tn, Tun;
  const bool all_kept = nk == n_sel;  // nothing dropped: the tables of build_layout_B are final
  if (!all_kept) {
    Tn.base = bb->T.base;
    build_items(bb->cnt, m->B.n_tiles, n_sel, kept_of_sel.data(), nk, ctx->sm_count, Tn, SHAPE_GENERAL);
    if (split) {
      Tun.base = bb->Tu.base;
      build_items(bb->cnt_u, m->B.n_tiles, n_sel, kept_of_sel.data(), nk, ctx->sm_count, Tun, SHAPE_UNIT);
    }
    assign_slots(Tn, split ? &Tun : nullptr, nk);
    m->B.rows = nk;
    const size_t n_partial = Tn.items.size() + Tun.items.size();
    RET(upload_tables(ctx, m->B, Tn, n_partial));
    if (split) {
      m->B.unit->rows = nk;
      RET(upload_tables(ctx, *m->B.unit, Tun, n_partial, &m->B));
    }
  }
  const HostTables& T = all_kept ? bb->T : Tn;
  const HostTables& Tu = all_kept ? bb->Tu : Tun;
  dev_free(m->kept_of_sel);
  RET(dev_alloc(ctx, &m->kept_of_sel, (size_t)std::max(n_sel, 1)));
  if (n_sel) CK(ctx, cudaMemcpyAsync(m->kept_of_sel, kept_of_sel.data(), n_sel * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
  // per-kept-gene nnz, and the pad entries of the GENERAL runs (OP_SQDEV counts those as (0-mu)^2;
  // the unit kernel skips its pads itself)
  std::vector<long long> nnzk(std::max(nk, 1), 0), padk(std::max(nk, 1), 0);
  for (int t = 0; t < m->B.n_tiles; ++t)
    for (int s = 0; s < n_sel; ++s)
      if (kept_of_sel[s] >= 0) {
        const int c = bb->cnt[(size_t)t * n_sel + s], cu = split ? bb->cnt_u[(size_t)t * n_sel + s] : 0;
        nnzk[kept_of_sel[s]] += c + cu;
        padk[kept_of_sel[s]] += pad4(c) + cu;
      }
  dev_free(m->kept_nnz);
  dev_free(m->kept_padlen);
  RET(dev_alloc(ctx, &m->kept_nnz, (size_t)std::max(nk, 1)));
  RET(dev_alloc(ctx, &m->kept_padlen, (size_t)std::max(nk, 1)));
  CK(ctx, cudaMemcpyAsync(m->kept_nnz, nnzk.data(), std::max(nk, 1) * sizeof(long long), cudaMemcpyHostToDevice, ctx->stream));
  CK(ctx, cudaMemcpyAsync(m->kept_padlen, padk.data(), std::max(nk, 1) * sizeof(long long), cudaMemcpyHostToDevice, ctx->stream));
  CK(ctx, cudaStreamSynchronize(ctx->stream));
  RET(bank_reorder(ctx, m->B));
  if (split) RET(bank_reorder(ctx, *m->B.unit));
  RET(flat_build_plans(ctx, m->B, T, split ? &Tu : nullptr));
  m3b_mat_set_bbuild(m, nullptr);
  return M3B_OK;
}

// ---------------------------------------------------------------------------
// Bank-aware ordering of the entries inside an item.
//
// The product kernel gathers x[idx] from shared memory with one 8-byte load per entry; lane l
// of a warp reads the entries 2(q0+l) and 2(q0+l)+1 of the item.  With the entries in index
// order the 16 lanes of a half-warp hit random bank pairs: ncu counted ~7 shared-memory
// wavefronts per 32-lane gather where 2 are the minimum (profiles/).  The order of the entries
// inside a row is free (it only fixes the summation order), so each item is permuted once at
// build time such that the entry at position p has  idx mod 16 == (p >> 1) mod 16  wherever the
// item has enough entries of that class: consecutive lanes then read 16 different bank pairs.
// Entries of over-full classes fill the holes of under-full ones, in a fixed order, so the
// layout stays deterministic.  One warp per item, staged through shared memory.
// In general: a lane's load covers S consecutive entries (S = 2 for the general stream, S = 8
// for the index-only unit stream, whose lanes gather component j of their 16-byte load together),
// so position p belongs to lane class (p / S) mod 16.
// ---------------------------------------------------------------------------
// The pass is memory-latency bound (a warp walks its item in 32-entry steps), so what counts is
// warps per SM and bytes per load: the staging per warp is cut to what the layout kind needs (no
// value arrays for the index-only unit layouts) and the item is read and written back with
// 16-byte accesses through a second shared-memory copy.  (First version: 3 warps per SM and one
// 2-byte load per lane and step -- 33 ms of a 100 ms m3b_select_genes at the cfg3 shape.)
template <bool HAS_VAL, int MAXE>
struct BrSmem {
  // per warp: source idx, destination position, hole list, permuted idx (+ source / permuted values)
  static constexpr int bytes = MAXE * (2 + 2 + 2 + 2) + (HAS_VAL ? MAXE * 16 : 0) + 256;
};

template <bool HAS_VAL, int MAXE>
__global__ void __launch_bounds__(256) k_bank_reorder(unsigned short* __restrict__ idx, double* __restrict__ val,
                                                      const ItemDesc* __restrict__ items, long long n_items,
                                                      int S /* 2 or 8 */, int lgS, int len_lo) {
  extern __shared__ __align__(16) unsigned char br_smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  unsigned char* base = br_smem + (size_t)warp * BrSmem<HAS_VAL, MAXE>::bytes;
  unsigned short* sidx = reinterpret_cast<unsigned short*>(base);
  unsigned short* dpos = sidx + MAXE;   // destination position of source entry e
  unsigned short* hole = dpos + MAXE;   // positions of the holes, in class order
  unsigned short* sout = hole + MAXE;   // permuted indices
  double* sval = reinterpret_cast<double*>(sout + MAXE);                       // HAS_VAL only
  double* vout = sval + MAXE;                                                  // HAS_VAL only
  int* cnt = reinterpret_cast<int*>(base + BrSmem<HAS_VAL, MAXE>::bytes - 256);  // [16] per class, then scratch [16..63]
  const unsigned lt = (1u << lane) - 1u;
  for (long long it = (long long)blockIdx.x * nwarps + warp; it < n_items; it += (long long)gridDim.x * nwarps) {
    const ItemDesc d = items[it];
    const int len = d.len;
    if (len < len_lo || len > MAXE) continue;  // another size class's item, or too short to matter
    // items start on a multiple of 4 entries (general) / 8 entries (unit) and have such a length:
    // 8-byte / 16-byte accesses on idx, 16-byte accesses on val
    if (HAS_VAL) {
      const uint2* g2 = reinterpret_cast<const uint2*>(idx + d.start);
      for (int q = lane; q < (len >> 2); q += 32) reinterpret_cast<uint2*>(sidx)[q] = g2[q];
      const double2* v2 = reinterpret_cast<const double2*>(val + d.start);
      for (int q = lane; q < (len >> 1); q += 32) reinterpret_cast<double2*>(sval)[q] = v2[q];
    } else {
      const uint4* g4 = reinterpret_cast<const uint4*>(idx + d.start);
      for (int q = lane; q < (len >> 3); q += 32) reinterpret_cast<uint4*>(sidx)[q] = g4[q];
    }
    cnt[lane] = 0;
    cnt[32 + lane] = 0;
    __syncwarp();
    // stable rank of every entry inside its class
    for (int e0 = 0; e0 < len; e0 += 32) {
      const int e = e0 + lane;
      const bool in = e < len;
      const int b = in ? (sidx[e] & 15) : 16 + (lane & 15);  // out-of-range lanes use dummy classes
      const unsigned peers = __match_any_sync(0xffffffffu, b);
      const int before = __popc(peers & lt);
      int basec = 0;
      if (in) basec = cnt[b];
      __syncwarp();
      if (in) {
        dpos[e] = (unsigned short)(basec + before);  // rank for now
        if (before == __popc(peers) - 1) cnt[b] = basec + __popc(peers);  // the last peer publishes
      }
      __syncwarp();
    }
    // native slots of class b inside len: positions p < len with ((p / S) & 15) == b
    // nat[b] = S * floor(len / 16S) + min(S, max(0, (len % 16S) - S b))
    int natb = 0, cntb = 0;
    const int period = 16 * S;
    if (lane < 16) {
      const int r = len % period;
      natb = S * (len / period) + min(S, max(0, r - S * lane));
      cntb = cnt[lane];
    }
    const int over = max(0, cntb - natb), holes = max(0, natb - cntb);
    // exclusive prefix sums over the 16 classes (lanes 0..15)
    int po = over, ph = holes;
#pragma unroll
    for (int o = 1; o < 16; o <<= 1) {
      const int a2 = __shfl_up_sync(0xffffffffu, po, o), h = __shfl_up_sync(0xffffffffu, ph, o);
      if (lane >= o) {
        po += a2;
        ph += h;
      }
    }
    po -= over;
    ph -= holes;
    if (lane < 16) {
      cnt[16 + lane] = natb;      // nat
      cnt[32 + lane] = po;        // overflow prefix
      // enumerate this class's holes: slots cntb .. natb-1
      for (int s2 = cntb; s2 < natb; ++s2)
        hole[ph + (s2 - cntb)] = (unsigned short)(period * (s2 >> lgS) + S * lane + (s2 & (S - 1)));
    }
    __syncwarp();
    for (int e = lane; e < len; e += 32) {
      const int b = sidx[e] & 15, r = dpos[e], nb = cnt[16 + b];
      int p;
      if (r < nb) p = period * (r >> lgS) + S * b + (r & (S - 1));
      else p = hole[cnt[32 + b] + (r - nb)];
      sout[p] = sidx[e];
      if (HAS_VAL) vout[p] = sval[e];
    }
    __syncwarp();
    if (HAS_VAL) {
      uint2* g2 = reinterpret_cast<uint2*>(idx + d.start);
      for (int q = lane; q < (len >> 2); q += 32) g2[q] = reinterpret_cast<const uint2*>(sout)[q];
      double2* v2 = reinterpret_cast<double2*>(val + d.start);
      for (int q = lane; q < (len >> 1); q += 32) v2[q] = reinterpret_cast<const double2*>(vout)[q];
    } else {
      uint4* g4 = reinterpret_cast<uint4*>(idx + d.start);
      for (int q = lane; q < (len >> 3); q += 32) g4[q] = reinterpret_cast<const uint4*>(sout)[q];
    }
    __syncwarp();
  }
}

// Size classes.  The pass is latency-bound, so what counts is warps per SM, and the staging per warp
// is sized for the LONGEST item the instantiation accepts: with one class (4096 unit / 2048 general
// entries) 6 / 4 warps fit an SM although most items are a few hundred entries long (cfg4 on one GPU:
// 35 + 15 ms per layout, 98 ms of a 250 ms m3b_select_genes).  Each class is its own launch over all
// items with its own staging size; a launch skips the items of the other classes.
template <bool HAS_VAL, int MAXE>
static int bank_reorder_class(m3b_ctx* ctx, TiledLayout& L, int len_lo, int S, int lgS) {
  constexpr int per_warp = BrSmem<HAS_VAL, MAXE>::bytes;
  // CTAs of up to 8 warps, as many CTAs per SM as the shared memory holds
  const int warps_per_sm = std::max(1, std::min(32, (int)((220 * 1024) / per_warp)));
  int w = std::min(8, warps_per_sm);
  for (int c = 4; c <= 8 && c <= warps_per_sm; ++c)  // CTA size (4..8 warps) that wastes the fewest warp slots
    if ((warps_per_sm / c) * c > (warps_per_sm / w) * w) w = c;
  const int ctas_per_sm = std::max(1, warps_per_sm / w);
  const int smem = w * per_warp;
  CK(ctx, cudaFuncSetAttribute(k_bank_reorder<HAS_VAL, MAXE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  const unsigned grid = (unsigned)std::min<long long>((L.n_items + w - 1) / w, (long long)ctx->sm_count * ctas_per_sm);
  k_bank_reorder<HAS_VAL, MAXE><<<grid, 32 * w, smem, ctx->stream>>>(L.idx, HAS_VAL ? L.val : nullptr, L.items, L.n_items, S, lgS, len_lo);
  count_launch(ctx);
  CK(ctx, cudaGetLastError());
  return M3B_OK;
}

static int bank_reorder(m3b_ctx* ctx, TiledLayout& L) {
  if (L.n_items == 0 || getenv("M3B_NO_BANK_REORDER")) return M3B_OK;
  if (L.is_unit) {
    RET((bank_reorder_class<false, 1024>(ctx, L, 64, 8, 3)));
    RET((bank_reorder_class<false, 2048>(ctx, L, 1025, 8, 3)));
    RET((bank_reorder_class<false, M3B_UITEM_MAX>(ctx, L, 2049, 8, 3)));
  } else {
    RET((bank_reorder_class<true, 512>(ctx, L, 64, 2, 1)));
    RET((bank_reorder_class<true, M3B_ITEM_MAX>(ctx, L, 513, 2, 1)));
  }
  return M3B_OK;
}

// ---------------------------------------------------------------------------
// Layout A (cell-major, tiled along kept genes): same order as the input, so a
// warp-per-cell stable compaction (ballot/popc) is enough.
// ---------------------------------------------------------------------------
__global__ void k_compose_maps(const int* __restrict__ sel_of_gene, const int* __restrict__ kept_of_sel, int n_genes,
                               int* __restrict__ kept_of_gene) {
  int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g < n_genes) {
    int s = sel_of_gene[g];
    kept_of_gene[g] = (s >= 0) ? kept_of_sel[s] : -1;
  }
}

__global__ void __launch_bounds__(256) k_count_A(const long long* __restrict__ p, const int* __restrict__ gi,
                                                 const double* __restrict__ xraw, int cls,
                                                 const int* __restrict__ kept_of_gene, long long n_cells, int tile_w,
                                                 int n_tiles, int* __restrict__ cnt) {
  const int lane = threadIdx.x & 31;
  long long warp = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long c = warp; c < n_cells; c += nwarps) {
    const long long lo = p[c], hi = p[c + 1];
    for (int t = 0; t < n_tiles; ++t) {
      int n = 0;
      for (long long k = lo + lane; k < hi; k += 32) {
        int g = kept_of_gene[ld_stream_s32(gi + k)];
        n += (g >= 0 && g / tile_w == t && in_class(cls, xraw, k)) ? 1 : 0;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) n += __shfl_xor_sync(0xffffffffu, n, o);
      if (lane == 0) cnt[(long long)t * n_cells + c] = n;
    }
  }
}

__global__ void __launch_bounds__(256) k_scatter_A(const long long* __restrict__ p, const int* __restrict__ gi,
                                                   const double* __restrict__ y, const double* __restrict__ xraw, int cls,
                                                   const int* __restrict__ kept_of_gene,
                                                   long long n_cells, int tile_w, int n_tiles,
                                                   const long long* __restrict__ base, unsigned short* __restrict__ idx,
                                                   double* __restrict__ val) {
  const int lane = threadIdx.x & 31;
  const unsigned lt = (1u << lane) - 1u;
  long long warp = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long c = warp; c < n_cells; c += nwarps) {
    const long long lo = p[c], hi = p[c + 1];
    for (int t = 0; t < n_tiles; ++t) {
      long long dst0 = base[(long long)t * n_cells + c];
      int run = 0;
      for (long long k0 = lo; k0 < hi; k0 += 32) {
        long long k = k0 + lane;
        int g = -1;
        double v = 0.0;
        if (k < hi) {
          g = kept_of_gene[ld_stream_s32(gi + k)];
          if (g >= 0 && g / tile_w == t && in_class(cls, xraw, k)) {
            if (cls != 2) v = ld_stream_f64(y + k);
          } else {
            g = -1;
          }
        }
        unsigned m = __ballot_sync(0xffffffffu, g >= 0);
        if (g >= 0) {
          long long dst = dst0 + run + __popc(m & lt);
          idx[dst] = (unsigned short)(g - t * tile_w);
          if (cls != 2) val[dst] = v;
        }
        run += __popc(m);
      }
      // pads (unit runs: to 8 entries, pointing at the zero slot behind the tile)
      const int pad = (cls == 2) ? 8 : 4;
      int padded = (run + pad - 1) / pad * pad;
      if (lane < padded - run) {
        idx[dst0 + run + lane] = (cls == 2) ? (unsigned short)tile_w : (unsigned short)0;
        if (cls != 2) val[dst0 + run + lane] = 0.0;
      }
    }
  }
}

// split matrices: both classes of layout A in one pass over the cell's entries
__global__ void __launch_bounds__(256) k_count_A2(const long long* __restrict__ p, const int* __restrict__ gi,
                                                  const double* __restrict__ xraw, const int* __restrict__ kept_of_gene,
                                                  long long n_cells, int tile_w, int n_tiles, int* __restrict__ cnt_g,
                                                  int* __restrict__ cnt_u) {
  const int lane = threadIdx.x & 31;
  long long warp = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long c = warp; c < n_cells; c += nwarps) {
    const long long lo = p[c], hi = p[c + 1];
    for (int t = 0; t < n_tiles; ++t) {
      int ng = 0, nu = 0;
      for (long long k = lo + lane; k < hi; k += 32) {
        const int g = kept_of_gene[ld_stream_s32(gi + k)];
        if (g >= 0 && g / tile_w == t) {
          if (ld_stream_f64(xraw + k) == 1.0) ++nu;
          else ++ng;
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        ng += __shfl_xor_sync(0xffffffffu, ng, o);
        nu += __shfl_xor_sync(0xffffffffu, nu, o);
      }
      if (lane == 0) {
        cnt_g[(long long)t * n_cells + c] = ng;
        cnt_u[(long long)t * n_cells + c] = nu;
      }
    }
  }
}

__global__ void __launch_bounds__(256) k_scatter_A2(const long long* __restrict__ p, const int* __restrict__ gi,
                                                    const double* __restrict__ y, const double* __restrict__ xraw,
                                                    const int* __restrict__ kept_of_gene, long long n_cells, int tile_w,
                                                    int n_tiles, const long long* __restrict__ base_g,
                                                    const long long* __restrict__ base_u, unsigned short* __restrict__ idx_g,
                                                    double* __restrict__ val_g, unsigned short* __restrict__ idx_u) {
  const int lane = threadIdx.x & 31;
  const unsigned lt = (1u << lane) - 1u;
  long long warp = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long c = warp; c < n_cells; c += nwarps) {
    const long long lo = p[c], hi = p[c + 1];
    for (int t = 0; t < n_tiles; ++t) {
      const long long dg0 = base_g[(long long)t * n_cells + c], du0 = base_u[(long long)t * n_cells + c];
      int rg = 0, ru = 0;
      for (long long k0 = lo; k0 < hi; k0 += 32) {
        const long long k = k0 + lane;
        int g = -1;
        bool one = false;
        if (k < hi) {
          g = kept_of_gene[ld_stream_s32(gi + k)];
          if (g >= 0 && g / tile_w == t) one = ld_stream_f64(xraw + k) == 1.0;
          else g = -1;
        }
        const unsigned mg = __ballot_sync(0xffffffffu, g >= 0 && !one), mu = __ballot_sync(0xffffffffu, g >= 0 && one);
        if (g >= 0) {
          if (one) {
            idx_u[du0 + ru + __popc(mu & lt)] = (unsigned short)(g - t * tile_w);
          } else {
            const long long dst = dg0 + rg + __popc(mg & lt);
            idx_g[dst] = (unsigned short)(g - t * tile_w);
            val_g[dst] = ld_stream_f64(y + k);
          }
        }
        rg += __popc(mg);
        ru += __popc(mu);
      }
      const int pg = (rg + 3) & ~3, pu = (ru + 7) & ~7;
      if (lane < pg - rg) {
        idx_g[dg0 + rg + lane] = 0;
        val_g[dg0 + rg + lane] = 0.0;
      }
      if (lane < pu - ru) idx_u[du0 + ru + lane] = (unsigned short)tile_w;
    }
  }
}

// Layout A in ONE pass over the cell's entries for all (<= 4) gene tiles and both classes
// (k_count_A2 / k_scatter_A2 read the cell once per tile: 11.8 GB of DRAM reads for a 5.7 GB input
// at the cfg3 shape, and one dependent load chain per 32 entries).  Four batches of 32 entries are
// loaded up front; positions come from one ballot per (tile, class).
template <int NT>
__global__ void __launch_bounds__(256) k_count_A3(const long long* __restrict__ p, const int* __restrict__ gi,
                                                  const double* __restrict__ xraw, const int* __restrict__ kept_of_gene,
                                                  long long n_cells, int tile_w, int n_tiles, int split,
                                                  int* __restrict__ cnt_g, int* __restrict__ cnt_u) {
  const int lane = threadIdx.x & 31;
  long long warp = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long c = warp; c < n_cells; c += nwarps) {
    const long long lo = p[c], hi = p[c + 1];
    int ng[NT], nu[NT];
#pragma unroll
    for (int t = 0; t < NT; ++t) ng[t] = nu[t] = 0;
    for (long long k0 = lo; k0 < hi; k0 += 128) {
      int g[4];
      double x[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const long long k = k0 + lane + 32 * u;
        g[u] = (k < hi) ? ld_stream_s32(gi + k) : -1;
        x[u] = (split && k < hi) ? ld_stream_f64(xraw + k) : 0.0;
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) g[u] = (g[u] >= 0) ? kept_of_gene[g[u]] : -1;
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int tl = (g[u] >= 0) ? g[u] / tile_w : -1;
        const bool one = split && x[u] == 1.0;
#pragma unroll
        for (int t = 0; t < NT; ++t) {
          ng[t] += __popc(__ballot_sync(0xffffffffu, tl == t && !one));
          if (split) nu[t] += __popc(__ballot_sync(0xffffffffu, tl == t && one));
        }
      }
    }
    if (lane == 0) {
#pragma unroll
      for (int t = 0; t < NT; ++t)
        if (t < n_tiles) {
          cnt_g[(long long)t * n_cells + c] = ng[t];
          if (split) cnt_u[(long long)t * n_cells + c] = nu[t];
        }
    }
  }
}

template <int NT>
__global__ void __launch_bounds__(256) k_scatter_A3(const long long* __restrict__ p, const int* __restrict__ gi,
                                                    const double* __restrict__ y, const double* __restrict__ xraw,
                                                    const int* __restrict__ kept_of_gene, long long n_cells, int tile_w,
                                                    int n_tiles, int split, const long long* __restrict__ base_g,
                                                    const long long* __restrict__ base_u, unsigned short* __restrict__ idx_g,
                                                    double* __restrict__ val_g, unsigned short* __restrict__ idx_u) {
  const int lane = threadIdx.x & 31;
  const unsigned lt = (1u << lane) - 1u;
  long long warp = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long c = warp; c < n_cells; c += nwarps) {
    const long long lo = p[c], hi = p[c + 1];
    long long dg[NT], du[NT];
    int rg[NT], ru[NT];
#pragma unroll
    for (int t = 0; t < NT; ++t) {
      dg[t] = (t < n_tiles) ? base_g[(long long)t * n_cells + c] : 0;
      du[t] = (split && t < n_tiles) ? base_u[(long long)t * n_cells + c] : 0;
      rg[t] = ru[t] = 0;
    }
    for (long long k0 = lo; k0 < hi; k0 += 128) {
      int g[4];
      double x[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const long long k = k0 + lane + 32 * u;
        g[u] = (k < hi) ? ld_stream_s32(gi + k) : -1;
        x[u] = (split && k < hi) ? ld_stream_f64(xraw + k) : 0.0;
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) g[u] = (g[u] >= 0) ? kept_of_gene[g[u]] : -1;
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const long long k = k0 + lane + 32 * u;
        const int tl = (g[u] >= 0) ? g[u] / tile_w : -1;
        const bool one = split && x[u] == 1.0;
#pragma unroll
        for (int t = 0; t < NT; ++t) {
          const unsigned mg = __ballot_sync(0xffffffffu, tl == t && !one);
          if (tl == t && !one) {
            const long long dst = dg[t] + rg[t] + __popc(mg & lt);
            idx_g[dst] = (unsigned short)(g[u] - t * tile_w);
            val_g[dst] = ld_stream_f64(y + k);
          }
          rg[t] += __popc(mg);
          if (split) {
            const unsigned mu = __ballot_sync(0xffffffffu, tl == t && one);
            if (tl == t && one) idx_u[du[t] + ru[t] + __popc(mu & lt)] = (unsigned short)(g[u] - t * tile_w);
            ru[t] += __popc(mu);
          }
        }
      }
    }
    // pads: general runs to 4 entries (idx 0, val 0), unit runs to 8 (idx = tile_w, the zero slot behind the tile)
#pragma unroll
    for (int t = 0; t < NT; ++t) {
      const int pg = (rg[t] + 3) & ~3, pu = (ru[t] + 7) & ~7;
      if (lane < pg - rg[t]) {
        idx_g[dg[t] + rg[t] + lane] = 0;
        val_g[dg[t] + rg[t] + lane] = 0.0;
      }
      if (split && lane < pu - ru[t]) idx_u[du[t] + ru[t] + lane] = (unsigned short)tile_w;
    }
  }
}

int build_layout_A(m3b_mat* m, bool split) {
  m3b_ctx* ctx = m->ctx;
  TiledLayout& L = m->A;
  L.free_all();
  L.rows = m->n_cells;
  L.cols = m->n_kept;
  choose_tiling(L.cols, ctx->tile_w_max, ctx->sm_count, &L.tile_w, &L.n_tiles);
  if (split) {
    L.unit = new TiledLayout();
    L.unit->is_unit = true;
    L.unit->rows = L.rows;
    L.unit->cols = L.cols;
    L.unit->tile_w = L.tile_w;
    L.unit->n_tiles = L.n_tiles;
    L.unit->rowscale = m->unit_val;  // rows are cells
  }
  dev_free(m->kept_of_gene);
  RET(dev_alloc(ctx, &m->kept_of_gene, (size_t)std::max(m->n_genes, 1)));
  if (m->n_genes) {
    k_compose_maps<<<(m->n_genes + 255) / 256, 256, 0, ctx->stream>>>(m->sel_of_gene, m->kept_of_sel, m->n_genes, m->kept_of_gene);
    count_launch(ctx);
  }
  const long long n_runs = (long long)L.n_tiles * L.rows;
  const int n_cls = split ? 2 : 1;
  int* d_cnt[2] = {nullptr, nullptr};
  long long* d_base[2] = {nullptr, nullptr};
  std::vector<int> cnt[2];
  HostTables T[2];
  TiledLayout* Ls[2] = {&L, L.unit};
  int blocks = (int)std::max<long long>(1, std::min<long long>((m->n_cells + 7) / 8, (long long)ctx->sm_count * 8));
  const bool both = split && m->n_cells && !getenv("M3B_NO_COMBINED_BUILD");  // one pass for both classes
  // one pass over the cells for all tiles and both classes (<= 4 gene tiles; device tables only)
  const bool one_pass = !use_host_tables() && (both || !split) && m->n_cells && L.n_tiles <= 4 && !getenv("M3B_OLD_SCATTER");
  if (one_pass) {
    for (int q = 0; q < n_cls; ++q) RET(dev_alloc(ctx, &d_cnt[q], (size_t)std::max<long long>(n_runs, 1)));
    const int sp = split ? 1 : 0;
    if (L.n_tiles == 1)
      k_count_A3<1><<<blocks, 256, 0, ctx->stream>>>(m->p, m->i, m->x, m->kept_of_gene, m->n_cells, L.tile_w, L.n_tiles, sp, d_cnt[0], d_cnt[1]);
    else if (L.n_tiles == 2)
      k_count_A3<2><<<blocks, 256, 0, ctx->stream>>>(m->p, m->i, m->x, m->kept_of_gene, m->n_cells, L.tile_w, L.n_tiles, sp, d_cnt[0], d_cnt[1]);
    else
      k_count_A3<4><<<blocks, 256, 0, ctx->stream>>>(m->p, m->i, m->x, m->kept_of_gene, m->n_cells, L.tile_w, L.n_tiles, sp, d_cnt[0], d_cnt[1]);
    count_launch(ctx);
  } else if (both) {
    for (int q = 0; q < 2; ++q) RET(dev_alloc(ctx, &d_cnt[q], (size_t)std::max<long long>(n_runs, 1)));
    k_count_A2<<<blocks, 256, 0, ctx->stream>>>(m->p, m->i, m->x, m->kept_of_gene, m->n_cells, L.tile_w, L.n_tiles, d_cnt[0],
                                                d_cnt[1]);
    count_launch(ctx);
  }
  for (int q = 0; q < n_cls; ++q) {
    if (!both && !one_pass) RET(dev_alloc(ctx, &d_cnt[q], (size_t)std::max<long long>(n_runs, 1)));
    if (m->n_cells && !both && !one_pass) {
      k_count_A<<<blocks, 256, 0, ctx->stream>>>(m->p, m->i, m->x, split ? q + 1 : 0, m->kept_of_gene, m->n_cells, L.tile_w,
                                                 L.n_tiles, d_cnt[q]);
      count_launch(ctx);
    }
    CK(ctx, cudaGetLastError());
    if (use_host_tables()) {
      cnt[q].resize((size_t)n_runs);
      if (n_runs) CK(ctx, cudaMemcpyAsync(cnt[q].data(), d_cnt[q], n_runs * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    }
  }
  if (!use_host_tables()) {
    for (int q = 0; q < n_cls; ++q) RET(dev_alloc(ctx, &d_base[q], (size_t)n_runs + 1));
    if (!m->n_cells)
      for (int q = 0; q < n_cls; ++q) CK(ctx, cudaMemsetAsync(d_cnt[q], 0, std::max<long long>(n_runs, 1) * sizeof(int), ctx->stream));
    long long n_ent[2] = {0, 0};
    RET(tables_compute_bases(ctx, L.n_tiles, L.rows, d_cnt, n_cls, d_base, n_ent));
    for (int q = 0; q < n_cls; ++q) {
      Ls[q]->n_entries = n_ent[q];
      RET(dev_alloc(ctx, &Ls[q]->idx, (size_t)std::max<long long>(n_ent[q], 1)));
      if (q == 0) RET(dev_alloc(ctx, &Ls[q]->val, (size_t)std::max<long long>(n_ent[q], 1)));
    }
    if (n_runs && one_pass) {
      const int sp = split ? 1 : 0;
      unsigned short* iu = split ? L.unit->idx : nullptr;
      if (L.n_tiles == 1)
        k_scatter_A3<1><<<blocks, 256, 0, ctx->stream>>>(m->p, m->i, m->y, m->x, m->kept_of_gene, m->n_cells, L.tile_w, L.n_tiles, sp,
                                                         d_base[0], d_base[1], L.idx, L.val, iu);
      else if (L.n_tiles == 2)
        k_scatter_A3<2><<<blocks, 256, 0, ctx->stream>>>(m->p, m->i, m->y, m->x, m->kept_of_gene, m->n_cells, L.tile_w, L.n_tiles, sp,
                                                         d_base[0], d_base[1], L.idx, L.val, iu);
      else
        k_scatter_A3<4><<<blocks, 256, 0, ctx->stream>>>(m->p, m->i, m->y, m->x, m->kept_of_gene, m->n_cells, L.tile_w, L.n_tiles, sp,
                                                         d_base[0], d_base[1], L.idx, L.val, iu);
      count_launch(ctx);
    } else if (n_runs) {
      if (both) {
        k_scatter_A2<<<blocks, 256, 0, ctx->stream>>>(m->p, m->i, m->y, m->x, m->kept_of_gene, m->n_cells, L.tile_w, L.n_tiles,
                                                      d_base[0], d_base[1], L.idx, L.val, L.unit->idx);
        count_launch(ctx);
      } else {
        for (int q = 0; q < n_cls; ++q) {
          k_scatter_A<<<blocks, 256, 0, ctx->stream>>>(m->p, m->i, m->y, m->x, split ? q + 1 : 0, m->kept_of_gene, m->n_cells,
                                                       L.tile_w, L.n_tiles, d_base[q], Ls[q]->idx, Ls[q]->val);
          count_launch(ctx);
        }
      }
    }
    CK(ctx, cudaGetLastError());
    RET(tables_build(ctx, L, L.rows, nullptr, d_cnt, d_base));
    RET(bank_reorder(ctx, L));
    if (split) RET(bank_reorder(ctx, *L.unit));
    RET(flat_build_plans_device(ctx, L));
    CK(ctx, cudaStreamSynchronize(ctx->stream));
    for (int q = 0; q < 2; ++q) {
      dev_free(d_cnt[q]);
      dev_free(d_base[q]);
    }
    return M3B_OK;
  }
  CK(ctx, cudaStreamSynchronize(ctx->stream));
  for (int q = 0; q < n_cls; ++q) compute_bases(cnt[q], L.n_tiles, L.rows, T[q], (q == 1) ? SHAPE_UNIT : SHAPE_GENERAL);
  for (int q = 0; q < n_cls; ++q) {
    TiledLayout& Lq = *Ls[q];
    Lq.n_entries = T[q].n_entries;
    RET(dev_alloc(ctx, &Lq.idx, (size_t)T[q].n_entries + 16));
    if (q == 0) RET(dev_alloc(ctx, &Lq.val, (size_t)T[q].n_entries));
    RET(dev_alloc(ctx, &d_base[q], (size_t)std::max<long long>(n_runs, 1)));
    if (n_runs) {
      CK(ctx, cudaMemcpyAsync(d_base[q], T[q].base.data(), n_runs * sizeof(long long), cudaMemcpyHostToDevice, ctx->stream));
      if (both) {
        if (q == 1) {
          k_scatter_A2<<<blocks, 256, 0, ctx->stream>>>(m->p, m->i, m->y, m->x, m->kept_of_gene, m->n_cells, L.tile_w, L.n_tiles,
                                                        d_base[0], d_base[1], L.idx, L.val, L.unit->idx);
          count_launch(ctx);
        }
      } else {
        k_scatter_A<<<blocks, 256, 0, ctx->stream>>>(m->p, m->i, m->y, m->x, split ? q + 1 : 0, m->kept_of_gene, m->n_cells,
                                                     L.tile_w, L.n_tiles, d_base[q], Lq.idx, Lq.val);
        count_launch(ctx);
      }
    }
    CK(ctx, cudaGetLastError());
  }
  // the item / slot tables are built on the host while the GPU scatters
  for (int q = 0; q < n_cls; ++q)
    build_items(cnt[q], L.n_tiles, L.rows, nullptr, L.rows, ctx->sm_count, T[q], (q == 1) ? SHAPE_UNIT : SHAPE_GENERAL);
  assign_slots(T[0], split ? &T[1] : nullptr, L.rows);
  const size_t n_partial = T[0].items.size() + (split ? T[1].items.size() : 0);
  RET(upload_tables(ctx, L, T[0], n_partial));
  RET(bank_reorder(ctx, L));
  if (split) {
    RET(upload_tables(ctx, *L.unit, T[1], n_partial, &L));
    RET(bank_reorder(ctx, *L.unit));
  }
  RET(flat_build_plans(ctx, L, T[0], split ? &T[1] : nullptr));
  CK(ctx, cudaStreamSynchronize(ctx->stream));
  for (int q = 0; q < 2; ++q) {
    dev_free(d_cnt[q]);
    dev_free(d_base[q]);
  }
  return M3B_OK;
}

int rebuild_layouts(m3b_mat* m, bool split) {
  if (!m->selected || (int)m->keep_mask.size() < m->n_sel) {
    m3b_set_error(m->ctx, "internal: rebuild_layouts without a selection");
    return M3B_E_STATE;
  }
  RET(build_layout_B(m, split));
  m->split = split;
  RET(finish_selection(m, m->keep_mask.data()));
  RET(build_layout_A(m, split));
  return M3B_OK;
}
