// Device-side construction of the layout tables (items, slots, chunks, row pointers) and of the
// flat work plans -- what m3b_select_genes used to do in host loops (tables_body.h has the story
// and the per-element code; the reference has no counterpart: this is the bookkeeping behind the
// Matrix::t() / FM[use_genes,] materialisations of R/preprocess_cds.R:117-122,139).
//
// Everything is a prefix sum or an independent per-run / per-item / per-row kernel; the only host
// round trips are three scalar read-backs per layout (entry counts -> allocation of the streams,
// item counts -> allocation of the item tables, chunk counts).
#include <algorithm>
#include <vector>

#include "m3b_internal.h"
#include "tables_body.h"

// ---------------------------------------------------------------------------
// exclusive prefix sum: out[0 .. n] (n + 1 entries, out[n] = total), three small kernels
// ---------------------------------------------------------------------------
#define SC_THREADS 256
#define SC_PER 16
#define SC_BLOCK (SC_THREADS * SC_PER)

template <typename Tin, typename Tout>
__global__ void __launch_bounds__(SC_THREADS) k_scan_sums(const Tin* __restrict__ in, long long n, Tout* __restrict__ bsum) {
  __shared__ Tout sh[SC_THREADS / 32];
  const long long b0 = (long long)blockIdx.x * SC_BLOCK;
  Tout s = 0;
  for (int u = 0; u < SC_PER; ++u) {
    const long long k = b0 + (long long)u * SC_THREADS + threadIdx.x;
    if (k < n) s += (Tout)in[k];
  }
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    Tout t = 0;
    for (int w = 0; w < SC_THREADS / 32; ++w) t += sh[w];
    bsum[blockIdx.x] = t;
  }
}

// single block: exclusive scan of the block sums in place; total -> bsum[nb]
template <typename Tout>
__global__ void __launch_bounds__(1024) k_scan_blocks(Tout* __restrict__ bsum, int nb) {
  __shared__ Tout sh[1024];
  __shared__ Tout carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int b0 = 0; b0 < nb; b0 += 1024) {
    const int k = b0 + threadIdx.x;
    const Tout v = (k < nb) ? bsum[k] : (Tout)0;
    sh[threadIdx.x] = v;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
      Tout t = (threadIdx.x >= o) ? sh[threadIdx.x - o] : (Tout)0;
      __syncthreads();
      sh[threadIdx.x] += t;
      __syncthreads();
    }
    if (k < nb) bsum[k] = carry + sh[threadIdx.x] - v;
    __syncthreads();
    if (threadIdx.x == 1023) carry += sh[1023];
    __syncthreads();
  }
  if (threadIdx.x == 0) bsum[nb] = carry;
}

template <typename Tin, typename Tout>
__global__ void __launch_bounds__(SC_THREADS) k_scan_write(const Tin* __restrict__ in, long long n, const Tout* __restrict__ bsum,
                                                           int nb, Tout* __restrict__ out) {
  // thread t owns SC_PER CONSECUTIVE elements: local scan, then a block scan of the thread sums
  __shared__ Tout sh[SC_THREADS];
  const long long b0 = (long long)blockIdx.x * SC_BLOCK + (long long)threadIdx.x * SC_PER;
  Tout v[SC_PER];
  Tout s = 0;
#pragma unroll
  for (int u = 0; u < SC_PER; ++u) {
    const long long k = b0 + u;
    v[u] = (k < n) ? (Tout)in[k] : (Tout)0;
    s += v[u];
  }
  sh[threadIdx.x] = s;
  __syncthreads();
  for (int o = 1; o < SC_THREADS; o <<= 1) {
    Tout t = (threadIdx.x >= o) ? sh[threadIdx.x - o] : (Tout)0;
    __syncthreads();
    sh[threadIdx.x] += t;
    __syncthreads();
  }
  Tout run = bsum[blockIdx.x] + sh[threadIdx.x] - s;
#pragma unroll
  for (int u = 0; u < SC_PER; ++u) {
    const long long k = b0 + u;
    if (k < n) out[k] = run;
    run += v[u];
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) out[n] = bsum[nb];
}

// k_scan_sums sums strided elements of a block, k_scan_write scans consecutive ones: both cover the
// same SC_BLOCK elements per block, so the block sums agree.
template <typename Tin, typename Tout>
static int scan_excl(m3b_ctx* ctx, const Tin* in, long long n, Tout* out) {
  if (n <= 0) {
    CK(ctx, cudaMemsetAsync(out, 0, sizeof(Tout), ctx->stream));
    return M3B_OK;
  }
  const int nb = (int)((n + SC_BLOCK - 1) / SC_BLOCK);
  Tout* bsum = nullptr;
  RET(dev_alloc(ctx, &bsum, (size_t)nb + 1));
  k_scan_sums<Tin, Tout><<<nb, SC_THREADS, 0, ctx->stream>>>(in, n, bsum);
  k_scan_blocks<Tout><<<1, 1024, 0, ctx->stream>>>(bsum, nb);
  k_scan_write<Tin, Tout><<<nb, SC_THREADS, 0, ctx->stream>>>(in, n, bsum, nb, out);
  count_launch(ctx, 3);
  cudaError_t e = cudaGetLastError();
  dev_free(bsum);  // stream-ordered reuse: the next user of the block runs behind these kernels
  CK(ctx, e);
  return M3B_OK;
}

// ---------------------------------------------------------------------------
// kernels: thin wrappers around tables_body.h
// ---------------------------------------------------------------------------
__global__ void k_tb_run_plen(const int* __restrict__ cnt, const int* __restrict__ row_map, long long rows_in, long long n_runs,
                              int pad, long long* __restrict__ plen_all, long long* __restrict__ plen_kept,
                              unsigned long long* __restrict__ nnz_kept) {
  const long long run = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  long long z = 0;
  if (run < n_runs) tb_run_plen(run, cnt, row_map, rows_in, pad, plen_all, plen_kept, &z);
  for (int o = 16; o > 0; o >>= 1) z += __shfl_xor_sync(0xffffffffu, z, o);
  if ((threadIdx.x & 31) == 0 && z) atomicAdd(nnz_kept, (unsigned long long)z);  // integers: order-free
}

__global__ void k_tb_run_items(const int* __restrict__ cnt, const long long* __restrict__ kcum, long long rows_in,
                               long long n_runs, TabShape sh, int* __restrict__ nit) {
  const long long run = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (run < n_runs) nit[run] = tb_run_items(run, cnt, kcum, rows_in, sh);
}

__global__ void k_tb_tile_len(const long long* __restrict__ kcum, long long rows_in, int n_tiles, long long* __restrict__ tlen) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n_tiles) tlen[t] = kcum[(long long)(t + 1) * rows_in] - kcum[(long long)t * rows_in];
}

__global__ void k_tb_row_slots(const int* __restrict__ row_map, long long rows_in, int n_tiles, const int* __restrict__ nit_g,
                               const int* __restrict__ nit_u, int* __restrict__ sbase_g, int* __restrict__ sbase_u,
                               int* __restrict__ rowtot, int* __restrict__ gtot) {
  const long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (r < rows_in) tb_row_slots(r, row_map, rows_in, n_tiles, nit_g, nit_u, sbase_g, sbase_u, rowtot, gtot);
}

__global__ void k_tb_fill_items(const int* __restrict__ cnt, const int* __restrict__ row_map, long long rows_in,
                                long long rows_out, long long n_runs, const long long* __restrict__ base,
                                const long long* __restrict__ kcum, const int* __restrict__ ioff,
                                const int* __restrict__ sbase, const int* __restrict__ rsp, const int* __restrict__ gtot,
                                TabShape sh, TabItemOut o) {
  const long long run = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (run < n_runs) tb_fill_items(run, cnt, row_map, rows_in, rows_out, base, kcum, ioff, sbase, rsp, gtot, sh, o);
}

__global__ void k_tb_chunk_heads(long long n_items, const long long* __restrict__ wcum, const int* __restrict__ item_tile,
                                 const int* __restrict__ row_item_ptr, long long rows_out, int item_max, int* __restrict__ head) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i < n_items) head[i] = tb_chunk_head(i, wcum, item_tile, row_item_ptr, rows_out, item_max);
}

// chunks[c] = {first item, end item}: a head item opens its chunk and closes the previous one
__global__ void k_tb_fill_chunks(long long n_items, const int* __restrict__ head, const int* __restrict__ cidx,
                                 int2* __restrict__ chunks) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n_items) return;
  if (head[i]) {
    const int c = cidx[i];
    chunks[c].x = (int)i;
    if (c > 0) chunks[c - 1].y = (int)i;
  }
  if (i == n_items - 1) chunks[cidx[i] + head[i] - 1].y = (int)n_items;
}

__global__ void k_tb_tile_chunk_ptr(int n_tiles, long long rows_out, const int* __restrict__ row_item_ptr,
                                    const int* __restrict__ cidx, int* __restrict__ tile_chunk_ptr) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n_tiles) tile_chunk_ptr[t] = cidx[row_item_ptr[(long long)t * (rows_out + 1)]];
  if (t == n_tiles) tile_chunk_ptr[t] = cidx[row_item_ptr[(long long)(n_tiles - 1) * (rows_out + 1) + rows_out]];
}

// CTA -> starting tile of the per-item kernels, proportional to the tiles' entry counts
__global__ void k_tb_cta_tile(int n_tiles, int n_cta, const long long* __restrict__ tlen, int* __restrict__ cta_tile) {
  if (blockIdx.x || threadIdx.x) return;
  for (int b = 0; b < n_cta; ++b) cta_tile[b] = 0;
  long long total = 0;
  for (int t = 0; t < n_tiles; ++t) total += tlen[t];
  if (total <= 0) return;
  long long acc = 0;
  int b = 0;
  for (int t = 0; t < n_tiles && b < n_cta; ++t) {
    acc += tlen[t];
    int upto = (int)((double)acc / (double)total * n_cta + 0.5);
    if (tlen[t] > 0 && upto <= b) upto = b + 1;
    upto = min(upto, n_cta);
    for (; b < upto; ++b) cta_tile[b] = t;
  }
  for (; b < n_cta; ++b) cta_tile[b] = n_tiles - 1;
}

// entries of dropped rows (possible only when a row with stored entries has a zero or non-finite
// sum) become pads, so that every tile's stream stays one contiguous, harmless range for the flat
// product kernel, which does not look at item boundaries of rows it was never told about
__global__ void k_tb_neutralize(const int* __restrict__ cnt, const int* __restrict__ row_map, long long rows_in,
                                long long n_runs, const long long* __restrict__ base, int pad, int pad_idx,
                                unsigned short* __restrict__ idx, double* __restrict__ val) {
  const long long run = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (run >= n_runs) return;
  const int c = cnt[run];
  if (c == 0 || row_map[run % rows_in] >= 0) return;
  const long long b = base[run], e = b + tb_padn(c, pad);
  for (long long k = b; k < e; ++k) {
    idx[k] = (unsigned short)pad_idx;
    if (val) val[k] = 0.0;
  }
}

// per-row bookkeeping: exact nnz and padded length (general pads only) over tiles and classes
__global__ void k_tb_row_stats(const int* __restrict__ cnt_g, const int* __restrict__ cnt_u, const int* __restrict__ row_map,
                               long long rows_in, int n_tiles, long long* __restrict__ nnz_out,
                               long long* __restrict__ padlen_out) {
  const long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (r >= rows_in) return;
  if (row_map && row_map[r] < 0) return;
  const long long ro = row_map ? row_map[r] : r;
  long long nz = 0, pl = 0;
  for (int t = 0; t < n_tiles; ++t) {
    const int c = cnt_g[(long long)t * rows_in + r], cu = cnt_u ? cnt_u[(long long)t * rows_in + r] : 0;
    nz += c + cu;
    pl += tb_padn(c, 4) + cu;
  }
  nnz_out[ro] = nz;
  if (padlen_out) padlen_out[ro] = pl;
}

// ---------------------------------------------------------------------------
// driver: tables of one layout (general class + optional unit sibling)
// ---------------------------------------------------------------------------
static const TabShape TB_GENERAL = {4, M3B_ITEM_MAX, M3B_ITEM_TAIL, 2};
static const TabShape TB_UNIT = {8, M3B_UITEM_MAX, M3B_UITEM_TAIL, 8};

// step 1 (only when the storage is laid out here): base = prefix sums of the padded run lengths of
// ALL input rows; returns the entry counts the caller allocates idx / val with.  One sync.
int tables_compute_bases(m3b_ctx* ctx, int n_tiles, long long rows_in, const int* const cnt[2], int n_cls,
                         long long* const base[2], long long n_entries[2]) {
  const long long n_runs = (long long)n_tiles * rows_in;
  long long* plen = nullptr;
  long long* dummy = nullptr;
  unsigned long long* nz = nullptr;
  RET(dev_alloc(ctx, &plen, (size_t)std::max<long long>(n_runs, 1)));
  RET(dev_alloc(ctx, &dummy, (size_t)std::max<long long>(n_runs, 1)));
  RET(dev_alloc(ctx, &nz, (size_t)1));
  int st = M3B_OK;
  long long h[2] = {0, 0};
  for (int q = 0; q < n_cls && st == M3B_OK; ++q) {
    if (n_runs) {
      k_tb_run_plen<<<(unsigned)((n_runs + 255) / 256), 256, 0, ctx->stream>>>(cnt[q], nullptr, rows_in, n_runs,
                                                                            q ? TB_UNIT.pad : TB_GENERAL.pad, plen, dummy, nz);
      count_launch(ctx);
    }
    st = scan_excl<long long, long long>(ctx, plen, n_runs, base[q]);
    if (st == M3B_OK &&
        cudaMemcpyAsync(&h[q], base[q] + n_runs, sizeof(long long), cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess)
      st = M3B_E_CUDA;
  }
  if (st == M3B_OK && cudaStreamSynchronize(ctx->stream) != cudaSuccess) st = M3B_E_CUDA;
  dev_free(plen);
  dev_free(dummy);
  dev_free(nz);
  if (st != M3B_OK) {
    m3b_set_error(ctx, "tables_compute_bases failed: %s", cudaGetErrorString(cudaGetLastError()));
    return st;
  }
  n_entries[0] = h[0];
  n_entries[1] = n_cls > 1 ? h[1] : 0;
  return M3B_OK;
}

static void free_tables(TiledLayout& L) {
  dev_free(L.items);
  dev_free(L.chunks);
  dev_free(L.tile_chunk_ptr);
  dev_free(L.cta_tile);
  dev_free(L.counters);
  dev_free(L.row_item_ptr);
  dev_free(L.item_slot);
  dev_free(L.tile_len);
  dev_free(L.item_tile);
  dev_free(L.item_cum);
  if (!L.is_unit) {
    dev_free(L.row_slot_ptr);
    dev_free(L.partial);
  }
  L.items = nullptr;
  L.chunks = nullptr;
  L.tile_chunk_ptr = nullptr;
  L.cta_tile = nullptr;
  L.counters = nullptr;
  L.row_item_ptr = nullptr;
  L.item_slot = nullptr;
  L.row_slot_ptr = nullptr;
  L.partial = nullptr;
  L.tile_len = nullptr;
  L.item_tile = nullptr;
  L.item_cum = nullptr;
}

#define TBK(call)                                                                                  \
  do {                                                                                             \
    cudaError_t e__ = (call);                                                                      \
    if (e__ != cudaSuccess) {                                                                      \
      m3b_set_error(ctx, "%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__));   \
      st = (e__ == cudaErrorMemoryAllocation) ? M3B_E_NOMEM : M3B_E_CUDA;                          \
      goto out;                                                                                    \
    }                                                                                              \
  } while (0)
#define TBR(expr)                   \
  do {                              \
    st = (expr);                    \
    if (st != M3B_OK) goto out;     \
  } while (0)

// step 2: everything else.  L.rows must already be rows_out; L.n_tiles / tile_w set; L.unit
// non-null iff cnt[1] is.  row_map: device array (input row -> output row or -1) or nullptr.
int tables_build(m3b_ctx* ctx, TiledLayout& L, long long rows_in, const int* row_map, const int* const cnt[2],
                 const long long* const base[2]) {
  const int n_tiles = L.n_tiles;
  const long long rows_out = L.rows;
  const long long n_runs = (long long)n_tiles * rows_in;
  const int n_cls = L.unit ? 2 : 1;
  TiledLayout* Ls[2] = {&L, L.unit};
  const TabShape shp[2] = {TB_GENERAL, TB_UNIT};
  int st = M3B_OK;
  long long *plen = nullptr, *kcum[2] = {nullptr, nullptr}, *wcum = nullptr;
  int *nit[2] = {nullptr, nullptr}, *ioff[2] = {nullptr, nullptr}, *sbase[2] = {nullptr, nullptr};
  int *rowtot = nullptr, *gtot = nullptr, *item_w = nullptr, *head = nullptr, *cidx = nullptr;
  unsigned long long* nz = nullptr;
  long long h_items[2] = {0, 0};
  unsigned long long h_nnz[2] = {0, 0};
  int h_chunks[2] = {0, 0};
  const unsigned rb = (unsigned)std::max<long long>(1, (n_runs + 255) / 256);
  free_tables(L);
  if (L.unit) free_tables(*L.unit);
  TBR(dev_alloc(ctx, &plen, (size_t)std::max<long long>(n_runs, 1)));
  TBR(dev_alloc(ctx, &nz, (size_t)2));
  TBK(cudaMemsetAsync(nz, 0, 2 * sizeof(unsigned long long), ctx->stream));
  for (int q = 0; q < n_cls; ++q) {
    TBR(dev_alloc(ctx, &kcum[q], (size_t)n_runs + 1));
    TBR(dev_alloc(ctx, &nit[q], (size_t)std::max<long long>(n_runs, 1)));
    TBR(dev_alloc(ctx, &ioff[q], (size_t)n_runs + 1));
    TBR(dev_alloc(ctx, &sbase[q], (size_t)std::max<long long>(n_runs, 1)));
    TBR(dev_alloc(ctx, &Ls[q]->tile_len, (size_t)n_tiles));
    if (n_runs) {
      k_tb_run_plen<<<rb, 256, 0, ctx->stream>>>(cnt[q], row_map, rows_in, n_runs, shp[q].pad, nullptr, plen, nz + q);
      count_launch(ctx);
    }
    TBR((scan_excl<long long, long long>(ctx, plen, n_runs, kcum[q])));
    if (n_runs) {
      k_tb_run_items<<<rb, 256, 0, ctx->stream>>>(cnt[q], kcum[q], rows_in, n_runs, shp[q], nit[q]);
      count_launch(ctx);
    }
    TBR((scan_excl<int, int>(ctx, nit[q], n_runs, ioff[q])));
    if (rows_in > 0) {
      k_tb_tile_len<<<(n_tiles + 255) / 256, 256, 0, ctx->stream>>>(kcum[q], rows_in, n_tiles, Ls[q]->tile_len);
      count_launch(ctx);
    } else {
      TBK(cudaMemsetAsync(Ls[q]->tile_len, 0, n_tiles * sizeof(long long), ctx->stream));
    }
    int hi = 0;
    TBK(cudaMemcpyAsync(&hi, ioff[q] + n_runs, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    TBK(cudaStreamSynchronize(ctx->stream));  // (the second class's kernels are short; one sync per class)
    h_items[q] = hi;
  }
  TBK(cudaMemcpyAsync(h_nnz, nz, 2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->stream));
  {
    // slots: per output row, general items then unit items
    const size_t n_partial = (size_t)(h_items[0] + h_items[1]);
    TBR(dev_alloc(ctx, &rowtot, (size_t)std::max<long long>(rows_out, 1)));
    TBR(dev_alloc(ctx, &gtot, (size_t)std::max<long long>(rows_out, 1)));
    TBR(dev_alloc(ctx, &L.row_slot_ptr, (size_t)rows_out + 1));
    TBR(dev_alloc(ctx, &L.partial, std::max<size_t>(n_partial, 1)));
    if (rows_in > 0) {
      k_tb_row_slots<<<(unsigned)((rows_in + 255) / 256), 256, 0, ctx->stream>>>(row_map, rows_in, n_tiles, nit[0], nit[1], sbase[0],
                                                                              sbase[1], rowtot, gtot);
      count_launch(ctx);
    }
    TBR((scan_excl<int, int>(ctx, rowtot, rows_out, L.row_slot_ptr)));
  }
  for (int q = 0; q < n_cls; ++q) {
    TiledLayout& Q = *Ls[q];
    const long long ni = h_items[q];
    Q.n_items = ni;
    Q.n_cta = ctx->sm_count;
    if (q == 1) {
      Q.partial = L.partial;
      Q.row_slot_ptr = L.row_slot_ptr;
    }
    TBR(dev_alloc(ctx, &Q.items, (size_t)std::max<long long>(ni, 1)));
    TBR(dev_alloc(ctx, &Q.item_slot, (size_t)std::max<long long>(ni, 1)));
    TBR(dev_alloc(ctx, &Q.item_cum, (size_t)std::max<long long>(ni, 1)));
    TBR(dev_alloc(ctx, &Q.item_tile, (size_t)std::max<long long>(ni, 1)));
    TBR(dev_alloc(ctx, &Q.row_item_ptr, (size_t)n_tiles * (rows_out + 1)));
    TBR(dev_alloc(ctx, &Q.chunks, (size_t)std::max<long long>(ni, 1)));
    TBR(dev_alloc(ctx, &Q.tile_chunk_ptr, (size_t)n_tiles + 1));
    TBR(dev_alloc(ctx, &Q.cta_tile, (size_t)ctx->sm_count));
    TBR(dev_alloc(ctx, &Q.counters, (size_t)n_tiles + 1));
    TBK(cudaMemsetAsync(Q.counters, 0, ((size_t)n_tiles + 1) * sizeof(unsigned int), ctx->stream));
    dev_free(item_w);
    dev_free(wcum);
    dev_free(head);
    dev_free(cidx);
    item_w = head = cidx = nullptr;
    wcum = nullptr;
    TBR(dev_alloc(ctx, &item_w, (size_t)std::max<long long>(ni, 1)));
    TBR(dev_alloc(ctx, &wcum, (size_t)ni + 1));
    TBR(dev_alloc(ctx, &head, (size_t)std::max<long long>(ni, 1)));
    TBR(dev_alloc(ctx, &cidx, (size_t)ni + 1));
    if (rows_in == 0) TBK(cudaMemsetAsync(Q.row_item_ptr, 0, (size_t)n_tiles * (rows_out + 1) * sizeof(int), ctx->stream));
    if (n_runs) {
      TabItemOut o;
      o.items = Q.items;
      o.item_slot = Q.item_slot;
      o.item_cum = Q.item_cum;
      o.item_tile = Q.item_tile;
      o.item_w = item_w;
      o.row_item_ptr = Q.row_item_ptr;
      k_tb_fill_items<<<rb, 256, 0, ctx->stream>>>(cnt[q], row_map, rows_in, rows_out, n_runs, base[q], kcum[q], ioff[q], sbase[q],
                                                   L.row_slot_ptr, q ? gtot : nullptr, shp[q], o);
      count_launch(ctx);
    }
    // chunks of the per-item kernels
    TBR((scan_excl<int, long long>(ctx, item_w, ni, wcum)));
    if (ni) {
      const unsigned ib = (unsigned)((ni + 255) / 256);
      k_tb_chunk_heads<<<ib, 256, 0, ctx->stream>>>(ni, wcum, Q.item_tile, Q.row_item_ptr, rows_out, shp[q].item_max, head);
      count_launch(ctx);
      TBR((scan_excl<int, int>(ctx, head, ni, cidx)));
      k_tb_fill_chunks<<<ib, 256, 0, ctx->stream>>>(ni, head, cidx, Q.chunks);
      k_tb_tile_chunk_ptr<<<(n_tiles + 1 + 255) / 256, 256, 0, ctx->stream>>>(n_tiles, rows_out, Q.row_item_ptr, cidx, Q.tile_chunk_ptr);
      count_launch(ctx, 2);
      TBK(cudaMemcpyAsync(&h_chunks[q], cidx + ni, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    } else {
      TBK(cudaMemsetAsync(Q.tile_chunk_ptr, 0, ((size_t)n_tiles + 1) * sizeof(int), ctx->stream));
    }
    k_tb_cta_tile<<<1, 32, 0, ctx->stream>>>(n_tiles, ctx->sm_count, Q.tile_len, Q.cta_tile);
    count_launch(ctx);
    TBK(cudaGetLastError());
  }
  TBK(cudaStreamSynchronize(ctx->stream));
  for (int q = 0; q < n_cls; ++q) {
    Ls[q]->n_chunks = h_chunks[q];
    Ls[q]->nnz = (long long)h_nnz[q];
  }
out:
  dev_free(plen);
  dev_free(nz);
  for (int q = 0; q < 2; ++q) {
    dev_free(kcum[q]);
    dev_free(nit[q]);
    dev_free(ioff[q]);
    dev_free(sbase[q]);
  }
  dev_free(rowtot);
  dev_free(gtot);
  dev_free(item_w);
  dev_free(wcum);
  dev_free(head);
  dev_free(cidx);
  return st;
}

int tables_neutralize_dropped(m3b_ctx* ctx, TiledLayout& L, long long rows_in, const int* row_map, const int* const cnt[2],
                              const long long* const base[2]) {
  const long long n_runs = (long long)L.n_tiles * rows_in;
  if (!row_map || n_runs == 0) return M3B_OK;
  const unsigned rb = (unsigned)((n_runs + 255) / 256);
  k_tb_neutralize<<<rb, 256, 0, ctx->stream>>>(cnt[0], row_map, rows_in, n_runs, base[0], 4, 0, L.idx, L.val);
  count_launch(ctx);
  if (L.unit) {
    k_tb_neutralize<<<rb, 256, 0, ctx->stream>>>(cnt[1], row_map, rows_in, n_runs, base[1], 8, L.tile_w, L.unit->idx, nullptr);
    count_launch(ctx);
  }
  CK(ctx, cudaGetLastError());
  return M3B_OK;
}

int tables_row_stats(m3b_ctx* ctx, int n_tiles, long long rows_in, const int* row_map, const int* cnt_g, const int* cnt_u,
                     long long* nnz_out, long long* padlen_out) {
  if (rows_in == 0) return M3B_OK;
  k_tb_row_stats<<<(unsigned)((rows_in + 255) / 256), 256, 0, ctx->stream>>>(cnt_g, cnt_u, row_map, rows_in, n_tiles, nnz_out,
                                                                          padlen_out);
  count_launch(ctx);
  CK(ctx, cudaGetLastError());
  return M3B_OK;
}

// ---------------------------------------------------------------------------
// flat plans on the device
// ---------------------------------------------------------------------------
__global__ void k_tb_plan_meta(int n_tiles, int n_cta, int warps, const long long* __restrict__ tlen_g,
                               const long long* __restrict__ tlen_u, long long target_g, long long target_u, PlanTile* pt,
                               int* cta_visit_ptr, int* visit_tile, int* gp_ptr, int* up_ptr, int* counts, int* n_of, int* order,
                               double* load, int* cta_of) {
  if (blockIdx.x || threadIdx.x) return;
  tb_plan_meta(n_tiles, n_cta, warps, tlen_g, tlen_u, target_g, target_u, pt, cta_visit_ptr, visit_tile, gp_ptr, up_ptr, counts,
               n_of, order, load, cta_of);
}

__global__ void k_tb_plan_pieces(long long n_items, int cls, const ItemDesc* __restrict__ items,
                                 const long long* __restrict__ item_cum, const int* __restrict__ item_tile,
                                 const int* __restrict__ row_item_ptr, long long rows_out, const long long* __restrict__ tlen,
                                 const PlanTile* __restrict__ pt, int warps, const int* __restrict__ ptr, int group,
                                 int4* __restrict__ slots) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i < n_items) tb_plan_piece(i, cls, items, item_cum, item_tile, row_item_ptr, rows_out, tlen, pt, warps, ptr, group, slots);
}

// .y = groups of the piece (0 for a slot no item fell into)
__global__ void k_tb_plan_fix(long long n_slots, int4* __restrict__ slots) {
  const long long j = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (j < n_slots) {
    int4 s = slots[j];
    s.y = (s.w > s.x) ? s.w - s.x : 0;
    slots[j] = s;
  }
}

int tables_build_plan(m3b_ctx* ctx, const TiledLayout& L, int n_cta, int warps, long long target_g, long long target_u,
                      int** cta_visit_ptr, int** visit_tile, int** gp_ptr, int4** gp, int** up_ptr, int4** up, int* n_visits,
                      long long* n_pieces) {
  const TiledLayout* U = L.unit;
  const int n_tiles = L.n_tiles;
  const int max_visits = std::max(n_cta, n_tiles);
  // slot bounds: sum over tiles of W_t * ceil(total_t / (W_t target)) <= entries / target + sum W_t
  const long long gcap = L.n_entries / target_g + (long long)max_visits * warps + 1;
  const long long ucap = U ? U->n_entries / target_u + (long long)max_visits * warps + 1 : 1;
  int st = M3B_OK;
  PlanTile* pt = nullptr;
  int *counts = nullptr, *n_of = nullptr, *order = nullptr, *cta_of = nullptr;
  double* load = nullptr;
  int h_counts[3] = {0, 0, 0};
  TBR(dev_alloc(ctx, cta_visit_ptr, (size_t)n_cta + 1));
  TBR(dev_alloc(ctx, visit_tile, (size_t)max_visits));
  TBR(dev_alloc(ctx, gp_ptr, (size_t)max_visits * warps + 1));
  TBR(dev_alloc(ctx, up_ptr, (size_t)max_visits * warps + 1));
  TBR(dev_alloc(ctx, gp, (size_t)gcap));
  TBR(dev_alloc(ctx, up, (size_t)ucap));
  {
    int* raw = nullptr;
    TBR(dev_alloc(ctx, &raw, (size_t)n_tiles * (sizeof(PlanTile) / sizeof(int))));
    pt = reinterpret_cast<PlanTile*>(raw);
  }
  TBR(dev_alloc(ctx, &counts, (size_t)4));
  TBR(dev_alloc(ctx, &n_of, (size_t)n_tiles));
  TBR(dev_alloc(ctx, &order, (size_t)n_tiles));
  TBR(dev_alloc(ctx, &cta_of, (size_t)n_tiles));
  TBR(dev_alloc(ctx, &load, (size_t)n_cta));
  TBK(cudaMemsetAsync(*gp, 0, (size_t)gcap * sizeof(int4), ctx->stream));
  TBK(cudaMemsetAsync(*up, 0, (size_t)ucap * sizeof(int4), ctx->stream));
  k_tb_plan_meta<<<1, 32, 0, ctx->stream>>>(n_tiles, n_cta, warps, L.tile_len, U ? U->tile_len : nullptr, target_g, target_u, pt,
                                            *cta_visit_ptr, *visit_tile, *gp_ptr, *up_ptr, counts, n_of, order, load, cta_of);
  count_launch(ctx);
  if (L.n_items) {
    k_tb_plan_pieces<<<(unsigned)((L.n_items + 255) / 256), 256, 0, ctx->stream>>>(L.n_items, 0, L.items, L.item_cum, L.item_tile,
                                                                                L.row_item_ptr, L.rows, L.tile_len, pt, warps,
                                                                                *gp_ptr, 2, *gp);
    k_tb_plan_fix<<<(unsigned)((gcap + 255) / 256), 256, 0, ctx->stream>>>(gcap, *gp);
    count_launch(ctx, 2);
  }
  if (U && U->n_items) {
    k_tb_plan_pieces<<<(unsigned)((U->n_items + 255) / 256), 256, 0, ctx->stream>>>(U->n_items, 1, U->items, U->item_cum,
                                                                                 U->item_tile, U->row_item_ptr, U->rows,
                                                                                 U->tile_len, pt, warps, *up_ptr, 8, *up);
    k_tb_plan_fix<<<(unsigned)((ucap + 255) / 256), 256, 0, ctx->stream>>>(ucap, *up);
    count_launch(ctx, 2);
  }
  TBK(cudaGetLastError());
  TBK(cudaMemcpyAsync(h_counts, counts, 3 * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  TBK(cudaStreamSynchronize(ctx->stream));
  *n_visits = h_counts[0];
  *n_pieces = (long long)h_counts[1] + h_counts[2];
  if (h_counts[1] > gcap || h_counts[2] > ucap) {
    m3b_set_error(ctx, "internal: flat plan needs %d + %d piece slots, bound was %lld + %lld", h_counts[1], h_counts[2], gcap, ucap);
    st = M3B_E_STATE;
  }
out:
  dev_free(pt);
  dev_free(counts);
  dev_free(n_of);
  dev_free(order);
  dev_free(cta_of);
  dev_free(load);
  return st;
}
