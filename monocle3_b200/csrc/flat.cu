// K4 / K5, the product kernel proper (OP_DOT): flat streaming with a static plan.
//
// Replaces CHOLMOD's cholmod_sdmult inside irlba (the two products per Lanczos step behind
// R/utils.R:417).  The per-item kernels of spmv.cu pay one exposed DRAM latency per item, which
// was invisible while items were ~1000 entries of a 10 B/entry stream but dominates once the
// count-1 entries live in the 2 B/entry unit stream (items of a few hundred bytes; ncu: 1.7 TB/s,
// long-scoreboard stalls).  Here the stream of a tile is cut ONCE, at build time, into
// item-aligned, memory-contiguous pieces of equal length for the warps of the CTAs that own the
// tile (at cfg2 sizes: one piece per warp), and a warp streams its pieces flat:
//   * two register batches of K rounds (32 lanes x one group each; K = 3 by default) keep a few KB
//     per warp in flight regardless of where items begin and end;
//   * bit 15 of the first index of a group marks "this group starts a new item" (tile-local
//     indices are < 28672, so the bit is free): one ballot per round tells the warp whether an
//     item boundary is inside the round.  No boundary: lanes just accumulate.  Boundary: the open
//     item is closed with a shuffle tree, items that lie completely inside the round with a
//     segmented scan -- all in a fixed order, so results stay bit-reproducible;
//   * the item's partial goes to its slot (spmv.cu's combine kernels are unchanged).
// The general stream (idx + val, 10 B/entry) and the unit stream (idx only, 2 B/entry) of a tile
// are processed by the same CTA from the same staged tile; on layout B the tile is rescaled by
// the per-cell unit value between the two phases.
#include <algorithm>
#include <vector>

#include "m3b_internal.h"

// Kernel variants <rounds per register batch, threads per CTA>; one CTA per SM either way (the tile
// takes the shared memory), so the thread count trades registers per thread (batch depth) against
// warps per scheduler (latency hiding).  M3B_FLAT_VARIANT selects one for sweeps.  Measured on cfg2
// (us per launch): <6,512> 78.7, <3,1024> 65.1 (default), <2,1024> 64.9; rotating three batches
// instead of two (<2,1024,3>, <1,1024,3>) changed nothing or lost: the kernel is no longer waiting
// for DRAM but spread thin over issue slots, shared-memory gathers and load latency (profiles/).
struct FlatVariant {
  int k, threads;
};
static const FlatVariant FL_VARIANTS[] = {{6, 512}, {3, 1024}, {2, 1024}, {1, 1024}, {2, 1024}};  // 3, 4: ring, R = k
static int fl_variant() {
  static int v = -1;
  static const bool dynamic = getenv("M3B_FLAT_DYNAMIC") != nullptr;  // sweeps: re-read the variant on every call
  if (v < 0 || dynamic) {
    const char* e = getenv("M3B_FLAT_VARIANT");
    v = e ? atoi(e) : 1;
    if (v < 0 || v >= (int)(sizeof(FL_VARIANTS) / sizeof(FL_VARIANTS[0]))) v = 0;
  }
  return v;
}
#define FL_WARPS (FL_VARIANTS[fl_variant()].threads / 32)

struct FlatPlan {
  int n_cta = 0, n_visits = 0;
  long long n_pieces = 0;        // descriptors in gp + up
  int* cta_visit_ptr = nullptr;  // [n_cta+1]
  int* visit_tile = nullptr;     // [n_visits]
  int* gp_ptr = nullptr;         // [n_visits*FL_WARPS+1] piece ranges of (visit, warp), general stream
  int4* gp = nullptr;            // {first group (pairs), groups, first item, -}
  int* up_ptr = nullptr;         // the same for the unit stream (groups of 8 entries)
  int4* up = nullptr;
};

long long flat_plan_pieces(const FlatPlan* p) { return p ? p->n_pieces : 0; }
bool flat_has_plan(const TiledLayout& L, int n_cta) { return L.plan[1] && L.plan[1]->n_cta == n_cta; }

void flat_plan_free(FlatPlan* p) {
  if (!p) return;
  dev_free(p->cta_visit_ptr);
  dev_free(p->visit_tile);
  dev_free(p->gp_ptr);
  dev_free(p->gp);
  dev_free(p->up_ptr);
  dev_free(p->up);
  delete p;
}

// ---------------------------------------------------------------------------
// host: the plan
// ---------------------------------------------------------------------------
// items [i0, i1) of one tile -> pieces for W warps.  The tile's stream is cut into item-aligned
// pieces of ~target entries and dealt round-robin: at any time the W warps work inside one
// window of W consecutive pieces that moves through the stream.  (Giving every warp ONE long
// contiguous share instead spread the 2368 concurrent read streams over the whole 700 MB array;
// ncu then showed 4 us of load latency -- TLB and DRAM-page misses -- and 2.7 TB/s.)
// A piece is contiguous in memory (holes exist only where dropped rows had entries).
static void split_tile(const std::vector<ItemDesc>& items, int i0, int i1, int W, int group, long long target,
                       std::vector<std::vector<int4>>& out) {
  out.assign(W, std::vector<int4>());
  long long total = 0;
  for (int i = i0; i < i1; ++i) total += items[i].len;
  if (total == 0) return;
  const long long segs = std::max<long long>(1, (total + (long long)W * target - 1) / ((long long)W * target));
  const long long P = segs * W;
  long long cum = 0, cur_p = -1;
  long long prev_end = -1;
  int w = 0;
  for (int i = i0; i < i1; ++i) {
    const ItemDesc& d = items[i];
    long long pc = std::min<long long>(P - 1, (long long)(((double)cum + 0.5 * d.len) * (double)P / (double)total));
    if (pc < cur_p) pc = cur_p;
    if (pc != cur_p || d.start != prev_end) {
      w = (int)(pc % W);
      out[w].push_back(make_int4((int)(d.start / group), 0, i, 0));
      cur_p = pc;
    }
    out[w].back().y += d.len / group;
    prev_end = d.start + d.len;
    cum += d.len;
  }
}

template <typename T>
static int upload_vec(m3b_ctx* ctx, T** dst, const std::vector<T>& v) {
  RET(dev_alloc(ctx, dst, std::max<size_t>(v.size(), 1)));
  if (!v.empty()) CK(ctx, cudaMemcpyAsync(*dst, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, ctx->stream));
  return M3B_OK;
}
static long long env_ll(const char* name, long long dflt) {
  const char* e = getenv(name);
  return e ? std::max(64LL, atoll(e)) : dflt;
}
// entries per piece (M3B_FLAT_PIECE_G / M3B_FLAT_PIECE_U override, for sweeps)
static long long fl_piece_general() { return env_ll("M3B_FLAT_PIECE_G", 8192); }
static long long fl_piece_unit() { return env_ll("M3B_FLAT_PIECE_U", 32768); }

// the plan as host vectors (pure host code: also what tools/host_tables_check.cu compares the
// device builder with)
struct HostPlan {
  std::vector<int> cta_visit_ptr, visit_tile, gp_ptr, up_ptr;
  std::vector<int4> gp, up;
};
static void plan_host(int n_tiles, long long rows, const HostTables& Tg, const HostTables* Tu, int n_cta, int warps,
                      long long target_g, long long target_u, HostPlan& HP) {
  auto tile_items = [&](const HostTables& T, int t, int* i0, int* i1) {
    *i0 = T.row_item_ptr[(size_t)t * (rows + 1)];
    *i1 = T.row_item_ptr[(size_t)t * (rows + 1) + rows];
  };
  std::vector<double> bytes(n_tiles, 0.0);
  for (int t = 0; t < n_tiles; ++t) {
    int i0, i1;
    tile_items(Tg, t, &i0, &i1);
    for (int i = i0; i < i1; ++i) bytes[t] += 10.0 * Tg.items[i].len;
    if (Tu) {
      tile_items(*Tu, t, &i0, &i1);
      for (int i = i0; i < i1; ++i) bytes[t] += 2.0 * Tu->items[i].len;
    }
  }
  double total = 0;
  for (double b : bytes) total += b;
  // tiles -> CTAs
  std::vector<std::vector<int>> cta_tiles(n_cta);  // visits of every CTA
  std::vector<std::vector<int>> tile_ctas(n_tiles);
  if (total > 0) {
    if (n_tiles <= n_cta) {
      // proportional to the tiles' bytes, every non-empty tile gets at least one CTA
      std::vector<int> n_of(n_tiles, 0);
      int used = 0, nonempty = 0;
      for (int t = 0; t < n_tiles; ++t) nonempty += bytes[t] > 0;
      for (int t = 0; t < n_tiles; ++t)
        if (bytes[t] > 0) {
          n_of[t] = std::max(1, (int)(bytes[t] / total * (n_cta - nonempty)) + 1);
          used += n_of[t];
        }
      // hand the remaining CTAs to the tiles with the most bytes per CTA
      while (used < n_cta) {
        int best = -1;
        for (int t = 0; t < n_tiles; ++t)
          if (bytes[t] > 0 && (best < 0 || bytes[t] / n_of[t] > bytes[best] / n_of[best])) best = t;
        n_of[best]++;
        used++;
      }
      while (used > n_cta) {
        int best = -1;
        for (int t = 0; t < n_tiles; ++t)
          if (n_of[t] > 1 && (best < 0 || bytes[t] / n_of[t] < bytes[best] / n_of[best])) best = t;
        n_of[best]--;
        used--;
      }
      int c = 0;
      for (int t = 0; t < n_tiles; ++t)
        for (int k = 0; k < n_of[t]; ++k, ++c) {
          cta_tiles[c].push_back(t);
          tile_ctas[t].push_back(c);
        }
    } else {
      // more tiles than CTAs: greedy, heaviest tile to the least loaded CTA
      std::vector<int> order(n_tiles);
      for (int t = 0; t < n_tiles; ++t) order[t] = t;
      std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return bytes[a] > bytes[b]; });
      std::vector<double> load(n_cta, 0.0);
      for (int t : order) {
        if (bytes[t] <= 0) continue;
        int c = (int)(std::min_element(load.begin(), load.end()) - load.begin());
        load[c] += bytes[t];
        cta_tiles[c].push_back(t);
        tile_ctas[t].push_back(c);
      }
    }
  }
  // visits, in CTA order
  std::vector<int> cta_visit_ptr(n_cta + 1, 0), visit_tile;
  std::vector<std::vector<int>> visit_of(n_cta);
  for (int c = 0; c < n_cta; ++c) {
    cta_visit_ptr[c] = (int)visit_tile.size();
    for (int t : cta_tiles[c]) {
      visit_of[c].push_back((int)visit_tile.size());
      visit_tile.push_back(t);
    }
  }
  cta_visit_ptr[n_cta] = (int)visit_tile.size();
  const int n_visits = (int)visit_tile.size();
  // shares of every (visit, warp)
  std::vector<std::vector<int4>> gshare((size_t)n_visits * warps), ushare((size_t)n_visits * warps);
  for (int t = 0; t < n_tiles; ++t) {
    const int nc = (int)tile_ctas[t].size();
    if (nc == 0) continue;
    const int W = nc * warps;
    for (int pass = 0; pass < (Tu ? 2 : 1); ++pass) {
      const HostTables& T = pass ? *Tu : Tg;
      int i0, i1;
      tile_items(T, t, &i0, &i1);
      std::vector<std::vector<int4>> sh;
      split_tile(T.items, i0, i1, W, pass ? 8 : 2, pass ? target_u : target_g, sh);
      for (int k = 0; k < nc; ++k) {
        const int c = tile_ctas[t][k];
        // the visit of CTA c that is tile t
        int v = -1;
        for (size_t q = 0; q < cta_tiles[c].size(); ++q)
          if (cta_tiles[c][q] == t) v = visit_of[c][q];
        for (int w = 0; w < warps; ++w) {
          if (sh.empty()) continue;
          (pass ? ushare : gshare)[(size_t)v * warps + w] = sh[(size_t)k * warps + w];
        }
      }
    }
  }
  std::vector<int>& gp_ptr = HP.gp_ptr;
  std::vector<int>& up_ptr = HP.up_ptr;
  std::vector<int4>& gp = HP.gp;
  std::vector<int4>& up = HP.up;
  gp_ptr.clear();
  up_ptr.clear();
  gp.clear();
  up.clear();
  for (size_t k = 0; k < gshare.size(); ++k) {
    gp_ptr.push_back((int)gp.size());
    gp.insert(gp.end(), gshare[k].begin(), gshare[k].end());
    up_ptr.push_back((int)up.size());
    up.insert(up.end(), ushare[k].begin(), ushare[k].end());
  }
  gp_ptr.push_back((int)gp.size());
  up_ptr.push_back((int)up.size());
  HP.cta_visit_ptr = cta_visit_ptr;
  HP.visit_tile = visit_tile;
}

static int build_one_plan(m3b_ctx* ctx, const TiledLayout& L, const HostTables& Tg, const HostTables* Tu, int n_cta,
                          FlatPlan** out) {
  HostPlan HP;
  plan_host(L.n_tiles, L.rows, Tg, Tu, n_cta, FL_WARPS, fl_piece_general(), fl_piece_unit(), HP);
  FlatPlan* P = new FlatPlan();
  *out = P;
  P->n_cta = n_cta;
  P->n_visits = (int)HP.visit_tile.size();
  P->n_pieces = (long long)HP.gp.size() + (long long)HP.up.size();
  RET(upload_vec(ctx, &P->cta_visit_ptr, HP.cta_visit_ptr));
  RET(upload_vec(ctx, &P->visit_tile, HP.visit_tile));
  RET(upload_vec(ctx, &P->gp_ptr, HP.gp_ptr));
  RET(upload_vec(ctx, &P->gp, HP.gp));
  RET(upload_vec(ctx, &P->up_ptr, HP.up_ptr));
  RET(upload_vec(ctx, &P->up, HP.up));
  CK(ctx, cudaStreamSynchronize(ctx->stream));  // the host vectors go out of scope
  return M3B_OK;
}

__global__ void k_mark_heads(unsigned short* __restrict__ idx, const ItemDesc* __restrict__ items, long long n_items) {
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i < n_items) idx[items[i].start] |= (unsigned short)0x8000;
}

__global__ void k_item_scale(const ItemDesc* __restrict__ items, long long n_items, const double* __restrict__ rowscale,
                             double* __restrict__ out) {
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i < n_items) out[i] = rowscale[items[i].row];
}

// Called when the items of a layout (and of its unit sibling) are final and bank-ordered.
int flat_build_plans(m3b_ctx* ctx, TiledLayout& L, const HostTables& Tg, const HostTables* Tu) {
  for (int k = 0; k < 2; ++k) {
    flat_plan_free(L.plan[k]);
    L.plan[k] = nullptr;
  }
  if (getenv("M3B_NO_FLAT")) return M3B_OK;
  if (L.n_entries / 2 >= (1LL << 31) || (L.unit && L.unit->n_entries / 8 >= (1LL << 31))) return M3B_OK;  // 32-bit group offsets
  RET(build_one_plan(ctx, L, Tg, Tu, ctx->sm_count, &L.plan[0]));
  if (L.n_tiles == 1 && ctx->sm_count > 2 * M3B_SIDE_CTA_RESERVE)
    RET(build_one_plan(ctx, L, Tg, Tu, ctx->sm_count - M3B_SIDE_CTA_RESERVE, &L.plan[1]));
  if (L.n_items) {
    k_mark_heads<<<(unsigned)((L.n_items + 255) / 256), 256, 0, ctx->stream>>>(L.idx, L.items, L.n_items);
    count_launch(ctx);
  }
  if (L.unit && L.unit->n_items) {
    TiledLayout& U = *L.unit;
    k_mark_heads<<<(unsigned)((U.n_items + 255) / 256), 256, 0, ctx->stream>>>(U.idx, U.items, U.n_items);
    count_launch(ctx);
    dev_free(U.item_scale);
    U.item_scale = nullptr;
    if (U.rowscale) {
      RET(dev_alloc(ctx, &U.item_scale, (size_t)U.n_items));
      k_item_scale<<<(unsigned)((U.n_items + 255) / 256), 256, 0, ctx->stream>>>(U.items, U.n_items, U.rowscale, U.item_scale);
      count_launch(ctx);
    }
  }
  CK(ctx, cudaGetLastError());
  return M3B_OK;
}

// Device-built plans (tables.cu): the same pieces as build_one_plan makes, computed per item on the
// GPU.  Every (visit, warp) owns a fixed number of piece slots; a slot no item fell into has zero
// groups and the kernel skips it.
int flat_build_plans_device(m3b_ctx* ctx, TiledLayout& L) {
  for (int k = 0; k < 2; ++k) {
    flat_plan_free(L.plan[k]);
    L.plan[k] = nullptr;
  }
  if (getenv("M3B_NO_FLAT")) return M3B_OK;
  if (L.n_entries / 2 >= (1LL << 31) || (L.unit && L.unit->n_entries / 8 >= (1LL << 31))) return M3B_OK;  // 32-bit group offsets
  // [1]: the reduced grid of the side-stream launch (the next cycle's first product under the restart SVD)
  const int n_plans = (ctx->sm_count > 2 * M3B_SIDE_CTA_RESERVE) ? 2 : 1;
  for (int k = 0; k < n_plans; ++k) {
    FlatPlan* P = new FlatPlan();
    L.plan[k] = P;
    P->n_cta = k == 0 ? ctx->sm_count : ctx->sm_count - M3B_SIDE_CTA_RESERVE;
    RET(tables_build_plan(ctx, L, P->n_cta, FL_WARPS, fl_piece_general(), fl_piece_unit(), &P->cta_visit_ptr, &P->visit_tile,
                          &P->gp_ptr, &P->gp, &P->up_ptr, &P->up, &P->n_visits, &P->n_pieces));
  }
  if (L.n_items) {
    k_mark_heads<<<(unsigned)((L.n_items + 255) / 256), 256, 0, ctx->stream>>>(L.idx, L.items, L.n_items);
    count_launch(ctx);
  }
  if (L.unit && L.unit->n_items) {
    TiledLayout& U = *L.unit;
    k_mark_heads<<<(unsigned)((U.n_items + 255) / 256), 256, 0, ctx->stream>>>(U.idx, U.items, U.n_items);
    count_launch(ctx);
    dev_free(U.item_scale);
    U.item_scale = nullptr;
    if (U.rowscale) {
      RET(dev_alloc(ctx, &U.item_scale, (size_t)U.n_items));
      k_item_scale<<<(unsigned)((U.n_items + 255) / 256), 256, 0, ctx->stream>>>(U.items, U.n_items, U.rowscale, U.item_scale);
      count_launch(ctx);
    }
  }
  CK(ctx, cudaGetLastError());
  return M3B_OK;
}

// ---------------------------------------------------------------------------
// device
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint4 fl_ld_uint4(const uint4* p) {
  uint4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
               : "l"(p));
  return v;
}
__device__ __forceinline__ unsigned int fl_ld_u32(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ double2 fl_ld_double2(const double2* p) {
  double2 v;
  asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
  return v;
}

// Gathers from the staged tile by shared-window byte address: index extraction + one IMAD
// (idx * 8 + base) + LDS.  Plain sx[idx] cost a shift, a mask and an add per gather.
__device__ __forceinline__ double fl_lds(unsigned addr) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
  return v;
}
// both addresses of an index pair: unpack the halves, one 16-bit multiply-add each
__device__ __forceinline__ void fl_addr2(unsigned pair, unsigned sbase, unsigned& alo, unsigned& ahi) {
  asm("{\n .reg .b16 l, h;\n mov.b32 {l, h}, %2;\n mad.wide.u16 %0, l, 8, %3;\n mad.wide.u16 %1, h, 8, %3;\n}"
      : "=r"(alo), "=r"(ahi)
      : "r"(pair), "r"(sbase));
}
// x[lo] + x[hi] of an index pair
__device__ __forceinline__ double fl_gsum(unsigned sbase, unsigned pair) {
  unsigned a0, a1;
  fl_addr2(pair, sbase, a0, a1);
  return fl_lds(a0) + fl_lds(a1);
}

struct FlatState {
  double acc;     // this lane's running sum of the open item
  int item;       // the open item, or (if none is open) the next item to start
  bool open;
  int wbase;      // window: lane l holds slot / scale of item wbase + l
  int wslot;
  double wscale;
};

__device__ __forceinline__ void fl_window(FlatState& st, int base, const int* __restrict__ slot,
                                          const double* __restrict__ scale, int n_items, int lane) {
  st.wbase = base;
  const int i = base + lane;
  st.wslot = (i < n_items) ? slot[i] : 0;
  st.wscale = (scale != nullptr && i < n_items) ? scale[i] : 1.0;
}

__device__ __forceinline__ double fl_warp_sum(double a) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
  return a;
}

// one round: v = this lane's group sum, mask = lanes whose group starts a new item
__device__ __forceinline__ void fl_round(FlatState& st, double v, unsigned mask, int lane, const int* __restrict__ slot,
                                         const double* __restrict__ scale, int n_items, double* __restrict__ partial) {
  if (mask == 0u) {
    st.acc += v;
    return;
  }
  const int h0 = __ffs(mask) - 1;
  const int nh = __popc(mask);
  const int first_new = st.open ? st.item + 1 : st.item;
  const int lo = st.open ? st.item : first_new;
  if (first_new + nh - 1 >= st.wbase + 32) fl_window(st, lo, slot, scale, n_items, lane);
  if (st.open) {
    const double tot = fl_warp_sum(st.acc + ((lane < h0) ? v : 0.0));
    const int sl = __shfl_sync(0xffffffffu, st.wslot, st.item - st.wbase);
    const double sc = __shfl_sync(0xffffffffu, st.wscale, st.item - st.wbase);
    if (lane == 0) partial[sl] = sc * tot;
  }
  if (nh == 1) {
    st.acc = (lane >= h0) ? v : 0.0;
  } else {
    // items that begin and end inside this round: segmented inclusive scan over the lanes >= h0
    const unsigned le = mask & (0xffffffffu >> (31 - lane));  // heads at or below this lane
    const bool in = lane >= h0;
    const int seg_start = in ? 31 - __clz(le) : 0;
    double s = in ? v : 0.0;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const double t = __shfl_up_sync(0xffffffffu, s, d);
      if (in && lane - d >= seg_start) s += t;
    }
    const bool last = (lane < 31) && ((mask >> (lane + 1)) & 1u);
    const int my_item = first_new + __popc(le) - 1;
    const int src = in ? (my_item - st.wbase) & 31 : 0;
    const int sl = __shfl_sync(0xffffffffu, st.wslot, src);
    const double sc = __shfl_sync(0xffffffffu, st.wscale, src);
    if (in && last) partial[sl] = sc * s;
    const int hl = 31 - __clz(mask);
    st.acc = (lane >= hl) ? v : 0.0;
  }
  st.item = first_new + nh - 1;
  st.open = true;
}

__device__ __forceinline__ void fl_finish(FlatState& st, int lane, const int* __restrict__ slot,
                                          const double* __restrict__ scale, int n_items, double* __restrict__ partial) {
  if (!st.open) return;
  const double tot = fl_warp_sum(st.acc);
  // (a round of 32 heads can leave the open item one past the window)
  if (st.item - st.wbase >= 32) fl_window(st, st.item, slot, scale, n_items, lane);
  const int sl = __shfl_sync(0xffffffffu, st.wslot, st.item - st.wbase);
  const double sc = __shfl_sync(0xffffffffu, st.wscale, st.item - st.wbase);
  if (lane == 0) partial[sl] = sc * tot;
}

template <int K>
__device__ __forceinline__ double fl_tree(const double (&v)[K]) {
  static_assert(K == 1 || K == 2 || K == 3 || K == 4 || K == 6, "fl_tree: unsupported batch depth");
  if (K == 1) return v[0];
  if (K == 6) return ((v[0] + v[1]) + (v[2] + v[3])) + (v[4 % K] + v[5 % K]);
  if (K == 4) return (v[0] + v[1]) + (v[2] + v[3 % K]);
  if (K == 3) return (v[0] + v[1]) + v[2 % K];
  return v[0] + v[1 % K];
}

struct FlatArgs {
  const unsigned short* gidx;
  const double* gval;
  const int* gslot;
  const int4* gp;
  const int* gp_ptr;
  int g_items;
  const unsigned short* uidx;
  const int* uslot;
  const double* uscale;
  const int4* up;
  const int* up_ptr;
  int u_items;
  const int* cta_visit_ptr;
  const int* visit_tile;
  int tile_w;
  long long cols;
  const double* x;
  const double* colscale;  // unit phase: tile *= colscale (layout B); nullptr on layout A
  double* partial;
};

template <int K, int THREADS>
__global__ void __launch_bounds__(THREADS, 1) k_flat_dot(const FlatArgs a) {
  constexpr int NWARPS = THREADS / 32;
  extern __shared__ double sx[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int tile_w = a.tile_w;
  const unsigned sbase = (unsigned)__cvta_generic_to_shared(sx);
  const int v0 = a.cta_visit_ptr[blockIdx.x], v1 = a.cta_visit_ptr[blockIdx.x + 1];
  for (int vis = v0; vis < v1; ++vis) {
    const int tile = a.visit_tile[vis];
    const long long off = (long long)tile * tile_w;
    const int len = (int)min((long long)tile_w, a.cols - off);
    if (vis > v0) __syncthreads();  // the previous tile is still being read
    {
      // stage the tile of the dense operand (coalesced 16-byte loads; x is L2-resident); the
      // slots [len, tile_w] read as 0 (tile_w is where the unit stream's pads point)
      const double* xs0 = a.x + off;
      if ((reinterpret_cast<unsigned long long>(xs0) & 15ull) == 0ull) {
        const double2* src = reinterpret_cast<const double2*>(xs0);
        double2* dst = reinterpret_cast<double2*>(sx);
        const int n2 = len >> 1;
        // five independent 16-byte loads in flight per thread (a plain copy loop kept one or two:
        // the staging showed up as 8 % of the stall samples)
        for (int k0 = threadIdx.x; k0 < n2; k0 += 5 * THREADS) {
          double2 v[5];
#pragma unroll
          for (int u = 0; u < 5; ++u) {
            const int k = k0 + u * THREADS;
            v[u] = (k < n2) ? src[k] : make_double2(0.0, 0.0);
          }
#pragma unroll
          for (int u = 0; u < 5; ++u) {
            const int k = k0 + u * THREADS;
            if (k < n2) dst[k] = v[u];
          }
        }
        if ((len & 1) && threadIdx.x == 0) sx[len - 1] = xs0[len - 1];
      } else {
        for (int k = threadIdx.x; k < len; k += blockDim.x) sx[k] = xs0[k];
      }
      for (int k = len + threadIdx.x; k <= tile_w; k += blockDim.x) sx[k] = 0.0;
      __syncthreads();
    }
    // ---- general stream: groups of 2 entries (u32 index pair + double2 values)
    {
      const int p0 = a.gp_ptr[vis * NWARPS + warp], p1 = a.gp_ptr[vis * NWARPS + warp + 1];
      for (int p = p0; p < p1; ++p) {
        const int4 pc = a.gp[p];
        const unsigned int* ip = reinterpret_cast<const unsigned int*>(a.gidx) + pc.x;
        const double2* vp = reinterpret_cast<const double2*>(a.gval) + pc.x;
        const int ng = pc.y, nr = (ng + 31) >> 5;
        if (ng <= 0) continue;  // an empty slot of a device-built plan
        FlatState st;
        st.acc = 0.0;
        st.item = pc.z;
        st.open = false;
        fl_window(st, pc.z, a.gslot, nullptr, a.g_items, lane);
        // Two register-resident batches of K rounds, software-pipelined: batch k+1 is in flight
        // while batch k is consumed.  What the SASS of earlier versions taught:
        //  * loads are UNCONDITIONAL (addresses past the end are clamped to the last group, the data
        //    is ignored when the round is consumed): "in range ? load : 0" became a load into a
        //    temporary with a select right behind it, which waits for the load where it is issued;
        //  * a finer ring (one load per step, 7 in flight) ran at 2.7 TB/s: a warp has six
        //    scoreboards, ptxas has to share them between the slots, and waiting for the oldest
        //    load then waits for the newest one on the same scoreboard.  Two big batches map onto
        //    two scoreboards.
        unsigned int i0[K], i1[K];
        double2 w0[K], w1[K];
#define FL_GLOAD(BI, BV, R0)                                        \
  _Pragma("unroll") for (int u = 0; u < K; ++u) {                \
    const int g = min(32 * ((R0) + u) + lane, ng - 1);              \
    BI[u] = fl_ld_u32(ip + g);                                      \
    BV[u] = fl_ld_double2(vp + g);                                  \
  }
  // all gathers / products of a batch first (independent => the LDS and DFMA latencies overlap),
  // then the item bookkeeping; a batch without an item start is just one tree of adds
#define FL_GUSE(BI, BV, R0)                                                                     \
  {                                                                                             \
    double v[K];                                                                             \
    unsigned mk[K], any = 0u;                                                                \
    _Pragma("unroll") for (int u = 0; u < K; ++u) {                                          \
      const bool valid = 32 * ((R0) + u) + lane < ng;                                           \
      mk[u] = __ballot_sync(0xffffffffu, valid && ((BI[u] >> 15) & 1u));                        \
      any |= mk[u];                                                                             \
      unsigned ga, gb;                                                                          \
      fl_addr2(BI[u] & 0xffff7fffu, sbase, ga, gb);                                             \
      const double t = fma(BV[u].y, fl_lds(gb), BV[u].x * fl_lds(ga));                          \
      v[u] = valid ? t : 0.0;                                                                   \
    }                                                                                           \
    if (any == 0u) {                                                                            \
      st.acc += fl_tree<K>(v);                                                                     \
    } else {                                                                                    \
      _Pragma("unroll") for (int u = 0; u < K; ++u)                                          \
        if ((R0) + u < nr) fl_round(st, v[u], mk[u], lane, a.gslot, nullptr, a.g_items, a.partial); \
    }                                                                                           \
  }
        FL_GLOAD(i0, w0, 0)
        for (int r0 = 0; r0 < nr; r0 += 2 * K) {
          FL_GLOAD(i1, w1, r0 + K)
          FL_GUSE(i0, w0, r0)
          if (r0 + K >= nr) break;
          FL_GLOAD(i0, w0, r0 + 2 * K)
          FL_GUSE(i1, w1, r0 + K)
        }
#undef FL_GLOAD
#undef FL_GUSE
        fl_finish(st, lane, a.gslot, nullptr, a.g_items, a.partial);
      }
    }
    // ---- unit stream: groups of 8 indices, value applied on the dense side
    if (a.uidx) {
      if (a.colscale) {
        __syncthreads();
        const double* cs = a.colscale + off;
        if ((reinterpret_cast<unsigned long long>(cs) & 15ull) == 0ull) {
          const double2* cs2 = reinterpret_cast<const double2*>(cs);
          double2* d2 = reinterpret_cast<double2*>(sx);
          const int n2 = len >> 1;
          for (int k0 = threadIdx.x; k0 < n2; k0 += 5 * THREADS) {
            double2 v[5];
#pragma unroll
            for (int u = 0; u < 5; ++u) {
              const int k = k0 + u * THREADS;
              v[u] = (k < n2) ? cs2[k] : make_double2(1.0, 1.0);
            }
#pragma unroll
            for (int u = 0; u < 5; ++u) {
              const int k = k0 + u * THREADS;
              if (k < n2) {
                double2 t = d2[k];
                t.x *= v[u].x;
                t.y *= v[u].y;
                d2[k] = t;
              }
            }
          }
          if ((len & 1) && threadIdx.x == 0) sx[len - 1] *= cs[len - 1];
        } else {
          for (int k = threadIdx.x; k < len; k += blockDim.x) sx[k] *= cs[k];
        }
        __syncthreads();
      }
      const int p0 = a.up_ptr[vis * NWARPS + warp], p1 = a.up_ptr[vis * NWARPS + warp + 1];
      for (int p = p0; p < p1; ++p) {
        const int4 pc = a.up[p];
        const uint4* ip = reinterpret_cast<const uint4*>(a.uidx) + pc.x;
        const int ng = pc.y, nr = (ng + 31) >> 5;
        if (ng <= 0) continue;
        FlatState st;
        st.acc = 0.0;
        st.item = pc.z;
        st.open = false;
        fl_window(st, pc.z, a.uslot, a.uscale, a.u_items, lane);
        uint4 q0[K], q1[K];
#define FL_ULOAD(B, R0) \
  _Pragma("unroll") for (int u = 0; u < K; ++u) B[u] = fl_ld_uint4(ip + min(32 * ((R0) + u) + lane, ng - 1));
#define FL_UUSE(B, R0)                                                                                      \
  {                                                                                                         \
    double v[K];                                                                                         \
    unsigned mk[K], any = 0u;                                                                            \
    _Pragma("unroll") for (int u = 0; u < K; ++u) {                                                      \
      const uint4 b = B[u];                                                                                 \
      const bool valid = 32 * ((R0) + u) + lane < ng;                                                       \
      mk[u] = __ballot_sync(0xffffffffu, valid && ((b.x >> 15) & 1u));                                      \
      any |= mk[u];                                                                                         \
      const double t = (fl_gsum(sbase, b.x & 0xffff7fffu) + fl_gsum(sbase, b.y)) +                          \
                       (fl_gsum(sbase, b.z) + fl_gsum(sbase, b.w));                                         \
      v[u] = valid ? t : 0.0;                                                                               \
    }                                                                                                       \
    if (any == 0u) {                                                                                        \
      st.acc += fl_tree<K>(v);                                                                                 \
    } else {                                                                                                \
      _Pragma("unroll") for (int u = 0; u < K; ++u)                                                      \
        if ((R0) + u < nr) fl_round(st, v[u], mk[u], lane, a.uslot, a.uscale, a.u_items, a.partial);        \
    }                                                                                                       \
  }
        FL_ULOAD(q0, 0)
        for (int r0 = 0; r0 < nr; r0 += 2 * K) {
          FL_ULOAD(q1, r0 + K)
          FL_UUSE(q0, r0)
          if (r0 + K >= nr) break;
          FL_ULOAD(q0, r0 + 2 * K)
          FL_UUSE(q1, r0 + K)
        }
#undef FL_ULOAD
#undef FL_UUSE
        fl_finish(st, lane, a.uslot, a.uscale, a.u_items, a.partial);
      }
    }
  }
}


// ---------------------------------------------------------------------------
// Ring variant (M3B_FLAT_VARIANT=3 / 4): same plan, same item bookkeeping, same results per item
// up to the order of the adds -- but the stream travels global -> SHARED memory with bulk async
// copies (cp.async.bulk + mbarrier complete_tx, the 1-D TMA path) into a small ring that every warp
// owns, instead of through register-resident batches.  Why: with the batches a warp has at most
// one batch (1.5-1.9 KB) in flight and only while it is busy with the other one; ncu on the batch
// kernel shows long_scoreboard as the top stall (6.6 warps per issue), issue slots 50 % busy, DRAM
// 48 %.  With the ring the copies are issued by one lane, cost no registers, and a stage is refilled
// the moment it has been read, so the whole ring (2-3 KB per warp, 64-96 KB per SM) stays in flight
// whatever the consumer is doing.  A stage is R rounds: 512 R bytes of the unit stream, or
// 128 R (+16: the index pairs of a piece start on a 4-byte boundary, the copy on a 16-byte one)
// + 512 R bytes of the general stream.  Shared memory: tile | ring (NWARPS x RB) | mbarriers.
// ---------------------------------------------------------------------------
#define FR_NS_MAX 8
__device__ __forceinline__ void fr_mbar_init(unsigned bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void fr_expect_tx(unsigned bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void fr_bulk(unsigned dst, const void* src, unsigned bytes, unsigned bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
               "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void fr_wait(unsigned bar, unsigned parity) {
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
__device__ __forceinline__ uint4 fr_lds128(unsigned addr) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ unsigned fr_lds32(unsigned addr) {
  unsigned v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ double2 fr_lds_d2(unsigned addr) {
  double2 v;
  asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr));
  return v;
}

// walks the chunks (<= 32 R groups, never across a piece boundary) of a warp's pieces of one phase
struct FrIter {
  int p, p1;  // current piece, end of the warp's range
  int g0, ng, item0;
  int off;    // groups of the current piece already handed out
};
__device__ __forceinline__ void fr_iter_load(FrIter& it, const int4* __restrict__ pieces) {
  it.ng = 0;
  while (it.p < it.p1) {
    const int4 pc = pieces[it.p];
    if (pc.y > 0) {
      it.g0 = pc.x;
      it.ng = pc.y;
      it.item0 = pc.z;
      break;
    }
    ++it.p;  // an empty slot of a device-built plan
  }
  it.off = 0;
}
__device__ __forceinline__ void fr_iter_next(FrIter& it, int n, const int4* __restrict__ pieces) {
  it.off += n;
  if (it.off >= it.ng) {
    ++it.p;
    fr_iter_load(it, pieces);
  }
}

template <int R, int THREADS>
__global__ void __launch_bounds__(THREADS, 1) k_flat_dot_ring(const FlatArgs a, const int ring_off, const int rb) {
  constexpr int NWARPS = THREADS / 32;
  constexpr int GPC = 32 * R;                                // groups per chunk
  constexpr int SBU = 512 * R;                               // stage bytes, unit stream
  constexpr int IDXR = (128 * R + 16 + 15) & ~15;            // index region of a general stage
  constexpr int SBG = IDXR + 512 * R;                        // stage bytes, general stream
  extern __shared__ double sx[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int tile_w = a.tile_w;
  const unsigned sbase = (unsigned)__cvta_generic_to_shared(sx);
  const unsigned ring = sbase + (unsigned)ring_off + (unsigned)(warp * rb);
  const unsigned bars = sbase + (unsigned)ring_off + (unsigned)(NWARPS * rb) + (unsigned)(warp * FR_NS_MAX * 8);
  const int nsu = min(FR_NS_MAX, rb / SBU), nsg = min(FR_NS_MAX, rb / SBG);
  if (lane < FR_NS_MAX) fr_mbar_init(bars + 8u * lane, 1u);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncwarp();
  unsigned phase = 0u;  // bit s: parity the next wait on barrier s uses
  const int v0 = a.cta_visit_ptr[blockIdx.x], v1 = a.cta_visit_ptr[blockIdx.x + 1];
  for (int vis = v0; vis < v1; ++vis) {
    const int tile = a.visit_tile[vis];
    const long long off = (long long)tile * tile_w;
    const int len = (int)min((long long)tile_w, a.cols - off);
    if (vis > v0) __syncthreads();  // the previous tile is still being read
    // ---- general stream: the first stages go out before the tile is staged (they only touch the ring)
    FrIter gi_is, gi_co;
    gi_is.p = gi_co.p = a.gp_ptr[vis * NWARPS + warp];
    gi_is.p1 = gi_co.p1 = a.gp_ptr[vis * NWARPS + warp + 1];
    fr_iter_load(gi_is, a.gp);
    gi_co = gi_is;
#define FR_GISSUE(S)                                                                              \
  {                                                                                               \
    const int n_ = min(GPC, gi_is.ng - gi_is.off);                                                \
    const long long g_ = (long long)gi_is.g0 + gi_is.off;                                         \
    const unsigned m_ = (unsigned)((4 * g_) & 15);                                                \
    const unsigned bi_ = (m_ + 4u * n_ + 15u) & ~15u, bv_ = 16u * n_;                             \
    if (lane == 0) {                                                                              \
      const unsigned st_ = ring + (unsigned)((S) * SBG), bar_ = bars + 8u * (S);                  \
      fr_expect_tx(bar_, bi_ + bv_);                                                              \
      fr_bulk(st_, reinterpret_cast<const char*>(a.gidx) + 4 * g_ - m_, bi_, bar_);               \
      fr_bulk(st_ + IDXR, reinterpret_cast<const char*>(a.gval) + 16 * g_, bv_, bar_);            \
    }                                                                                             \
    fr_iter_next(gi_is, n_, a.gp);                                                                \
  }
    for (int s = 0; s < nsg; ++s)
      if (gi_is.p < gi_is.p1) FR_GISSUE(s)
    {
      // stage the tile of the dense operand (coalesced 16-byte loads; x is L2-resident); the
      // slots [len, tile_w] read as 0 (tile_w is where the unit stream's pads point)
      const double* xs0 = a.x + off;
      if ((reinterpret_cast<unsigned long long>(xs0) & 15ull) == 0ull) {
        const double2* src = reinterpret_cast<const double2*>(xs0);
        double2* dst = reinterpret_cast<double2*>(sx);
        const int n2 = len >> 1;
        for (int k0 = threadIdx.x; k0 < n2; k0 += 5 * THREADS) {
          double2 v[5];
#pragma unroll
          for (int u = 0; u < 5; ++u) {
            const int k = k0 + u * THREADS;
            v[u] = (k < n2) ? src[k] : make_double2(0.0, 0.0);
          }
#pragma unroll
          for (int u = 0; u < 5; ++u) {
            const int k = k0 + u * THREADS;
            if (k < n2) dst[k] = v[u];
          }
        }
        if ((len & 1) && threadIdx.x == 0) sx[len - 1] = xs0[len - 1];
      } else {
        for (int k = threadIdx.x; k < len; k += blockDim.x) sx[k] = xs0[k];
      }
      for (int k = len + threadIdx.x; k <= tile_w; k += blockDim.x) sx[k] = 0.0;
      __syncthreads();
    }
    {
      FlatState st;
      st.acc = 0.0;
      st.item = 0;
      st.open = false;
      st.wbase = 0;
      st.wslot = 0;
      st.wscale = 1.0;
      bool any_piece = false;
      int s = 0;
      while (gi_co.p < gi_co.p1) {
        if (gi_co.off == 0) {
          if (any_piece) fl_finish(st, lane, a.gslot, nullptr, a.g_items, a.partial);
          any_piece = true;
          st.acc = 0.0;
          st.item = gi_co.item0;
          st.open = false;
          fl_window(st, gi_co.item0, a.gslot, nullptr, a.g_items, lane);
        }
        const int n = min(GPC, gi_co.ng - gi_co.off);
        const unsigned m = (unsigned)((4 * ((long long)gi_co.g0 + gi_co.off)) & 15);
        const unsigned stg = ring + (unsigned)(s * SBG);
        fr_wait(bars + 8u * s, (phase >> s) & 1u);
        phase ^= 1u << s;
        double v[R];
        unsigned mk[R], any = 0u;
#pragma unroll
        for (int u = 0; u < R; ++u) {
          const int g = 32 * u + lane;
          const bool valid = g < n;
          const int gc = min(g, n - 1);
          const unsigned pair = fr_lds32(stg + m + 4u * gc);
          const double2 w = fr_lds_d2(stg + IDXR + 16u * gc);
          mk[u] = __ballot_sync(0xffffffffu, valid && ((pair >> 15) & 1u));
          any |= mk[u];
          unsigned ga, gb;
          fl_addr2(pair & 0xffff7fffu, sbase, ga, gb);
          const double t = fma(w.y, fl_lds(gb), w.x * fl_lds(ga));
          v[u] = valid ? t : 0.0;
        }
        if (any == 0u) {
          st.acc += fl_tree<R>(v);
        } else {
#pragma unroll
          for (int u = 0; u < R; ++u)
            if (32 * u < n) fl_round(st, v[u], mk[u], lane, a.gslot, nullptr, a.g_items, a.partial);
        }
        __syncwarp();
        if (gi_is.p < gi_is.p1) FR_GISSUE(s)
        fr_iter_next(gi_co, n, a.gp);
        s = (s + 1 == nsg) ? 0 : s + 1;
      }
      if (any_piece) fl_finish(st, lane, a.gslot, nullptr, a.g_items, a.partial);
    }
#undef FR_GISSUE
    // ---- unit stream: groups of 8 indices, value applied on the dense side
    if (a.uidx) {
      FrIter ui_is, ui_co;
      ui_is.p = a.up_ptr[vis * NWARPS + warp];
      ui_is.p1 = a.up_ptr[vis * NWARPS + warp + 1];
      fr_iter_load(ui_is, a.up);
      ui_co = ui_is;
#define FR_UISSUE(S)                                                                              \
  {                                                                                               \
    const int n_ = min(GPC, ui_is.ng - ui_is.off);                                                \
    const long long g_ = (long long)ui_is.g0 + ui_is.off;                                         \
    if (lane == 0) {                                                                              \
      const unsigned bar_ = bars + 8u * (S);                                                      \
      fr_expect_tx(bar_, 16u * n_);                                                               \
      fr_bulk(ring + (unsigned)((S) * SBU), reinterpret_cast<const char*>(a.uidx) + 16 * g_, 16u * n_, bar_); \
    }                                                                                             \
    fr_iter_next(ui_is, n_, a.up);                                                                \
  }
      // (the general phase has drained the ring: every copy it issued has been waited for and read)
      for (int s = 0; s < nsu; ++s)
        if (ui_is.p < ui_is.p1) FR_UISSUE(s)
      if (a.colscale) {
        __syncthreads();
        const double* cs = a.colscale + off;
        if ((reinterpret_cast<unsigned long long>(cs) & 15ull) == 0ull) {
          const double2* cs2 = reinterpret_cast<const double2*>(cs);
          double2* d2 = reinterpret_cast<double2*>(sx);
          const int n2 = len >> 1;
          for (int k0 = threadIdx.x; k0 < n2; k0 += 5 * THREADS) {
            double2 v[5];
#pragma unroll
            for (int u = 0; u < 5; ++u) {
              const int k = k0 + u * THREADS;
              v[u] = (k < n2) ? cs2[k] : make_double2(1.0, 1.0);
            }
#pragma unroll
            for (int u = 0; u < 5; ++u) {
              const int k = k0 + u * THREADS;
              if (k < n2) {
                double2 t = d2[k];
                t.x *= v[u].x;
                t.y *= v[u].y;
                d2[k] = t;
              }
            }
          }
          if ((len & 1) && threadIdx.x == 0) sx[len - 1] *= cs[len - 1];
        } else {
          for (int k = threadIdx.x; k < len; k += blockDim.x) sx[k] *= cs[k];
        }
        __syncthreads();
      }
      FlatState st;
      st.acc = 0.0;
      st.item = 0;
      st.open = false;
      st.wbase = 0;
      st.wslot = 0;
      st.wscale = 1.0;
      bool any_piece = false;
      int s = 0;
      while (ui_co.p < ui_co.p1) {
        if (ui_co.off == 0) {
          if (any_piece) fl_finish(st, lane, a.uslot, a.uscale, a.u_items, a.partial);
          any_piece = true;
          st.acc = 0.0;
          st.item = ui_co.item0;
          st.open = false;
          fl_window(st, ui_co.item0, a.uslot, a.uscale, a.u_items, lane);
        }
        const int n = min(GPC, ui_co.ng - ui_co.off);
        const unsigned stg = ring + (unsigned)(s * SBU);
        fr_wait(bars + 8u * s, (phase >> s) & 1u);
        phase ^= 1u << s;
        double v[R];
        unsigned mk[R], any = 0u;
#pragma unroll
        for (int u = 0; u < R; ++u) {
          const int g = 32 * u + lane;
          const bool valid = g < n;
          const uint4 b = fr_lds128(stg + 16u * min(g, n - 1));
          mk[u] = __ballot_sync(0xffffffffu, valid && ((b.x >> 15) & 1u));
          any |= mk[u];
          const double t = (fl_gsum(sbase, b.x & 0xffff7fffu) + fl_gsum(sbase, b.y)) +
                           (fl_gsum(sbase, b.z) + fl_gsum(sbase, b.w));
          v[u] = valid ? t : 0.0;
        }
        if (any == 0u) {
          st.acc += fl_tree<R>(v);
        } else {
#pragma unroll
          for (int u = 0; u < R; ++u)
            if (32 * u < n) fl_round(st, v[u], mk[u], lane, a.uslot, a.uscale, a.u_items, a.partial);
        }
        __syncwarp();
        if (ui_is.p < ui_is.p1) FR_UISSUE(s)
        fr_iter_next(ui_co, n, a.up);
        s = (s + 1 == nsu) ? 0 : s + 1;
      }
      if (any_piece) fl_finish(st, lane, a.uslot, a.uscale, a.u_items, a.partial);
#undef FR_UISSUE
    }
  }
}

// shared-memory carve of the ring variant for a tile width; false when the ring would be too small
static bool fr_carve(int tile_w, int warps, int rounds, int* ring_off, int* rb, int* total) {
  const int tile_bytes = (int)(((size_t)tile_w + 2) * sizeof(double) + 127) & ~127;
  const int bar_bytes = warps * FR_NS_MAX * 8;
  const int avail = 232448 - tile_bytes - bar_bytes;  // 227 KB per CTA
  if (avail <= 0) return false;
  int per = (avail / warps) & ~127;
  per = std::min(per, 4096);
  const int sbg = ((128 * rounds + 16 + 15) & ~15) + 512 * rounds;
  if (per / sbg < 2) return false;
  *ring_off = tile_bytes;
  *rb = per;
  *total = tile_bytes + warps * per + bar_bytes;
  return true;
}

template <int K, int THREADS>
static int fl_init_one(m3b_ctx* ctx) {
  CK(ctx, cudaFuncSetAttribute(k_flat_dot<K, THREADS>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               (int)((M3B_TILE_W_MAX + 2) * sizeof(double))));
  return M3B_OK;
}

int flat_init(m3b_ctx* ctx) {
  RET((fl_init_one<6, 512>(ctx)));
  RET((fl_init_one<3, 1024>(ctx)));
  RET((fl_init_one<2, 1024>(ctx)));
  CK(ctx, cudaFuncSetAttribute(k_flat_dot_ring<1, 1024>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
  CK(ctx, cudaFuncSetAttribute(k_flat_dot_ring<2, 1024>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
  return M3B_OK;
}

int flat_launch_dot(m3b_ctx* ctx, const TiledLayout& L, const double* x, cudaStream_t stream, int max_cta) {
  const FlatPlan* P = L.plan[0];
  if (max_cta > 0 && L.plan[1] && L.plan[1]->n_cta == max_cta) P = L.plan[1];
  if (!P || P->n_visits == 0) return M3B_OK;
  const TiledLayout* U = L.unit;
  FlatArgs a;
  a.gidx = L.idx;
  a.gval = L.val;
  a.gslot = L.item_slot;
  a.gp = P->gp;
  a.gp_ptr = P->gp_ptr;
  a.g_items = (int)L.n_items;
  a.uidx = (U && U->n_items) ? U->idx : nullptr;
  a.uslot = U ? U->item_slot : nullptr;
  a.uscale = U ? U->item_scale : nullptr;
  a.up = P->up;
  a.up_ptr = P->up_ptr;
  a.u_items = U ? (int)U->n_items : 0;
  a.cta_visit_ptr = P->cta_visit_ptr;
  a.visit_tile = P->visit_tile;
  a.tile_w = L.tile_w;
  a.cols = L.cols;
  a.x = x;
  a.colscale = U ? U->colscale : nullptr;
  a.partial = L.partial;
  const size_t smem = ((size_t)L.tile_w + 2) * sizeof(double);
  cudaStream_t st = stream ? stream : ctx->stream;
  int variant = fl_variant();
  int ring_off = 0, rb = 0, ring_total = 0;
  if (variant >= 3 && !fr_carve(L.tile_w, 32, FL_VARIANTS[variant].k, &ring_off, &rb, &ring_total)) variant = 1;  // tile too wide for a ring
  switch (variant) {
    case 3: k_flat_dot_ring<1, 1024><<<P->n_cta, 1024, ring_total, st>>>(a, ring_off, rb); break;
    case 4: k_flat_dot_ring<2, 1024><<<P->n_cta, 1024, ring_total, st>>>(a, ring_off, rb); break;
    case 1: k_flat_dot<3, 1024><<<P->n_cta, 1024, smem, st>>>(a); break;
    case 2: k_flat_dot<2, 1024><<<P->n_cta, 1024, smem, st>>>(a); break;
    default: k_flat_dot<6, 512><<<P->n_cta, 512, smem, st>>>(a); break;
  }
  count_launch(ctx);
  CK(ctx, cudaGetLastError());
  return M3B_OK;
}
