// Per-element bodies of the device-side table builder (tables.cu).
//
// m3b_select_genes used to build the item / slot / chunk tables and the flat work plans of the
// two tiled layouts with host loops over ~10^6 runs and items (101 ms of a 494 ms step at the
// cfg3 shape, 240 ms of the cfg4 step with eight processes fighting for the host).  The same
// tables are now produced by small kernels and prefix sums on the device.  Every kernel is a thin
// wrapper around one of the functions below, which are __host__ __device__ so that the host test
// (tools/host_tables_check.cu) can run the very same code against the old host builders.
//
// Vocabulary (m3b_internal.h): a RUN is the entries of one input row inside one tile (padded to
// 4 / 8 entries); an ITEM is a run or a <= item_max slice of one; a row's partial SLOTS are its
// general items (tile order) followed by its unit items; a CHUNK is a short item range (the unit
// the per-item kernels claim); a PIECE is an item-aligned, contiguous slice of a tile's stream
// owned by one warp of the flat product kernel.
#pragma once
#include <stdint.h>

#include "m3b_internal.h"

#ifdef __CUDACC__
#define TB_HD __host__ __device__ __forceinline__
#else
#define TB_HD inline
#endif

struct TabShape {
  int pad;        // run padding (4 general, 8 unit)
  int item_max;   // entries per item
  int item_tail;  // entries per item in the last fifth of a tile
  int group;      // entries per group of the flat kernel (2 general, 8 unit)
};

TB_HD long long tb_padn(long long n, int pad) { return (n + pad - 1) / pad * pad; }

// ---- runs ---------------------------------------------------------------------------------
// padded length of a run, for all rows (storage) and for kept rows only (work)
TB_HD void tb_run_plen(long long run, const int* cnt, const int* row_map, long long rows_in, int pad, long long* plen_all,
                       long long* plen_kept, long long* nnz_kept) {
  const long long r = run % rows_in;
  const long long pl = tb_padn(cnt[run], pad);
  if (plen_all) plen_all[run] = pl;
  const bool kept = !row_map || row_map[r] >= 0;
  plen_kept[run] = kept ? pl : 0;
  *nnz_kept = kept ? cnt[run] : 0;
}

// number of items a run of `len` padded entries is cut into when it starts at position `seen` of
// its tile's (kept) stream: <= item_max entries per item, <= item_tail from `tail_from` on
TB_HD int tb_item_count(long long len, long long seen, long long tail_from, int item_max, int item_tail) {
  int n = 0;
  while (len > 0) {
    const int lim = (seen >= tail_from) ? item_tail : item_max;
    const long long l = len < lim ? len : (long long)lim;
    ++n;
    seen += l;
    len -= l;
  }
  return n;
}

// kcum: exclusive prefix sums of the kept padded lengths over all runs (n_runs + 1 entries)
TB_HD int tb_run_items(long long run, const int* cnt, const long long* kcum, long long rows_in, const TabShape sh) {
  const long long t = run / rows_in;
  const long long len = kcum[run + 1] - kcum[run];
  if (len == 0 || cnt[run] == 0) return 0;
  const long long t0 = kcum[t * rows_in], tlen = kcum[(t + 1) * rows_in] - t0;
  return tb_item_count(len, kcum[run] - t0, tlen - tlen / 5, sh.item_max, sh.item_tail);
}

// ---- slots --------------------------------------------------------------------------------
// one call per INPUT row: relative slot offsets of the row's runs (general then unit) and the
// row's totals
TB_HD void tb_row_slots(long long r, const int* row_map, long long rows_in, int n_tiles, const int* nit_g, const int* nit_u,
                        int* sbase_g, int* sbase_u, int* rowtot, int* gtot) {
  if (row_map && row_map[r] < 0) return;
  const long long ro = row_map ? row_map[r] : r;
  int acc = 0;
  for (int t = 0; t < n_tiles; ++t) {
    sbase_g[(long long)t * rows_in + r] = acc;
    acc += nit_g[(long long)t * rows_in + r];
  }
  const int g = acc;
  if (nit_u) {
    acc = 0;
    for (int t = 0; t < n_tiles; ++t) {
      sbase_u[(long long)t * rows_in + r] = acc;
      acc += nit_u[(long long)t * rows_in + r];
    }
    rowtot[ro] = g + acc;
  } else {
    rowtot[ro] = g;
  }
  gtot[ro] = g;
}

// ---- items --------------------------------------------------------------------------------
struct TabItemOut {
  ItemDesc* items;
  int* item_slot;
  long long* item_cum;  // position of the item inside its tile's kept stream (entries)
  int* item_tile;
  int* item_w;          // scheduling weight of the item (per-item kernels' chunks)
  int* row_item_ptr;    // [n_tiles][rows_out + 1]
};

// one call per run: writes the run's items and the (tile, row) -> item range table
TB_HD void tb_fill_items(long long run, const int* cnt, const int* row_map, long long rows_in, long long rows_out,
                         const long long* base, const long long* kcum, const int* ioff, const int* sbase,
                         const int* rsp, const int* gtot_or_null, const TabShape sh, const TabItemOut o) {
  const long long t = run / rows_in, r = run % rows_in;
  const bool kept = !row_map || row_map[r] >= 0;
  const long long ro = row_map ? row_map[r] : r;
  if (kept) o.row_item_ptr[t * (rows_out + 1) + ro] = ioff[run];
  if (r == rows_in - 1) o.row_item_ptr[t * (rows_out + 1) + rows_out] = ioff[run + 1];
  const int n = ioff[run + 1] - ioff[run];
  if (n == 0) return;
  const long long t0 = kcum[t * rows_in], tlen = kcum[(t + 1) * rows_in] - t0;
  const long long tail_from = tlen - tlen / 5;
  long long len = kcum[run + 1] - kcum[run], seen = kcum[run] - t0, start = base[run];
  int i = ioff[run];
  int slot = rsp[ro] + (gtot_or_null ? gtot_or_null[ro] : 0) + sbase[run];
  while (len > 0) {
    const int lim = (seen >= tail_from) ? sh.item_tail : sh.item_max;
    const int l = (int)(len < lim ? len : (long long)lim);
    ItemDesc d;
    d.start = start;
    d.len = l;
    d.row = (int)ro;
    o.items[i] = d;
    o.item_slot[i] = slot;
    o.item_cum[i] = seen;
    o.item_tile[i] = (int)t;
    // tail items weigh item_max / item_tail times more, so that a fixed weight per chunk gives
    // ~item_max entries per chunk in the body of a tile and ~item_tail at its end
    o.item_w[i] = (l + 32) * (lim == sh.item_tail ? sh.item_max / sh.item_tail : 1);
    ++i;
    ++slot;
    start += l;
    seen += l;
    len -= l;
  }
}

// ---- chunks -------------------------------------------------------------------------------
// wcum: exclusive prefix sums of item_w.  An item starts a chunk when it is the first of its tile
// or when the running weight inside the tile crosses a multiple of item_max.
TB_HD int tb_chunk_head(long long i, const long long* wcum, const int* item_tile, const int* row_item_ptr, long long rows_out,
                        int item_max) {
  const int t = item_tile[i];
  const int f = row_item_ptr[(long long)t * (rows_out + 1)];
  if (i == f) return 1;
  const long long w0 = wcum[f];
  return ((wcum[i] - w0) / item_max) != ((wcum[i - 1] - w0) / item_max);
}

// ---- flat plan ----------------------------------------------------------------------------
// Everything the piece kernel needs to know about a tile (one record per tile and plan)
struct PlanTile {
  int nc;        // CTAs that own the tile (0: the tile is empty)
  int visit0;    // visit id of its first CTA (visits of a tile's CTAs are consecutive)
  int segs[2];   // pieces per warp, general / unit stream
};

// tiles -> CTAs, visits and the (visit, warp) -> piece-slot ranges.  Serial: n_tiles and n_cta
// are a few hundred at most.  Mirrors build_one_plan of the host builder (flat.cu).
//   bytes[t] = 10 * tlen_g[t] + 2 * tlen_u[t]
//   outputs: pt[n_tiles], cta_visit_ptr[n_cta + 1], visit_tile[<= max(n_cta, n_tiles)],
//            gp_ptr / up_ptr [n_visits * warps + 1], counts[0] = n_visits, [1] = gp slots, [2] = up slots
// scratch: n_of[n_tiles] ints, order[n_tiles] ints, load[n_cta] doubles, cta_of[n_tiles] ints
TB_HD void tb_plan_meta(int n_tiles, int n_cta, int warps, const long long* tlen_g, const long long* tlen_u,
                        long long target_g, long long target_u, PlanTile* pt, int* cta_visit_ptr, int* visit_tile,
                        int* gp_ptr, int* up_ptr, int* counts, int* n_of, int* order, double* load, int* cta_of) {
  double total = 0;
  int nonempty = 0;
  for (int t = 0; t < n_tiles; ++t) {
    const double b = 10.0 * (double)tlen_g[t] + 2.0 * (double)(tlen_u ? tlen_u[t] : 0);
    total += b;
    nonempty += b > 0;
    pt[t].nc = 0;
    pt[t].visit0 = 0;
    pt[t].segs[0] = pt[t].segs[1] = 0;
    n_of[t] = 0;
    cta_of[t] = -1;
  }
#define TB_BYTES(t) (10.0 * (double)tlen_g[t] + 2.0 * (double)(tlen_u ? tlen_u[t] : 0))
  int n_visits = 0;
  for (int c = 0; c <= n_cta; ++c) cta_visit_ptr[c] = 0;
  if (total > 0) {
    if (n_tiles <= n_cta) {
      int used = 0;
      for (int t = 0; t < n_tiles; ++t)
        if (TB_BYTES(t) > 0) {
          int v = (int)(TB_BYTES(t) / total * (n_cta - nonempty)) + 1;
          n_of[t] = v < 1 ? 1 : v;
          used += n_of[t];
        }
      while (used < n_cta) {
        int best = -1;
        for (int t = 0; t < n_tiles; ++t)
          if (TB_BYTES(t) > 0 && (best < 0 || TB_BYTES(t) / n_of[t] > TB_BYTES(best) / n_of[best])) best = t;
        n_of[best]++;
        used++;
      }
      while (used > n_cta) {
        int best = -1;
        for (int t = 0; t < n_tiles; ++t)
          if (n_of[t] > 1 && (best < 0 || TB_BYTES(t) / n_of[t] < TB_BYTES(best) / n_of[best])) best = t;
        n_of[best]--;
        used--;
      }
      int c = 0;
      for (int t = 0; t < n_tiles; ++t) {
        pt[t].nc = n_of[t];
        pt[t].visit0 = n_visits;
        for (int k = 0; k < n_of[t]; ++k, ++c) {
          cta_visit_ptr[c] = n_visits;
          visit_tile[n_visits++] = t;
        }
      }
      for (; c <= n_cta; ++c) cta_visit_ptr[c] = n_visits;
    } else {
      // more tiles than CTAs: heaviest tile first (stable), each to the least loaded CTA
      for (int t = 0; t < n_tiles; ++t) order[t] = t;
      for (int a = 1; a < n_tiles; ++a) {  // stable insertion sort, descending bytes
        const int v = order[a];
        int b = a - 1;
        while (b >= 0 && TB_BYTES(order[b]) < TB_BYTES(v)) {
          order[b + 1] = order[b];
          --b;
        }
        order[b + 1] = v;
      }
      for (int c = 0; c < n_cta; ++c) load[c] = 0.0;
      for (int q = 0; q < n_tiles; ++q) {
        const int t = order[q];
        if (!(TB_BYTES(t) > 0)) continue;
        int c = 0;
        for (int c2 = 1; c2 < n_cta; ++c2)
          if (load[c2] < load[c]) c = c2;
        load[c] += TB_BYTES(t);
        cta_of[t] = c;
      }
      for (int c = 0; c < n_cta; ++c) {
        cta_visit_ptr[c] = n_visits;
        for (int q = 0; q < n_tiles; ++q) {
          const int t = order[q];
          if (cta_of[t] == c) {
            pt[t].nc = 1;
            pt[t].visit0 = n_visits;
            visit_tile[n_visits++] = t;
          }
        }
      }
      cta_visit_ptr[n_cta] = n_visits;
    }
  }
#undef TB_BYTES
  // pieces per warp and the slot ranges of every (visit, warp)
  for (int t = 0; t < n_tiles; ++t) {
    if (pt[t].nc == 0) continue;
    const long long W = (long long)pt[t].nc * warps;
    const long long tg = tlen_g[t], tu = tlen_u ? tlen_u[t] : 0;
    long long sg = (tg + W * target_g - 1) / (W * target_g), su = (tu + W * target_u - 1) / (W * target_u);
    pt[t].segs[0] = tg > 0 ? (int)(sg < 1 ? 1 : sg) : 0;
    pt[t].segs[1] = tu > 0 ? (int)(su < 1 ? 1 : su) : 0;
  }
  int ng = 0, nu = 0;
  for (int v = 0; v < n_visits; ++v) {
    const int t = visit_tile[v];
    for (int w = 0; w < warps; ++w) {
      gp_ptr[v * warps + w] = ng;
      up_ptr[v * warps + w] = nu;
      ng += pt[t].segs[0];
      nu += pt[t].segs[1];
    }
  }
  gp_ptr[n_visits * warps] = ng;
  up_ptr[n_visits * warps] = nu;
  counts[0] = n_visits;
  counts[1] = ng;
  counts[2] = nu;
}

// piece index of an item inside its tile's stream (P pieces over `total` entries)
TB_HD long long tb_piece_of(long long cum, int len, long long P, long long total) {
  long long pc = (long long)(((double)cum + 0.5 * (double)len) * (double)P / (double)total);
  return pc < P - 1 ? pc : P - 1;
}

// one call per item: an item that starts a piece writes the piece's first group / first item into
// its slot and the END of the previous piece into that piece's slot (.w); the last item of a
// tile closes its own piece.  A second pass turns (.x, .w) into the group count .y.
TB_HD void tb_plan_piece(long long i, int cls, const ItemDesc* items, const long long* item_cum, const int* item_tile,
                         const int* row_item_ptr, long long rows_out, const long long* tlen, const PlanTile* pt, int warps,
                         const int* ptr, int group, int4* slots) {
  const int t = item_tile[i];
  const PlanTile T = pt[t];
  if (T.nc == 0) return;
  const long long total = tlen[t];
  const long long W = (long long)T.nc * warps, P = W * T.segs[cls];
  const int f = row_item_ptr[(long long)t * (rows_out + 1)], e = row_item_ptr[(long long)t * (rows_out + 1) + rows_out];
  const ItemDesc d = items[i];
  const long long pc = tb_piece_of(item_cum[i], d.len, P, total);
  long long pp = -1;
  if (i > f) pp = tb_piece_of(item_cum[i - 1], items[i - 1].len, P, total);
  if (pc != pp) {
    const long long wg = pc % W;
    const int j = ptr[(T.visit0 + (int)(wg / warps)) * warps + (int)(wg % warps)] + (int)(pc / W);
    slots[j].x = (int)(d.start / group);
    slots[j].z = (int)i;
    if (pp >= 0) {
      const long long wq = pp % W;
      const int j2 = ptr[(T.visit0 + (int)(wq / warps)) * warps + (int)(wq % warps)] + (int)(pp / W);
      slots[j2].w = (int)(d.start / group);
    }
  }
  if (i == e - 1) {
    const long long wg = pc % W;
    const int j = ptr[(T.visit0 + (int)(wg / warps)) * warps + (int)(wg % warps)] + (int)(pc / W);
    slots[j].w = (int)((d.start + d.len) / group);
  }
}
