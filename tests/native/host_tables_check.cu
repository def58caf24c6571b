// Host check of the device-side table builder (tables.cu / tables_body.h) against the host
// builders it replaces (layout.cu: compute_bases, build_items, assign_slots; flat.cu: plan_host).
// The per-element bodies are __host__ __device__: this program runs them in plain loops (prefix
// sums with a serial scan) on random run-count tables and compares items, slots, row pointers and
// flat-plan pieces with the host builders' output.  No GPU needed.  Built and run by
// tests/test_tables_host.py:  nvcc -std=c++17 -I monocle3_b200/csrc tests/native/host_tables_check.cu
#include <cstdio>
#include <cstdlib>
#include <numeric>
#include <random>
#include <vector>

#include "layout.cu"   // static host builders (the CUDA kernels in there are compiled but never launched)
#include "flat.cu"
#include "tables_body.h"

// stubs for symbols the included sources reference but this program never calls
int launch_tile_stream(m3b_ctx*, const TiledLayout&, int, const double*, cudaStream_t, int) { return 0; }
int launch_combine_rows(m3b_ctx*, const TiledLayout&, double*) { return 0; }
int comm_allreduce_f64(m3b_ctx*, double*, size_t) { return 0; }
void m3b_set_error(m3b_ctx*, const char*, ...) {}
int tables_compute_bases(m3b_ctx*, int, long long, const int* const*, int, long long* const*, long long*) { return 0; }
int tables_build(m3b_ctx*, TiledLayout&, long long, const int*, const int* const*, const long long* const*) { return 0; }
int tables_neutralize_dropped(m3b_ctx*, TiledLayout&, long long, const int*, const int* const*, const long long* const*) { return 0; }
int tables_row_stats(m3b_ctx*, int, long long, const int*, const int*, const int*, long long*, long long*) { return 0; }
int tables_build_plan(m3b_ctx*, const TiledLayout&, int, int, long long, long long, int**, int**, int**, int4**, int**, int4**, int*,
                      long long*) { return 0; }

template <typename Tin, typename Tout>
static std::vector<Tout> excl(const std::vector<Tin>& in) {
  std::vector<Tout> out(in.size() + 1);
  Tout s = 0;
  for (size_t k = 0; k < in.size(); ++k) {
    out[k] = s;
    s += (Tout)in[k];
  }
  out[in.size()] = s;
  return out;
}

struct DevTab {  // what tables_build produces for one class, emulated on the host
  std::vector<long long> base, kcum, tile_len, item_cum;
  std::vector<int> nit, ioff, sbase, item_slot, item_tile, item_w, rip;
  std::vector<ItemDesc> items;
};

static int fails = 0;
#define CHECK(cond, ...)            \
  do {                              \
    if (!(cond)) {                  \
      if (fails < 20) {             \
        printf("FAIL %s:%d: ", __FILE__, __LINE__); \
        printf(__VA_ARGS__);        \
        printf("\n");               \
      }                             \
      ++fails;                      \
    }                               \
  } while (0)

static void run_case(int n_tiles, long long rows_in, double drop_frac, double empty_frac, int max_cnt, bool split, int n_cta, int warps,
                     long long target_g, long long target_u, unsigned seed) {
  std::mt19937 rng(seed);
  const long long n_runs = (long long)n_tiles * rows_in;
  std::vector<int> cnt[2];
  for (int q = 0; q < 2; ++q) {
    cnt[q].assign(n_runs, 0);
    for (auto& c : cnt[q]) {
      const double u = (rng() % 10000) / 10000.0;
      if (u < empty_frac) c = 0;
      else if (u < 0.9) c = 1 + rng() % 40;
      else c = 1 + rng() % max_cnt;
    }
  }
  std::vector<int> row_map(rows_in);
  long long rows_out = 0;
  bool any_drop = false;
  for (long long r = 0; r < rows_in; ++r) {
    const bool drop = ((rng() % 10000) / 10000.0) < drop_frac;
    row_map[r] = drop ? -1 : (int)rows_out++;
    any_drop |= drop;
  }
  const int* rm = any_drop ? row_map.data() : nullptr;
  const int n_cls = split ? 2 : 1;
  const ItemShape hs[2] = {SHAPE_GENERAL, SHAPE_UNIT};
  const TabShape ds[2] = {{4, M3B_ITEM_MAX, M3B_ITEM_TAIL, 2}, {8, M3B_UITEM_MAX, M3B_UITEM_TAIL, 8}};
  // ---- host builders ----
  HostTables T[2];
  for (int q = 0; q < n_cls; ++q) {
    compute_bases(cnt[q], n_tiles, rows_in, T[q], hs[q]);
    build_items(cnt[q], n_tiles, rows_in, rm, rows_out, n_cta, T[q], hs[q]);
  }
  assign_slots(T[0], split ? &T[1] : nullptr, rows_out);
  // ---- device bodies, emulated ----
  DevTab D[2];
  for (int q = 0; q < n_cls; ++q) {
    std::vector<long long> pa(n_runs), pk(n_runs);
    long long nz;
    for (long long run = 0; run < n_runs; ++run) tb_run_plen(run, cnt[q].data(), rm, rows_in, ds[q].pad, pa.data(), pk.data(), &nz);
    D[q].base = excl<long long, long long>(pa);
    D[q].kcum = excl<long long, long long>(pk);
    D[q].nit.resize(n_runs);
    for (long long run = 0; run < n_runs; ++run) D[q].nit[run] = tb_run_items(run, cnt[q].data(), D[q].kcum.data(), rows_in, ds[q]);
    D[q].ioff = excl<int, int>(D[q].nit);
    D[q].tile_len.resize(n_tiles);
    for (int t = 0; t < n_tiles; ++t) D[q].tile_len[t] = D[q].kcum[(long long)(t + 1) * rows_in] - D[q].kcum[(long long)t * rows_in];
    D[q].sbase.assign(n_runs, 0);
  }
  std::vector<int> rowtot(rows_out, -1), gtot(rows_out, -1);
  for (long long r = 0; r < rows_in; ++r)
    tb_row_slots(r, rm, rows_in, n_tiles, D[0].nit.data(), split ? D[1].nit.data() : nullptr, D[0].sbase.data(),
                 split ? D[1].sbase.data() : nullptr, rowtot.data(), gtot.data());
  std::vector<int> rsp = excl<int, int>(rowtot);
  for (int q = 0; q < n_cls; ++q) {
    const int ni = D[q].ioff[n_runs];
    D[q].items.resize(ni);
    D[q].item_slot.assign(ni, -1);
    D[q].item_cum.resize(ni);
    D[q].item_tile.resize(ni);
    D[q].item_w.resize(ni);
    D[q].rip.assign((size_t)n_tiles * (rows_out + 1), -7);
    TabItemOut o{D[q].items.data(), D[q].item_slot.data(), D[q].item_cum.data(), D[q].item_tile.data(), D[q].item_w.data(), D[q].rip.data()};
    for (long long run = 0; run < n_runs; ++run)
      tb_fill_items(run, cnt[q].data(), rm, rows_in, rows_out, D[q].base.data(), D[q].kcum.data(), D[q].ioff.data(), D[q].sbase.data(),
                    rsp.data(), q ? gtot.data() : nullptr, ds[q], o);
    // compare with the host tables
    CHECK(D[q].base[n_runs] == T[q].n_entries, "n_entries %lld vs %lld", D[q].base[n_runs], T[q].n_entries);
    for (long long run = 0; run < n_runs; ++run) CHECK(D[q].base[run] == T[q].base[run], "base[%lld]", run);
    CHECK((size_t)ni == T[q].items.size(), "class %d: n_items %d vs %zu", q, ni, T[q].items.size());
    if ((size_t)ni == T[q].items.size()) {
      for (int i = 0; i < ni; ++i) {
        const ItemDesc &a = D[q].items[i], &b = T[q].items[i];
        CHECK(a.start == b.start && a.len == b.len && a.row == b.row, "class %d item %d: (%lld,%d,%d) vs (%lld,%d,%d)", q, i, a.start,
              a.len, a.row, b.start, b.len, b.row);
        CHECK(D[q].item_slot[i] == T[q].item_slot[i], "class %d slot %d: %d vs %d", q, i, D[q].item_slot[i], T[q].item_slot[i]);
      }
    }
    CHECK(D[q].rip.size() == T[q].row_item_ptr.size(), "rip size");
    for (size_t k = 0; k < D[q].rip.size(); ++k) CHECK(D[q].rip[k] == T[q].row_item_ptr[k], "class %d rip[%zu] %d vs %d", q, k, D[q].rip[k], T[q].row_item_ptr[k]);
    // chunks: every item in exactly one chunk, chunks inside tiles
    std::vector<long long> wcum = excl<int, long long>(D[q].item_w);
    std::vector<int> head(ni);
    for (int i = 0; i < ni; ++i) head[i] = tb_chunk_head(i, wcum.data(), D[q].item_tile.data(), D[q].rip.data(), rows_out, ds[q].item_max);
    for (int t = 0; t < n_tiles; ++t) {
      const int f = D[q].rip[(size_t)t * (rows_out + 1)], e = D[q].rip[(size_t)t * (rows_out + 1) + rows_out];
      if (f < e) CHECK(head[f] == 1, "tile %d first item is not a chunk head", t);
    }
  }
  for (long long ro = 0; ro <= rows_out; ++ro) CHECK(rsp[ro] == T[0].row_slot_ptr[ro], "rsp[%lld] %d vs %d", ro, rsp[ro], T[0].row_slot_ptr[ro]);
  if (getenv("TB_VERBOSE")) printf("case seed %u: tiles %d rows %lld->%lld items %zu/%zu entries %lld/%lld\n", seed, n_tiles, rows_in, rows_out, D[0].items.size(), D[1].items.size(), D[0].base[n_runs], split ? D[1].base[n_runs] : 0LL);
  // ---- flat plan ----
  // the host builder breaks pieces at memory gaps (dropped rows with entries); the device builder
  // turns such entries into pads and never breaks, so compare only when nothing with entries was dropped
  bool gaps = false;
  if (rm)
    for (long long run = 0; run < n_runs; ++run)
      if (row_map[run % rows_in] < 0 && (cnt[0][run] > 0 || (split && cnt[1][run] > 0))) gaps = true;
  HostPlan HP;
  plan_host(n_tiles, rows_out, T[0], split ? &T[1] : nullptr, n_cta, warps, target_g, target_u, HP);
  const int max_visits = std::max(n_cta, n_tiles);
  std::vector<PlanTile> pt(n_tiles);
  std::vector<int> cvp(n_cta + 1), vt(max_visits), gpp((size_t)max_visits * warps + 1), upp((size_t)max_visits * warps + 1), counts(4),
      n_of(n_tiles), order(n_tiles), cta_of(n_tiles);
  std::vector<double> load(n_cta);
  tb_plan_meta(n_tiles, n_cta, warps, D[0].tile_len.data(), split ? D[1].tile_len.data() : nullptr, target_g, target_u, pt.data(),
               cvp.data(), vt.data(), gpp.data(), upp.data(), counts.data(), n_of.data(), order.data(), load.data(), cta_of.data());
  CHECK((size_t)counts[0] == HP.visit_tile.size(), "n_visits %d vs %zu", counts[0], HP.visit_tile.size());
  for (int c = 0; c <= n_cta; ++c) CHECK(cvp[c] == HP.cta_visit_ptr[c], "cta_visit_ptr[%d] %d vs %d", c, cvp[c], HP.cta_visit_ptr[c]);
  for (size_t v = 0; v < HP.visit_tile.size() && v < (size_t)counts[0]; ++v) CHECK(vt[v] == HP.visit_tile[v], "visit_tile[%zu]", v);
  for (int q = 0; q < n_cls; ++q) {
    const long long cap = D[q].base[n_runs] / (q ? target_u : target_g) + (long long)max_visits * warps + 1;
    CHECK(counts[1 + q] <= cap, "slot bound %d > %lld", counts[1 + q], cap);
    std::vector<int4> slots(cap, make_int4(0, 0, 0, 0));
    const std::vector<int>& ptr = q ? upp : gpp;
    for (int i = 0; i < (int)D[q].items.size(); ++i)
      tb_plan_piece(i, q, D[q].items.data(), D[q].item_cum.data(), D[q].item_tile.data(), D[q].rip.data(), rows_out, D[q].tile_len.data(),
                    pt.data(), warps, ptr.data(), ds[q].group, slots.data());
    for (auto& s : slots) s.y = s.w > s.x ? s.w - s.x : 0;
    // every item covered exactly once by the pieces, in order, inside its tile
    {
      std::vector<int> covered(D[q].items.size(), 0);
      for (int v = 0; v < counts[0]; ++v)
        for (int w = 0; w < warps; ++w)
          for (int j = ptr[v * warps + w]; j < ptr[v * warps + w + 1]; ++j) {
            if (slots[j].y == 0) continue;
            long long g0 = slots[j].x, g1 = g0 + slots[j].y;
            int i = slots[j].z;
            CHECK(D[q].items[i].start / ds[q].group == g0, "piece does not start at its first item");
            CHECK(D[q].item_tile[i] == vt[v], "piece of tile %d in a visit of tile %d", D[q].item_tile[i], vt[v]);
            while (i < (int)D[q].items.size() && D[q].items[i].start / ds[q].group < g1 && D[q].item_tile[i] == vt[v]) {
              CHECK((D[q].items[i].start + D[q].items[i].len) / ds[q].group <= g1, "item %d cut by a piece end", i);
              covered[i]++;
              ++i;
            }
          }
      for (size_t i = 0; i < covered.size(); ++i) CHECK(covered[i] == 1, "class %d item %zu covered %d times", q, i, covered[i]);
    }
    if (!gaps) {
      const std::vector<int>& hptr = q ? HP.up_ptr : HP.gp_ptr;
      const std::vector<int4>& hp = q ? HP.up : HP.gp;
      for (int v = 0; v < counts[0] && v < (int)HP.visit_tile.size(); ++v)
        for (int w = 0; w < warps; ++w) {
          std::vector<int4> mine;
          for (int j = ptr[v * warps + w]; j < ptr[v * warps + w + 1]; ++j)
            if (slots[j].y > 0) mine.push_back(slots[j]);
          const int h0 = hptr[v * warps + w], h1 = hptr[v * warps + w + 1];
          CHECK((int)mine.size() == h1 - h0, "class %d (visit %d, warp %d): %zu pieces vs %d", q, v, w, mine.size(), h1 - h0);
          for (int k = 0; k < (int)mine.size() && k < h1 - h0; ++k)
            CHECK(mine[k].x == hp[h0 + k].x && mine[k].y == hp[h0 + k].y && mine[k].z == hp[h0 + k].z,
                  "class %d (visit %d, warp %d) piece %d: (%d,%d,%d) vs (%d,%d,%d)", q, v, w, k, mine[k].x, mine[k].y, mine[k].z,
                  hp[h0 + k].x, hp[h0 + k].y, hp[h0 + k].z);
        }
    }
  }
}

int main() {
  // (n_tiles, rows_in, drop, empty, max_cnt, split, n_cta, warps, target_g, target_u, seed)
  run_case(1, 500, 0.0, 0.1, 3000, true, 148, 32, 8192, 32768, 1);
  run_case(1, 3000, 0.0, 0.3, 9000, true, 120, 32, 8192, 32768, 2);
  run_case(3, 700, 0.0, 0.2, 5000, true, 148, 32, 8192, 32768, 3);
  run_case(13, 900, 0.0, 0.5, 20000, true, 148, 32, 8192, 32768, 4);
  run_case(2, 4000, 0.0, 0.05, 300, false, 148, 32, 8192, 32768, 5);
  run_case(5, 600, 0.15, 0.2, 4000, true, 148, 16, 1024, 4096, 6);   // dropped rows (with entries: gaps)
  run_case(4, 800, 0.2, 0.9, 2000, true, 148, 32, 8192, 32768, 7);
  run_case(200, 40, 0.0, 0.3, 2500, true, 148, 32, 8192, 32768, 8);  // more tiles than CTAs
  run_case(160, 30, 0.1, 0.3, 900, false, 120, 32, 2048, 8192, 9);
  run_case(1, 1, 0.0, 0.0, 10, true, 148, 32, 8192, 32768, 10);
  run_case(2, 50, 0.0, 1.0, 10, true, 148, 32, 8192, 32768, 11);     // all runs empty
  printf(fails ? "FAILED: %d mismatches\n" : "ok: device table bodies agree with the host builders\n", fails);
  return fails ? 1 : 0;
}
