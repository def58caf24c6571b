// Internal declarations shared by the translation units of libm3b.so.
// Nothing here is part of the ABI (include/m3b.h is).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string>
#include <vector>

#include "../../include/m3b.h"

// ---------------------------------------------------------------------------
// error plumbing: every CUDA call goes through CK(); the first failure is
// recorded in the context and turned into a status code at the ABI.
// ---------------------------------------------------------------------------
struct m3b_ctx;
void m3b_set_error(m3b_ctx* ctx, const char* fmt, ...);

#define CK(ctx, call)                                                                   \
  do {                                                                                  \
    cudaError_t e__ = (call);                                                           \
    if (e__ != cudaSuccess) {                                                           \
      m3b_set_error((ctx), "%s:%d: %s -> %s", __FILE__, __LINE__, #call,                \
                    cudaGetErrorString(e__));                                           \
      return (e__ == cudaErrorMemoryAllocation) ? M3B_E_NOMEM : M3B_E_CUDA;             \
    }                                                                                   \
  } while (0)

#define RET(expr)                  \
  do {                             \
    int s__ = (expr);              \
    if (s__ != M3B_OK) return s__; \
  } while (0)

// ---------------------------------------------------------------------------
// NCCL (loaded with dlopen so that a host process that already carries its own
// NCCL, e.g. torch, shares that copy instead of clashing with a second one)
// ---------------------------------------------------------------------------
struct NcclApi;
#define M3B_XCHG_CAP 131072      // doubles per exchange slot (kept genes <= 131072 for the fused path)
#define M3B_XCHG_FLAG_STRIDE 16  // doubles (128 B) between two flags
// flag-in-data ("LL") areas behind the two slots, used by the fused vector kernel (k_orth_fused):
// 16-byte entries {lo32, flag, hi32, flag}; per parity: [world][CAPF] gene-vector rows,
// [world][CAPW] Gram-Schmidt coefficients, [world][CAPG][2] per-CTA {|w|^2, sum w} partials
#define M3B_LL_CAPF 65536
#define M3B_LL_CAPW 512
#define M3B_LL_CAPG 256
#define M3B_LL_ENTRIES(world) ((size_t)(world) * (M3B_LL_CAPF + M3B_LL_CAPW + 2 * M3B_LL_CAPG))
struct Comm {
  int rank = 0, world = 1;
  void* comm = nullptr;  // ncclComm_t
  // peer-memory exchange (one box, NVLink/NVSwitch): every rank owns one buffer
  //   [flags: world x 2 slots, 128 B apart][slot 0: CAP doubles][slot 1: CAP doubles]
  // mapped into every other rank with CUDA IPC.  See k_combine_xchg_fix_F (irlba.cu).
  double* xchg = nullptr;          // own buffer
  double** peers_dev = nullptr;    // [world] device array of the ranks' buffers (own included)
  void* peer_base[16] = {nullptr}; // opened IPC mappings (to close)
  unsigned long long xchg_seq = 0; // use counter (same on every rank)
  unsigned int ll_launch[2] = {0, 0};  // fused launches so far, gene side / cell side (parity of the LL areas)
  bool p2p = false;
};
int comm_get_unique_id(m3b_ctx* ctx, char* id);
int comm_init(m3b_ctx* ctx, int rank, int world, const char* id);
void comm_destroy(m3b_ctx* ctx);
// in-place sum-allreduce of n doubles / int64 on the context stream (no-op if world==1)
int comm_allreduce_f64(m3b_ctx* ctx, double* dev, size_t n);
int comm_allreduce_i64(m3b_ctx* ctx, long long* dev, size_t n);
int comm_allreduce_max_i64(m3b_ctx* ctx, long long* dev, size_t n);
// all ranks agree on the next exchange sequence number (max over the ranks' counters + 2); a no-op when unsharded
int comm_xchg_resync(m3b_ctx* ctx);
int comm_p2p_handle(m3b_ctx* ctx, char* out64);
int comm_p2p_open(m3b_ctx* ctx, const char* handles /* world x 64 bytes */);

#define M3B_TILE_W_MAX 28672  // 224 KB of doubles: the most a CTA can stage
// Default tile width.  The full 224 KB leaves the SM no L1: the streaming loads of the
// product kernel are then throttled for lack of line buffers (ncu: lg_throttle, DRAM 56% vs 75%
// for the same kernel with a 120 KB tile; profiles/).  160 KB tiles leave ~60 KB of L1 (sweep in DESIGN.md).
#define M3B_TILE_W_DEFAULT 20480
// SMs left to the restart SVD while the next cycle's first product runs on the side stream
#define M3B_SIDE_CTA_RESERVE 28

// ---------------------------------------------------------------------------
// context
// ---------------------------------------------------------------------------
struct m3b_ctx {
  int device = 0;
  int flags = 0;
  int sm_count = 148;
  size_t smem_optin = 0;
  int tile_w_max = M3B_TILE_W_DEFAULT;  // columns per shared-memory tile of the dense operand
  cudaStream_t stream = nullptr;
  cudaStream_t stream2 = nullptr;  // side stream: next cycle's first product runs under the restart SVD
  std::string err;
  Comm comm;
  m3b_small_svd_fn small_svd = nullptr;
  void* small_svd_user = nullptr;
  m3b_should_abort_fn should_abort = nullptr;
  void* should_abort_user = nullptr;
  m3b_rnorm_fn rnorm = nullptr;
  void* rnorm_user = nullptr;
  m3b_stats stats{};
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;  // m3b_timer_start / m3b_timer_stop
  cudaEvent_t evk0 = nullptr, evk1 = nullptr;  // bracket the last K1 / K2 kernel
  // M3B_FLAG_TIME_SPMV: event pairs recorded around every SpMV launch, read back after the run
  std::vector<cudaEvent_t> ev_pool;
  std::vector<int> ev_kind;  // 0 = cells-out (K4), 1 = genes-out (K5)
  size_t ev_used = 0;
  bool k1_timed = false, k2_timed = false;   // ev0/ev1 bracket the last K1 / K2 launch
  bool k10_timed = false;                    // ... the K10 (projection) kernel
  unsigned int* sched = nullptr;             // [2] dynamic unit counter + done counter (zeroed)
  int clock_khz = 1900000;                        // nominal SM clock (m3b_create)
  long long spin_timeout_cycles = 60000000000LL;  // bound of the in-kernel waits (m3b_create; M3B_XCHG_TIMEOUT_S)
  int* ingest_flag = nullptr;                // device word raised by the ingest validation kernels
};

// ---------------------------------------------------------------------------
// tiled row-major sparse layout: the single storage format behind both
// Lanczos products (see DESIGN.md "Data layout in HBM").
//   rows x cols matrix; the gathered (cols) dimension is cut into tiles of
//   tile_w columns so that the dense operand of one tile fits in shared memory;
//   inside a tile the non-zeros are stored row by row, every run padded to a
//   multiple of 4 entries (pad entries: idx 0, val 0.0).  Because the index is
//   tile-local it fits 16 bits, so the stream is 10 B/entry instead of the
//   12 B/entry of an int32-indexed CSR.  An ITEM is one run (or
//   a <= M3B_ITEM_MAX slice of a long run): the unit of work of one warp.
// ---------------------------------------------------------------------------
#define M3B_ITEM_MAX 2048     // entries per item (multiple of 128)
#define M3B_ITEM_TAIL 512     // entries per item in the last part of a tile (short tail of the launch)
// UNIT layouts (see "unit entries" below): index-only streams, 2 B/entry, runs padded to 8
#define M3B_UITEM_MAX 4096
#define M3B_UITEM_TAIL 1024

struct ItemDesc {  // 16 bytes, loaded with one 128-bit load
  long long start;  // entry offset into idx/val (multiple of 4)
  int len;          // padded entries (multiple of 4)
  int row;          // output row
};

struct TiledLayout {
  long long rows = 0, cols = 0;
  int tile_w = 0, n_tiles = 0;
  long long n_entries = 0;  // padded entries allocated
  long long nnz = 0;        // true stored entries referenced by items
  unsigned short* idx = nullptr;  // tile-local column index (tile_w <= 28672 < 2^16: 2 bytes/entry)
  double* val = nullptr;
  long long n_items = 0;
  ItemDesc* items = nullptr;
  // scheduling: a CHUNK is a short contiguous item range (~2048 entries, ~512 near the end of
  // a tile) that one warp claims with an atomic on its tile's counter
  int n_chunks = 0;
  int2* chunks = nullptr;          // {item_begin, item_end}, grouped by tile
  int* tile_chunk_ptr = nullptr;   // [n_tiles+1] chunk ranges per tile
  int* cta_tile = nullptr;         // [n_cta] tile a CTA starts on (CTAs split over tiles by entries)
  int n_cta = 0;
  unsigned int* counters = nullptr;  // [n_tiles+1]: per-tile chunk cursor, last = CTAs finished
  int* row_item_ptr = nullptr; // [n_tiles][rows+1] item ranges per (tile,row)
  // The partial of item i is written to slot item_slot[i]; the slots of a row are contiguous
  // (tile, then segment order): out[row] = sum of partial[row_slot_ptr[row] .. row_slot_ptr[row+1]).
  int* item_slot = nullptr;    // [n_items]
  int* row_slot_ptr = nullptr; // [rows+1]
  double* partial = nullptr;   // [n_items], indexed by slot
  // UNIT ENTRIES.  Most stored counts are exactly 1 (51-84% in the reference's fixtures and the
  // synthetic benchmark matrix), and the normalised value of a count-1 entry depends on its cell
  // only: u_c = f(1 / size_factor_c).  Those entries are kept in a sibling layout that stores NO
  // values -- a 2 B/entry index stream -- and are multiplied by u on the dense side:
  //   layout A (row = cell):  partial = u[row] * sum x[idx]        (rowscale)
  //   layout B (col = cell):  the staged tile holds x[c] * u[c]     (colscale)
  // The sibling shares the parent's partial array and slot numbering (a row's slots are its
  // general items followed by its unit items), so the combine kernels see one layout.
  TiledLayout* unit = nullptr;       // owned; nullptr when the matrix is not split
  bool is_unit = false;              // val == nullptr, runs padded to 8 with idx == tile_w (a zero slot)
  const double* colscale = nullptr;  // unit layouts: per-column factor (length cols) or nullptr
  const double* rowscale = nullptr;  // unit layouts: per-row factor (length rows) or nullptr
  // static work plans of the flat product kernel (flat.cu); [0] = whole grid, [1] = the reduced
  // grid of the side-stream launch (single-tile layouts only).  Owned by the general layout.
  struct FlatPlan* plan[2] = {nullptr, nullptr};
  double* item_scale = nullptr;  // unit layouts with a rowscale: rowscale[row of item], per item
  // device-built tables only (tables.cu): kept entries per tile, tile and in-tile position of every item
  long long* tile_len = nullptr;  // [n_tiles]
  int* item_tile = nullptr;       // [n_items]
  long long* item_cum = nullptr;  // [n_items]
  void free_all();
};

// host-side tables of a layout while it is being built (layout.cu), input of the flat plan (flat.cu)
struct HostTables {
  std::vector<long long> base;
  std::vector<ItemDesc> items;
  std::vector<int> row_item_ptr, item_slot, row_slot_ptr;
  std::vector<int2> chunks;
  std::vector<int> tile_chunk_ptr, cta_tile;
  long long n_entries = 0, nnz = 0;
};

// the matrix handle
struct m3b_mat {
  m3b_ctx* ctx = nullptr;
  int n_genes = 0;
  long long n_cells = 0;        // local
  long long n_cells_total = 0;  // over all ranks
  // raw CSC (cell-major) as appended
  long long nnz = 0, nnz_cap = 0, cells_appended = 0;
  long long* p = nullptr;  // [n_cells+1]
  int* i = nullptr;        // [nnz]
  double* x = nullptr;     // [nnz] raw counts
  double* y = nullptr;     // [nnz] normalised values (after m3b_normalize)
  double* unit_val = nullptr;  // [n_cells] normalised value of a count-1 entry of the cell
  bool split = false;          // the layouts keep count-1 entries in index-only siblings
  std::vector<unsigned char> keep_mask;  // zero-gene filter of the last selection (per selected gene)
  bool ended = false, normalized = false, selected = false, moments_done = false;
  bool tfidf_applied = false;  // the layouts hold TF-IDF values (LSI branch), not the normalised ones
  bool sel_monotone = true;    // kept index grows with the gene id (no use_genes re-ordering)
  bool layouts_built = false;  // A / B exist (m3b_select_genes); false after m3b_select_genes_for_transform
  int norm_method = -1;
  double pseudo_count = 0, f0 = 0;  // f0 = implicit value of the non-stored entries
  // gene selection
  int n_sel = 0, n_kept = 0;
  int* sel_of_gene = nullptr;   // [n_genes] -> selected position or -1
  int* kept_of_sel = nullptr;   // [n_sel]   -> kept index or -1
  int* kept_of_gene = nullptr;  // [n_genes] -> kept index or -1
  long long* sel_nnz = nullptr; // [n_sel] exact local nnz per selected gene
  double* rowsum = nullptr;     // [n_sel] global row sums (incl. f0 term)
  long long* kept_nnz = nullptr;// [n_kept] local nnz
  long long* kept_padlen = nullptr;  // [n_kept] local padded run length (>= nnz)
  double* mean_sparse = nullptr;     // [n_kept] mean of the stored sparse part (center - f0)
  void* b_host_tables = nullptr;  // host copy of B's run counts between build and re-map (M3B_HOST_TABLES=1)
  // device-side counterpart: per-(tile, selected row) entry counts and run offsets of B, general / unit
  int* b_cnt[2] = {nullptr, nullptr};
  long long* b_base[2] = {nullptr, nullptr};
  TiledLayout A;  // rows = local cells, cols = kept genes   (cells-out product, K4)
  TiledLayout B;  // rows = kept genes,  cols = local cells  (genes-out product, K5)
  double* center = nullptr;  // [n_kept] mean of the (implicit-dense) matrix
  double* scale = nullptr;   // [n_kept] sd
  // embedding left on the device by the last PCA (n_cells x nv, ld n_cells)
  double* Xdev = nullptr;
  int Xdev_nv = 0;
};

// ---------------------------------------------------------------------------
// kernels / device-side building blocks (implemented in the .cu files)
// ---------------------------------------------------------------------------
template <typename T>
int dev_alloc(m3b_ctx* ctx, T** p, size_t n);
void dev_free(void* p);
void dev_pool_release(int device);  // return the cached blocks of a device to the driver
// Frees the registered device blocks when the scope is left -- on EVERY path: CK() / RET() return
// early, and a block that is not given back stays in the pool's live map for good.  Register the
// pointer VARIABLE before allocating into it; a block handed on to the caller is taken out by
// setting the variable to nullptr.
struct DevScope {
  std::vector<void**> slots;
  template <typename T>
  void own(T** p) {
    slots.push_back(reinterpret_cast<void**>(p));
  }
  ~DevScope() {
    for (void** q : slots) {
      dev_free(*q);
      *q = nullptr;
    }
  }
};

// layout.cu
int k1_cell_totals(m3b_mat* m, int round_exprs, double* d_tot);
int k2_normalize(m3b_mat* m, const double* d_sf, int method, double pc);
int build_layout_B(m3b_mat* m, bool split);
int finish_selection(m3b_mat* m, unsigned char* keep_host);
int build_layout_A(m3b_mat* m, bool split);
int rebuild_layouts(m3b_mat* m, bool split);  // same selection and keep mask, other split mode
void m3b_mat_set_bbuild(m3b_mat* m, void* p);
int ingest_append(m3b_mat* m, long long n_cells_chunk, const int* p, const int* i, const double* x);

// spmv.cu
enum StreamOp { OP_DOT = 0, OP_SUM = 1, OP_SQDEV = 2, OP_COUNTPOS = 3 };
// partial[item] = sum over the item's entries of op(val, x[idx]); x is the dense operand
// (length L.cols) for OP_DOT, ignored for OP_SUM, the per-row mean for OP_SQDEV.
int launch_tile_stream(m3b_ctx* ctx, const TiledLayout& L, int op, const double* x, cudaStream_t stream = nullptr,
                       int max_cta = 0);
// out[row] = sum of the row's partials (fixed order: tile, then segment)
int launch_combine_rows(m3b_ctx* ctx, const TiledLayout& L, double* out);
long long spmv_algorithmic_bytes(const TiledLayout& L);
long long spmv_stream_bytes(const TiledLayout& L);

// flat.cu: the product kernel proper (OP_DOT) -- static, item-aligned partition of each tile's
// entry stream over the warps, streamed flat with a register ring and reduced per item
void flat_plan_free(struct FlatPlan* p);
int flat_build_plans(m3b_ctx* ctx, TiledLayout& L, const HostTables& Tg, const HostTables* Tu);
int flat_launch_dot(m3b_ctx* ctx, const TiledLayout& L, const double* x, cudaStream_t stream, int max_cta);
int flat_init(m3b_ctx* ctx);
bool flat_has_plan(const TiledLayout& L, int n_cta);  // a plan for exactly n_cta CTAs exists (side-stream launch)
int flat_build_plans_device(m3b_ctx* ctx, TiledLayout& L);

// tables.cu: the layout tables and flat plans built on the device (default; M3B_HOST_TABLES=1 keeps
// the host builders of layout.cu / flat.cu)
int tables_compute_bases(m3b_ctx* ctx, int n_tiles, long long rows_in, const int* const cnt[2], int n_cls,
                         long long* const base[2], long long n_entries[2]);
int tables_build(m3b_ctx* ctx, TiledLayout& L, long long rows_in, const int* row_map, const int* const cnt[2],
                 const long long* const base[2]);
int tables_neutralize_dropped(m3b_ctx* ctx, TiledLayout& L, long long rows_in, const int* row_map, const int* const cnt[2],
                              const long long* const base[2]);
int tables_row_stats(m3b_ctx* ctx, int n_tiles, long long rows_in, const int* row_map, const int* cnt_g, const int* cnt_u,
                     long long* nnz_out, long long* padlen_out);
int tables_build_plan(m3b_ctx* ctx, const TiledLayout& L, int n_cta, int warps, long long target_g, long long target_u,
                      int** cta_visit_ptr, int** visit_tile, int** gp_ptr, int4** gp, int** up_ptr, int4** up, int* n_visits,
                      long long* n_pieces);
static inline bool use_host_tables() { return getenv("M3B_HOST_TABLES") != nullptr; }

// aggregate.cu
int run_detect_genes(m3b_mat* m, double min_expr, long long* num_cells_expressed, long long* num_genes_expressed);
int run_aggregate(m3b_mat* m, const int* gene_group, int n_groups, int gene_mean, const int* cell_group, int n_cell_groups,
                  int cell_mean, double* out, long long ld_out);

// jaccard.cu
int run_jaccard(m3b_ctx* ctx, const double* idx, long long n, int k, int weight, double* out);

// knn.cu
int run_knn_self(m3b_ctx* ctx, const double* X, long long n, int d, int k, int* idx_out, double* dist_out);

// align.cu
int run_align_residuals(m3b_ctx* ctx, double* X, long long n, int nv, long long ldx, const double* M, int p, double* beta,
                        bool fit);

// tfidf.cu (LSI branch)
int run_tfidf(m3b_mat* m, int frequencies, int log_scale_tf, double scale_factor, const double* row_sums_in,
              double num_cols, double* col_sums_out, double* row_sums_out);

// irlba.cu
int run_irlba(m3b_mat* m, int nv, int work, int maxit, double tol, double svtol, const double* v0,
              int center, int scale, double* d, double* V, double* X, long long ldx,
              double* center_out, double* scale_out, int* iter, int* mprod, bool copy_x);
int compute_moments(m3b_mat* m);
// tiny.cu: irlba's dense branch for min(cells, genes) < 6
int run_irlba_tiny(m3b_mat* m, int nv, int center, int scale, double* d, double* V, double* X, long long ldx, double* center_out,
                   double* scale_out, int* iter, int* mprod, bool copy_x);

// project.cu
int run_select_for_transform(m3b_mat* m, const int* gene_order, int n_sel, unsigned char* keep, int* n_kept, int* fallback);
int run_project(m3b_mat* q, const int* model_row, int n_model, const double* V, const double* center,
                const double* scale, int nv, double* Y, long long ldy);

// small_svd.cpp
int jacobi_svd(int n, const double* B, double* U, double* s, double* Vt);

static inline void count_launch(m3b_ctx* ctx, int n = 1) { ctx->stats.kernel_launches += n; }
int spmv_init(m3b_ctx* ctx);
int svd_init(m3b_ctx* ctx);
int irlba_init(m3b_ctx* ctx);
// event-pool helpers (api.cu)
void spmv_timer_begin(m3b_ctx* ctx, int kind, cudaStream_t stream = nullptr);
void spmv_timer_end(m3b_ctx* ctx, cudaStream_t stream = nullptr);
void spmv_timer_collect(m3b_ctx* ctx);
