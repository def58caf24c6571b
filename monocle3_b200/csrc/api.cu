// extern "C" entry points of libm3b.so (declared in include/m3b.h).
#include <math.h>
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <chrono>
#include <new>
#include <thread>
#include <vector>

#include "m3b_internal.h"

void m3b_set_error(m3b_ctx* ctx, const char* fmt, ...) {
  if (!ctx) return;
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  ctx->err = buf;
}

void spmv_timer_begin(m3b_ctx* ctx, int kind, cudaStream_t stream) {
  if (!(ctx->flags & M3B_FLAG_TIME_SPMV)) return;
  if (ctx->ev_used + 2 > ctx->ev_pool.size()) {
    for (int k = 0; k < 512; ++k) {
      cudaEvent_t e;
      if (cudaEventCreate(&e) != cudaSuccess) return;
      ctx->ev_pool.push_back(e);
    }
  }
  ctx->ev_kind.push_back(kind);
  cudaEventRecord(ctx->ev_pool[ctx->ev_used++], stream ? stream : ctx->stream);
}
void spmv_timer_end(m3b_ctx* ctx, cudaStream_t stream) {
  if (!(ctx->flags & M3B_FLAG_TIME_SPMV)) return;
  if (ctx->ev_used & 1) cudaEventRecord(ctx->ev_pool[ctx->ev_used++], stream ? stream : ctx->stream);
}
void spmv_timer_collect(m3b_ctx* ctx) {
  if (!(ctx->flags & M3B_FLAG_TIME_SPMV)) return;
  cudaStreamSynchronize(ctx->stream2);
  cudaStreamSynchronize(ctx->stream);
  for (size_t k = 0; k + 1 < ctx->ev_used; k += 2) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, ctx->ev_pool[k], ctx->ev_pool[k + 1]) != cudaSuccess) continue;
    const int kind = ctx->ev_kind[k / 2];
    if (kind == 0) ctx->stats.spmv_cells_out_ms += ms;
    if (kind == 1) ctx->stats.spmv_genes_out_ms += ms;
    if (kind >= 0 && kind < 8) ctx->stats.phase_ms[kind] += ms;
  }
  ctx->ev_used = 0;
  ctx->ev_kind.clear();
}

extern "C" {

int m3b_abi_version(void) { return M3B_ABI_VERSION; }

const char* m3b_status_string(int s) {
  switch (s) {
    case M3B_OK: return "ok";
    case M3B_E_DIMS: return "invalid dimensions";
    case M3B_W_NOT_CONVERGED: return "didn't converge";
    case M3B_E_NOMEM: return "out of memory";
    case M3B_E_NULLSPACE: return "starting vector near the null space";
    case M3B_E_LINDEP: return "linear dependency encountered";
    case M3B_E_CUDA: return "CUDA failure";
    case M3B_E_NCCL: return "NCCL failure";
    case M3B_E_STATE: return "call sequence violated";
    case M3B_E_NODEVICE: return "no usable CUDA device (there is no CPU fallback)";
    case M3B_E_ABORTED: return "aborted";
    default: return "unknown status";
  }
}

int m3b_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

int m3b_create(int device, int flags, m3b_ctx** out) {
  if (!out) return M3B_E_DIMS;
  *out = nullptr;
  int n = m3b_device_count();
  if (n <= 0 || device < 0 || device >= n) return M3B_E_NODEVICE;
  m3b_ctx* ctx = new (std::nothrow) m3b_ctx();
  if (!ctx) return M3B_E_NOMEM;
  ctx->device = device;
  ctx->flags = flags;
  cudaDeviceProp prop;
  if (cudaSetDevice(device) != cudaSuccess || cudaGetDeviceProperties(&prop, device) != cudaSuccess) {
    delete ctx;
    return M3B_E_NODEVICE;
  }
  if (prop.major < 10) {  // the kernels are built for sm_100a only
    delete ctx;
    return M3B_E_NODEVICE;
  }
  ctx->sm_count = prop.multiProcessorCount;
  ctx->smem_optin = prop.sharedMemPerBlockOptin;
  {
    // bound of the in-kernel waits (peer exchange, grid barrier): seconds -> SM cycles.  Generous by
    // default: ordinary rank skew (a host LAPACK callback, paging, a debugger) must not turn into a
    // failure; M3B_XCHG_TIMEOUT_S overrides.
    double secs = 30.0;
    if (const char* e = getenv("M3B_XCHG_TIMEOUT_S")) secs = std::max(0.1, atof(e));
    int khz = 1900000;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, device);
    ctx->clock_khz = khz;  // (queried ONCE: the attribute goes to the driver and can take milliseconds)
    ctx->spin_timeout_cycles = (long long)(secs * 1e3 * (double)khz);
  }
  if (const char* e = getenv("M3B_TILE_W")) {  // tuning knob (multiple of 32, <= 28672)
    int v = atoi(e);
    if (v >= 32 && v <= M3B_TILE_W_MAX) ctx->tile_w_max = v & ~31;
  }
  if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess ||
      cudaStreamCreateWithFlags(&ctx->stream2, cudaStreamNonBlocking) != cudaSuccess ||
      cudaEventCreate(&ctx->ev0) != cudaSuccess || cudaEventCreate(&ctx->ev1) != cudaSuccess ||
      cudaEventCreate(&ctx->evk0) != cudaSuccess || cudaEventCreate(&ctx->evk1) != cudaSuccess ||
      cudaMalloc((void**)&ctx->sched, 4 * sizeof(unsigned int)) != cudaSuccess ||
      cudaMemset(ctx->sched, 0, 4 * sizeof(unsigned int)) != cudaSuccess) {
    m3b_destroy(ctx);
    return M3B_E_CUDA;
  }
  if (spmv_init(ctx) != M3B_OK || svd_init(ctx) != M3B_OK || irlba_init(ctx) != M3B_OK) {
    m3b_destroy(ctx);
    return M3B_E_CUDA;
  }
  *out = ctx;
  return M3B_OK;
}

void m3b_destroy(m3b_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  comm_destroy(ctx);
  if (ctx->stream) cudaStreamSynchronize(ctx->stream);
  if (ctx->sched) cudaFree(ctx->sched);
  if (ctx->ingest_flag) cudaFree(ctx->ingest_flag);
  if (ctx->ev0) cudaEventDestroy(ctx->ev0);
  if (ctx->ev1) cudaEventDestroy(ctx->ev1);
  if (ctx->evk0) cudaEventDestroy(ctx->evk0);
  if (ctx->evk1) cudaEventDestroy(ctx->evk1);
  for (cudaEvent_t e : ctx->ev_pool) cudaEventDestroy(e);
  if (ctx->stream2) cudaStreamDestroy(ctx->stream2);
  if (ctx->stream) cudaStreamDestroy(ctx->stream);
  dev_pool_release(ctx->device);
  delete ctx;
}

const char* m3b_last_error(const m3b_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }

int m3b_set_small_svd(m3b_ctx* ctx, m3b_small_svd_fn fn, void* user) {
  if (!ctx) return M3B_E_DIMS;
  ctx->small_svd = fn;
  ctx->small_svd_user = user;
  return M3B_OK;
}

int m3b_set_rnorm(m3b_ctx* ctx, m3b_rnorm_fn fn, void* user) {
  if (!ctx) return M3B_E_DIMS;
  ctx->rnorm = fn;
  ctx->rnorm_user = user;
  return M3B_OK;
}

int m3b_set_should_abort(m3b_ctx* ctx, m3b_should_abort_fn fn, void* user) {
  if (!ctx) return M3B_E_DIMS;
  ctx->should_abort = fn;
  ctx->should_abort_user = user;
  return M3B_OK;
}

int m3b_comm_unique_id(m3b_ctx* ctx, char id[M3B_COMM_ID_BYTES]) {
  if (!ctx || !id) return M3B_E_DIMS;
  return comm_get_unique_id(ctx, id);
}
int m3b_comm_init(m3b_ctx* ctx, int rank, int world, const char id[M3B_COMM_ID_BYTES]) {
  if (!ctx) return M3B_E_DIMS;
  cudaSetDevice(ctx->device);
  return comm_init(ctx, rank, world, id);
}
int m3b_comm_world(const m3b_ctx* ctx) { return ctx ? ctx->comm.world : 1; }
int m3b_comm_p2p_handle(m3b_ctx* ctx, char handle[M3B_P2P_HANDLE_BYTES]) {
  if (!ctx || !handle) return M3B_E_DIMS;
  return comm_p2p_handle(ctx, handle);
}
int m3b_comm_p2p_open(m3b_ctx* ctx, const char* handles) {
  if (!ctx || !handles) return M3B_E_DIMS;
  return comm_p2p_open(ctx, handles);
}

// ---------------------------------------------------------------------------
int m3b_matrix_begin(m3b_ctx* ctx, int32_t n_genes, int64_t n_cells_local, int64_t n_cells_total, int64_t nnz_hint,
                     m3b_mat** out) {
  if (!ctx || !out) return M3B_E_DIMS;
  *out = nullptr;
  if (n_genes < 0 || n_cells_local < 0 || n_cells_total < n_cells_local) {
    m3b_set_error(ctx, "m3b_matrix_begin: bad dimensions");
    return M3B_E_DIMS;
  }
  cudaSetDevice(ctx->device);
  m3b_mat* m = new (std::nothrow) m3b_mat();
  if (!m) return M3B_E_NOMEM;
  m->ctx = ctx;
  m->n_genes = n_genes;
  m->n_cells = n_cells_local;
  m->n_cells_total = n_cells_total;
  m->nnz_cap = std::max<long long>(nnz_hint, 0);
  int st;
  if ((st = dev_alloc(ctx, &m->p, (size_t)n_cells_local + 1)) != M3B_OK ||
      (st = dev_alloc(ctx, &m->i, (size_t)std::max<long long>(m->nnz_cap, 1))) != M3B_OK ||
      (st = dev_alloc(ctx, &m->x, (size_t)std::max<long long>(m->nnz_cap, 1))) != M3B_OK) {
    m3b_matrix_free(m);
    return st;
  }
  cudaMemsetAsync(m->p, 0, sizeof(long long), ctx->stream);
  *out = m;
  return M3B_OK;
}

int m3b_matrix_append_csc(m3b_mat* m, int64_t n_cells_chunk, const int32_t* p, const int32_t* i, const double* x) {
  if (!m || !p) return M3B_E_DIMS;
  cudaSetDevice(m->ctx->device);
  return ingest_append(m, n_cells_chunk, p, i, x);
}

int m3b_matrix_end(m3b_mat* m) {
  if (!m) return M3B_E_DIMS;
  if (m->cells_appended != m->n_cells) {
    m3b_set_error(m->ctx, "m3b_matrix_end: %lld of %lld cells appended", m->cells_appended, m->n_cells);
    return M3B_E_STATE;
  }
  m->ended = true;
  return M3B_OK;
}

void m3b_matrix_free(m3b_mat* m) {
  if (!m) return;
  cudaSetDevice(m->ctx->device);
  cudaStreamSynchronize(m->ctx->stream);
  dev_free(m->p);
  dev_free(m->i);
  dev_free(m->x);
  dev_free(m->y);
  dev_free(m->unit_val);
  dev_free(m->sel_of_gene);
  dev_free(m->kept_of_sel);
  dev_free(m->kept_of_gene);
  dev_free(m->sel_nnz);
  dev_free(m->rowsum);
  dev_free(m->kept_nnz);
  dev_free(m->kept_padlen);
  dev_free(m->mean_sparse);
  dev_free(m->center);
  dev_free(m->scale);
  dev_free(m->Xdev);
  m3b_mat_set_bbuild(m, nullptr);
  for (int q = 0; q < 2; ++q) {
    dev_free(m->b_cnt[q]);
    dev_free(m->b_base[q]);
  }
  m->A.free_all();
  m->B.free_all();
  delete m;
}

int m3b_matrix_dims(const m3b_mat* m, int32_t* n_genes, int64_t* n_cells_local, int64_t* n_cells_total,
                    int64_t* nnz_local) {
  if (!m) return M3B_E_DIMS;
  if (n_genes) *n_genes = m->n_genes;
  if (n_cells_local) *n_cells_local = m->n_cells;
  if (n_cells_total) *n_cells_total = m->n_cells_total;
  if (nnz_local) *nnz_local = m->nnz;
  return M3B_OK;
}

// ---------------------------------------------------------------------------
int m3b_cell_totals(m3b_mat* m, int round_exprs, double* cell_total) {
  if (!m || !cell_total) return M3B_E_DIMS;
  m3b_ctx* ctx = m->ctx;
  if (!m->ended) {
    m3b_set_error(ctx, "m3b_cell_totals before m3b_matrix_end");
    return M3B_E_STATE;
  }
  cudaSetDevice(ctx->device);
  double* d = nullptr;
  RET(dev_alloc(ctx, &d, (size_t)std::max<long long>(m->n_cells, 1)));
  int st = k1_cell_totals(m, round_exprs, d);
  if (st == M3B_OK && m->n_cells) {
    cudaError_t e = cudaMemcpyAsync(cell_total, d, m->n_cells * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) {
      m3b_set_error(ctx, "m3b_cell_totals: %s", cudaGetErrorString(e));
      st = M3B_E_CUDA;
    }
    ctx->stats.d2h_bytes += m->n_cells * 8;
    if (st == M3B_OK && ctx->k1_timed) {
      float ms = 0.f;
      if (cudaEventElapsedTime(&ms, ctx->evk0, ctx->evk1) == cudaSuccess) ctx->stats.k1_ms = ms;
      ctx->k1_timed = false;
    }
  }
  dev_free(d);
  return st;
}

// R's mean(): long-double accumulate + one refinement pass (summary.c real_mean)
static double r_mean(const double* x, int64_t n) {
  long double s = 0.0L;
  for (int64_t k = 0; k < n; ++k) s += x[k];
  s /= (long double)n;
  long double t = 0.0L;
  for (int64_t k = 0; k < n; ++k) t += (x[k] - s);
  s += t / (long double)n;
  return (double)s;
}

int m3b_size_factors_from_totals(const double* tot, int64_t n, int method, double* sf, int* any_zero) {
  if (!tot || !sf || n < 0) return M3B_E_DIMS;
  if (method != M3B_SF_MEAN_GEOMETRIC_MEAN_TOTAL && method != M3B_SF_MEAN_GEOMETRIC_MEAN_LOG_TOTAL) return M3B_E_DIMS;
  std::vector<double> lg((size_t)n);
  // The element-wise parts (one or two libm logs per cell: ~50 ms for 2M cells on one core, on EVERY
  // rank of a sharded run) are spread over a few host threads; the mean itself stays the strictly
  // sequential long-double loop of R's mean(), so the result does not depend on the thread count.
  const int n_thr = (int)std::max<int64_t>(1, std::min<int64_t>(8, n / 65536));
  std::vector<int> zero_seen((size_t)n_thr, 0);
  auto span = [&](int t, int64_t* lo, int64_t* hi) {
    *lo = n * t / n_thr;
    *hi = n * (t + 1) / n_thr;
  };
  auto run = [&](auto&& body) {
    std::vector<std::thread> th;
    for (int t = 1; t < n_thr; ++t) th.emplace_back(body, t);
    body(0);
    for (auto& x : th) x.join();
  };
  run([&](int t) {
    int64_t lo, hi;
    span(t, &lo, &hi);
    for (int64_t k = lo; k < hi; ++k) {
      if (tot[k] == 0.0) zero_seen[t] = 1;
      lg[k] = (method == M3B_SF_MEAN_GEOMETRIC_MEAN_TOTAL) ? log(tot[k]) : log(log(tot[k]));
    }
  });
  int az = 0;
  for (int t = 0; t < n_thr; ++t) az |= zero_seen[t];
  const double denom = n ? exp(r_mean(lg.data(), n)) : NAN;
  run([&](int t) {
    int64_t lo, hi;
    span(t, &lo, &hi);
    for (int64_t k = lo; k < hi; ++k) {
      double num = (method == M3B_SF_MEAN_GEOMETRIC_MEAN_TOTAL) ? tot[k] : log(tot[k]);
      double v = num / denom;
      sf[k] = (v != v) ? 1.0 : v;  // sfs[is.na(sfs)] <- 1   (R/utils.R:68)
    }
  });
  if (any_zero) *any_zero = az;
  return M3B_OK;
}

int m3b_size_factors(m3b_mat* m, int round_exprs, int method, double* cell_total, double* sf, int* any_zero) {
  if (!m || !sf) return M3B_E_DIMS;
  if (m->ctx->comm.world != 1 || m->n_cells != m->n_cells_total) {
    m3b_set_error(m->ctx, "m3b_size_factors is for unsharded matrices; gather m3b_cell_totals and call "
                          "m3b_size_factors_from_totals");
    return M3B_E_STATE;
  }
  std::vector<double> tmp;
  if (!cell_total) {
    tmp.resize((size_t)std::max<long long>(m->n_cells, 1));
    cell_total = tmp.data();
  }
  RET(m3b_cell_totals(m, round_exprs, cell_total));
  return m3b_size_factors_from_totals(cell_total, m->n_cells, method, sf, any_zero);
}

// ---------------------------------------------------------------------------
int m3b_normalize(m3b_mat* m, const double* sf, int norm_method, double pseudo_count) {
  if (!m) return M3B_E_DIMS;
  m3b_ctx* ctx = m->ctx;
  if (!m->ended) {
    m3b_set_error(ctx, "m3b_normalize before m3b_matrix_end");
    return M3B_E_STATE;
  }
  if (norm_method < M3B_NORM_LOG2 || norm_method > M3B_NORM_BINARY) {
    m3b_set_error(ctx, "norm_method must be one of 'log', 'size_only' or 'none'");
    return M3B_E_DIMS;
  }
  const bool need_sf = !(norm_method == M3B_NORM_NONE || norm_method == M3B_NORM_BINARY);
  if (need_sf && !sf) {
    m3b_set_error(ctx, "size factors required");
    return M3B_E_DIMS;
  }
  if (norm_method == M3B_NORM_LOG10 && pseudo_count != 1.0) {
    m3b_set_error(ctx, "Pseudocount must equal 1 with sparse expression matrices");  // R/utils.R:498
    return M3B_E_DIMS;
  }
  cudaSetDevice(ctx->device);
  double* d_sf = nullptr;
  if (need_sf) {
    RET(dev_alloc(ctx, &d_sf, (size_t)std::max<long long>(m->n_cells, 1)));
    if (m->n_cells) {
      cudaError_t e = cudaMemcpyAsync(d_sf, sf, m->n_cells * sizeof(double), cudaMemcpyHostToDevice, ctx->stream);
      if (e != cudaSuccess) {
        dev_free(d_sf);
        m3b_set_error(ctx, "m3b_normalize: %s", cudaGetErrorString(e));
        return M3B_E_CUDA;
      }
      ctx->stats.h2d_bytes += m->n_cells * 8;
    }
  }
  int st = k2_normalize(m, d_sf, norm_method, pseudo_count);
  cudaError_t e = cudaStreamSynchronize(ctx->stream);
  dev_free(d_sf);
  if (st != M3B_OK) return st;
  if (e != cudaSuccess) {
    m3b_set_error(ctx, "m3b_normalize: %s", cudaGetErrorString(e));
    return M3B_E_CUDA;
  }
  if (ctx->k2_timed) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, ctx->evk0, ctx->evk1) == cudaSuccess) ctx->stats.k2_ms = ms;
    ctx->k2_timed = false;
  }
  m->normalized = true;
  m->selected = false;
  m->moments_done = false;
  m->norm_method = norm_method;
  m->pseudo_count = pseudo_count;
  return M3B_OK;
}

int m3b_export_values(m3b_mat* m, double* y) {
  if (!m || !y) return M3B_E_DIMS;
  m3b_ctx* ctx = m->ctx;
  if (!m->normalized) {
    m3b_set_error(ctx, "m3b_export_values before m3b_normalize");
    return M3B_E_STATE;
  }
  cudaSetDevice(ctx->device);
  if (m->nnz) {
    CK(ctx, cudaMemcpyAsync(y, m->y, m->nnz * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CK(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->stats.d2h_bytes += m->nnz * 8;
  }
  return M3B_OK;
}

// ---------------------------------------------------------------------------
int m3b_select_genes(m3b_mat* m, const int32_t* gene_order, int32_t n_sel, uint8_t* keep, int32_t* n_kept) {
  if (!m) return M3B_E_DIMS;
  m3b_ctx* ctx = m->ctx;
  if (!m->normalized) {
    m3b_set_error(ctx, "m3b_select_genes before m3b_normalize");
    return M3B_E_STATE;
  }
  cudaSetDevice(ctx->device);
  if (!gene_order) n_sel = m->n_genes;
  if (n_sel < 0) return M3B_E_DIMS;
  std::vector<int> sel_of_gene(std::max(m->n_genes, 1), -1);
  for (int s = 0; s < n_sel; ++s) {
    int g = gene_order ? gene_order[s] : s;
    if (g < 0 || g >= m->n_genes) {
      m3b_set_error(ctx, "use_genes index %d out of range", g);
      return M3B_E_DIMS;
    }
    if (sel_of_gene[g] != -1) {
      m3b_set_error(ctx, "use_genes lists gene %d twice", g);
      return M3B_E_DIMS;
    }
    sel_of_gene[g] = s;
  }
  m->sel_monotone = true;
  if (gene_order)
    for (int s = 1; s < n_sel; ++s)
      if (gene_order[s] <= gene_order[s - 1]) m->sel_monotone = false;
  const bool dbg = getenv("M3B_DEBUG_TIMING") != nullptr;
  auto now = []() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
  const double t0 = now();
  m->n_sel = n_sel;
  dev_free(m->sel_of_gene);
  RET(dev_alloc(ctx, &m->sel_of_gene, sel_of_gene.size()));
  CK(ctx, cudaMemcpyAsync(m->sel_of_gene, sel_of_gene.data(), sel_of_gene.size() * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
  // count-1 entries go to the index-only unit layouts whenever their value is a function of the
  // cell alone (every size-factor based method); TF-IDF rewrites the values per (gene, cell), so
  // the LSI branch re-builds unsplit (tfidf.cu)
  const bool split = !(ctx->flags & M3B_FLAG_NO_UNIT_SPLIT) && !getenv("M3B_NO_UNIT_SPLIT") &&
                     (m->norm_method == M3B_NORM_LOG2 || m->norm_method == M3B_NORM_LOG10 ||
                      m->norm_method == M3B_NORM_SIZE_ONLY);
  RET(build_layout_B(m, split));
  m->split = split;
  const double t1 = now();
  // global row sums of the (implicit dense) normalised matrix -> zero / non-finite filter
  dev_free(m->rowsum);
  RET(dev_alloc(ctx, &m->rowsum, (size_t)std::max(n_sel, 1)));
  std::vector<double> rs(std::max(n_sel, 1), 0.0);
  if (n_sel) {
    CK(ctx, cudaMemsetAsync(m->rowsum, 0, n_sel * sizeof(double), ctx->stream));
    RET(launch_tile_stream(ctx, m->B, OP_SUM, nullptr));
    RET(launch_combine_rows(ctx, m->B, m->rowsum));
    RET(comm_allreduce_f64(ctx, m->rowsum, n_sel));
    CK(ctx, cudaMemcpyAsync(rs.data(), m->rowsum, n_sel * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  }
  CK(ctx, cudaStreamSynchronize(ctx->stream));
  std::vector<unsigned char> kp(std::max(n_sel, 1), 0);
  for (int s = 0; s < n_sel; ++s) {
    double v = rs[s] + (double)m->n_cells_total * m->f0;
    kp[s] = (std::isfinite(v) && v != 0.0) ? 1 : 0;  // R/preprocess_cds.R:121-122
  }
  const double t2 = now();
  m->keep_mask = kp;
  RET(finish_selection(m, kp.data()));
  const double t3 = now();
  RET(build_layout_A(m, split));
  const double t4 = now();
  ctx->stats.select_ms[0] = t1 - t0;
  ctx->stats.select_ms[1] = t2 - t1;
  ctx->stats.select_ms[2] = t3 - t2;
  ctx->stats.select_ms[3] = t4 - t3;
  if (dbg)
    fprintf(stderr, "m3b_select_genes: layout B %.2f ms, row sums %.2f ms, remap %.2f ms, layout A %.2f ms\n", t1 - t0,
            t2 - t1, t3 - t2, t4 - t3);
  if (keep) memcpy(keep, kp.data(), n_sel);
  if (n_kept) *n_kept = m->n_kept;
  m->selected = true;
  m->layouts_built = true;
  m->moments_done = false;
  m->tfidf_applied = false;
  ctx->stats.spmv_bytes_cells_out = spmv_algorithmic_bytes(m->A);
  ctx->stats.spmv_bytes_genes_out = spmv_algorithmic_bytes(m->B);
  ctx->stats.spmv_stream_bytes_cells_out = spmv_stream_bytes(m->A);
  ctx->stats.spmv_stream_bytes_genes_out = spmv_stream_bytes(m->B);
  return M3B_OK;
}

int m3b_select_genes_for_transform(m3b_mat* m, const int32_t* gene_order, int32_t n_sel, uint8_t* keep, int32_t* n_kept) {
  if (!m) return M3B_E_DIMS;
  m3b_ctx* ctx = m->ctx;
  if (!m->normalized) {
    m3b_set_error(ctx, "m3b_select_genes_for_transform before m3b_normalize");
    return M3B_E_STATE;
  }
  cudaSetDevice(ctx->device);
  int fallback = 0, nk = 0;
  RET(run_select_for_transform(m, gene_order, n_sel, keep, &nk, &fallback));
  if (fallback) return m3b_select_genes(m, gene_order, n_sel, keep, n_kept);
  if (n_kept) *n_kept = nk;
  return M3B_OK;
}

int m3b_detect_genes(m3b_mat* m, double min_expr, int64_t* num_cells_expressed, int64_t* num_genes_expressed) {
  if (!m || !num_cells_expressed || !num_genes_expressed) return M3B_E_DIMS;
  if (!m->ended) {
    m3b_set_error(m->ctx, "m3b_detect_genes before m3b_matrix_end");
    return M3B_E_STATE;
  }
  cudaSetDevice(m->ctx->device);
  static_assert(sizeof(long long) == sizeof(int64_t), "int64_t");
  return run_detect_genes(m, min_expr, reinterpret_cast<long long*>(num_cells_expressed),
                          reinterpret_cast<long long*>(num_genes_expressed));
}

int m3b_aggregate_expression(m3b_mat* m, const int32_t* gene_group, int32_t n_gene_groups, int gene_agg_mean,
                             const int32_t* cell_group, int32_t n_cell_groups, int cell_agg_mean, double* out,
                             int64_t ld_out) {
  if (!m || !gene_group || !out) return M3B_E_DIMS;
  cudaSetDevice(m->ctx->device);
  return run_aggregate(m, gene_group, n_gene_groups, gene_agg_mean, cell_group, n_cell_groups, cell_agg_mean, out, ld_out);
}

int m3b_jaccard_coeff(m3b_ctx* ctx, const double* idx, int64_t n, int32_t k, int weight, double* out) {
  if (!ctx || !idx || !out) return M3B_E_DIMS;
  cudaSetDevice(ctx->device);
  return run_jaccard(ctx, idx, (long long)n, (int)k, weight, out);
}

int m3b_knn_self(m3b_ctx* ctx, const double* X, int64_t n, int32_t d, int32_t k, int32_t* nn_idx, double* nn_dists) {
  if (!ctx || !X || !nn_idx || !nn_dists) return M3B_E_DIMS;
  cudaSetDevice(ctx->device);
  return run_knn_self(ctx, X, (long long)n, (int)d, (int)k, nn_idx, nn_dists);
}

int m3b_align_residuals(m3b_ctx* ctx, double* X, int64_t n, int nv, int64_t ldx, const double* M, int p, double* beta) {
  if (!ctx || !X || !M || !beta) return M3B_E_DIMS;
  cudaSetDevice(ctx->device);
  return run_align_residuals(ctx, X, (long long)n, nv, (long long)ldx, M, p, beta, true);
}

int m3b_align_apply_beta(m3b_ctx* ctx, double* X, int64_t n, int nv, int64_t ldx, const double* M, int p, const double* beta) {
  if (!ctx || !X || !M || !beta) return M3B_E_DIMS;
  cudaSetDevice(ctx->device);
  return run_align_residuals(ctx, X, (long long)n, nv, (long long)ldx, M, p, const_cast<double*>(beta), false);
}

int m3b_gene_nnz(m3b_mat* m, int64_t* nnz_per_sel_gene) {
  if (!m || !nnz_per_sel_gene) return M3B_E_DIMS;
  m3b_ctx* ctx = m->ctx;
  if (!m->selected || !m->layouts_built) {
    m3b_set_error(ctx, "m3b_gene_nnz before m3b_select_genes");
    return M3B_E_STATE;
  }
  cudaSetDevice(ctx->device);
  if (m->n_sel) {
    CK(ctx, cudaMemcpyAsync(nnz_per_sel_gene, m->sel_nnz, m->n_sel * sizeof(long long), cudaMemcpyDeviceToHost, ctx->stream));
    CK(ctx, cudaStreamSynchronize(ctx->stream));
  }
  return M3B_OK;
}

int m3b_gene_moments(m3b_mat* m, double* center, double* scale) {
  if (!m) return M3B_E_DIMS;
  m3b_ctx* ctx = m->ctx;
  if (!m->selected || !m->layouts_built) {
    m3b_set_error(ctx, "m3b_gene_moments before m3b_select_genes");
    return M3B_E_STATE;
  }
  cudaSetDevice(ctx->device);
  RET(compute_moments(m));
  if (m->n_kept) {
    if (center) CK(ctx, cudaMemcpyAsync(center, m->center, m->n_kept * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    if (scale) CK(ctx, cudaMemcpyAsync(scale, m->scale, m->n_kept * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CK(ctx, cudaStreamSynchronize(ctx->stream));
  }
  return M3B_OK;
}

int m3b_tfidf(m3b_mat* m, int frequencies, int log_scale_tf, double scale_factor, const double* row_sums_in,
              double num_cols, double* col_sums_out, double* row_sums_out) {
  if (!m) return M3B_E_DIMS;
  cudaSetDevice(m->ctx->device);
  return run_tfidf(m, frequencies, log_scale_tf, scale_factor, row_sums_in, num_cols, col_sums_out, row_sums_out);
}

int m3b_pca_irlba(m3b_mat* m, int nv, int work, int maxit, double tol, double svtol, const double* v0, int center,
                  int scale, double* d, double* V, double* X, int64_t ldx, double* center_out, double* scale_out,
                  int* iter, int* mprod) {
  if (!m || !v0 || !d) return M3B_E_DIMS;
  if (!m->selected || !m->layouts_built) {
    m3b_set_error(m->ctx, "m3b_pca_irlba before m3b_select_genes");
    return M3B_E_STATE;
  }
  cudaSetDevice(m->ctx->device);
  return run_irlba(m, nv, work, maxit, tol, svtol, v0, center, scale, d, V, X, ldx, center_out, scale_out, iter, mprod,
                   true);
}

int m3b_pca_irlba_device(m3b_mat* m, int nv, int work, int maxit, double tol, double svtol, const double* v0,
                         int center, int scale, double* d, int* iter, int* mprod) {
  if (!m || !v0 || !d) return M3B_E_DIMS;
  if (!m->selected || !m->layouts_built) {
    m3b_set_error(m->ctx, "m3b_pca_irlba_device before m3b_select_genes");
    return M3B_E_STATE;
  }
  cudaSetDevice(m->ctx->device);
  return run_irlba(m, nv, work, maxit, tol, svtol, v0, center, scale, d, nullptr, nullptr, 0, nullptr, nullptr, iter,
                   mprod, false);
}

int m3b_project(m3b_mat* q, const int32_t* model_row, int32_t n_model, const double* V, const double* center,
                const double* scale, int nv, double* Y, int64_t ldy) {
  if (!q || !model_row || !V || !Y) return M3B_E_DIMS;
  cudaSetDevice(q->ctx->device);
  return run_project(q, model_row, n_model, V, center, scale, nv, Y, ldy);
}

// fp64 FMA peak of this device, measured: the roof the projection kernel (K10) is reported against
// (MEASURED_PEAKS.json has no fp64 figure).  8 independent DFMA chains per thread, 2 x 1024-thread
// CTAs per SM, CUDA events; best of 5.
__global__ void __launch_bounds__(1024) k_fp64_peak(double* __restrict__ out, int iters, double x, double y) {
  double a[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) a[k] = 1e-3 * (threadIdx.x + k);
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int k = 0; k < 8; ++k) a[k] = fma(a[k], x, y);
  }
  double s = 0.0;
#pragma unroll
  for (int k = 0; k < 8; ++k) s += a[k];
  out[blockIdx.x * (size_t)blockDim.x + threadIdx.x] = s;
}

int m3b_measure_fp64_peak(m3b_ctx* ctx, double* tflops) {
  if (!ctx || !tflops) return M3B_E_DIMS;
  cudaSetDevice(ctx->device);
  const int grid = ctx->sm_count * 2, threads = 1024, iters = 4096;
  double* d = nullptr;
  RET(dev_alloc(ctx, &d, (size_t)grid * threads));
  double best = 0.0;
  for (int rep = 0; rep < 6; ++rep) {
    cudaEventRecord(ctx->evk0, ctx->stream);
    k_fp64_peak<<<grid, threads, 0, ctx->stream>>>(d, iters, 0.999999, 1e-7);
    cudaEventRecord(ctx->evk1, ctx->stream);
    if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) break;
    float ms = 0.f;
    cudaEventElapsedTime(&ms, ctx->evk0, ctx->evk1);
    const double fl = 2.0 * 32.0 * (double)iters * (double)grid * (double)threads;
    if (rep > 0 && ms > 0.f) best = std::max(best, fl / (ms * 1e-3) / 1e12);
  }
  count_launch(ctx, 6);
  dev_free(d);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess || best == 0.0) {
    m3b_set_error(ctx, "m3b_measure_fp64_peak: %s", cudaGetErrorString(e));
    return M3B_E_CUDA;
  }
  *tflops = best;
  return M3B_OK;
}

int m3b_get_stats(const m3b_ctx* ctx, m3b_stats* out) {
  if (!ctx || !out) return M3B_E_DIMS;
  *out = ctx->stats;
  return M3B_OK;
}

int m3b_set_flags(m3b_ctx* ctx, int flags) {
  if (!ctx) return M3B_E_DIMS;
  ctx->flags = flags;
  return M3B_OK;
}

int m3b_timer_start(m3b_ctx* ctx) {
  if (!ctx) return M3B_E_DIMS;
  cudaSetDevice(ctx->device);
  CK(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
  return M3B_OK;
}

int m3b_timer_stop(m3b_ctx* ctx, double* elapsed_ms) {
  if (!ctx || !elapsed_ms) return M3B_E_DIMS;
  cudaSetDevice(ctx->device);
  CK(ctx, cudaEventRecord(ctx->ev1, ctx->stream));
  CK(ctx, cudaEventSynchronize(ctx->ev1));
  float ms = 0.f;
  CK(ctx, cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
  *elapsed_ms = ms;
  return M3B_OK;
}

int m3b_reset_stats(m3b_ctx* ctx) {
  if (!ctx) return M3B_E_DIMS;
  long long a = ctx->stats.spmv_bytes_cells_out, b = ctx->stats.spmv_bytes_genes_out;
  long long c = ctx->stats.spmv_stream_bytes_cells_out, d = ctx->stats.spmv_stream_bytes_genes_out;
  ctx->stats = m3b_stats{};
  ctx->stats.spmv_bytes_cells_out = a;
  ctx->stats.spmv_bytes_genes_out = b;
  ctx->stats.spmv_stream_bytes_cells_out = c;
  ctx->stats.spmv_stream_bytes_genes_out = d;
  return M3B_OK;
}

}  // extern "C"
