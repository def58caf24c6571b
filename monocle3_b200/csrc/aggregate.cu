// detect_genes and the group sums behind aggregate_gene_expression,
// both single passes over the cell-major CSC (the K1 shape: one warp per cell).
//
//   detect_genes                R/utils.R:449-457
//       rowData$num_cells_expressed = Matrix::rowSums(counts > min_expr)
//       colData$num_genes_expressed = Matrix::colSums(counts > min_expr)
//   aggregate_gene_expression   R/cluster_genes.R:447-536
//       agg_mat = normalized_counts(cds, norm_method, pseudocount); per gene group colSums /
//       colMeans of the member rows; per cell group my.aggregate.Matrix(fun = mean | sum);
//       the row scaling / clamping that follows works on the small dense result and stays on
//       the host (api.py).
#include <algorithm>
#include <vector>

#include "m3b_internal.h"

__device__ __forceinline__ double ag_ld_f64(const double* p) {
  double v;
  asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p));
  return v;
}

// per cell: stored entries above the threshold; per gene: the same (integer atomics: exact and
// order-independent) plus the stored-entry count, needed when the threshold is negative and the
// implicit zeros qualify too
// use_smem: the per-gene counters of a CTA live in shared memory (2 x n_genes x 4 B) and are
// flushed once; otherwise global atomics per entry (very many genes)
__global__ void __launch_bounds__(512) k_detect(const long long* __restrict__ p, const int* __restrict__ gi,
                                                const double* __restrict__ x, long long n_cells, int n_genes,
                                                double min_expr, int need_nnz, int use_smem,
                                                unsigned long long* __restrict__ gene_gt,
                                                unsigned long long* __restrict__ gene_nnz,
                                                long long* __restrict__ cell_gt) {
  extern __shared__ unsigned int dg_hist[];  // [2][n_genes] when use_smem
  unsigned int* h_gt = dg_hist;
  unsigned int* h_nz = dg_hist + n_genes;
  if (use_smem) {
    for (int g = threadIdx.x; g < 2 * n_genes; g += blockDim.x) dg_hist[g] = 0u;
    __syncthreads();
  }
  const int lane = threadIdx.x & 31;
  long long warp = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long c = warp; c < n_cells; c += nwarps) {
    const long long lo = p[c], hi = p[c + 1];
    int n = 0;
    for (long long k = lo + lane; k < hi; k += 32) {
      const int g = gi[k];
      const bool on = ag_ld_f64(x + k) > min_expr;
      if (use_smem) {
        if (need_nnz) atomicAdd(h_nz + g, 1u);
        if (on) atomicAdd(h_gt + g, 1u);
      } else {
        if (need_nnz) atomicAdd(gene_nnz + g, 1ULL);
        if (on) atomicAdd(gene_gt + g, 1ULL);
      }
      n += on ? 1 : 0;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) n += __shfl_xor_sync(0xffffffffu, n, o);
    if (lane == 0) cell_gt[c] = n;
  }
  if (use_smem) {
    __syncthreads();
    for (int g = threadIdx.x; g < n_genes; g += blockDim.x) {
      if (h_gt[g]) atomicAdd(gene_gt + g, (unsigned long long)h_gt[g]);
      if (need_nnz && h_nz[g]) atomicAdd(gene_nnz + g, (unsigned long long)h_nz[g]);
    }
  }
}

int run_detect_genes(m3b_mat* m, double min_expr, long long* num_cells_expressed, long long* num_genes_expressed) {
  m3b_ctx* ctx = m->ctx;
  const int G = m->n_genes;
  const long long C = m->n_cells;
  unsigned long long *d_gt = nullptr, *d_nnz = nullptr;
  long long* d_cell = nullptr;
  DevScope guard;
  guard.own(&d_gt);
  guard.own(&d_cell);
  RET(dev_alloc(ctx, &d_gt, (size_t)std::max(2 * G, 1)));
  d_nnz = d_gt + G;
  RET(dev_alloc(ctx, &d_cell, (size_t)std::max<long long>(C, 1)));
  int st = M3B_OK;
  cudaError_t e = cudaSuccess;
  do {
    if ((e = cudaMemsetAsync(d_gt, 0, (size_t)std::max(2 * G, 1) * sizeof(unsigned long long), ctx->stream)) != cudaSuccess) break;
    if (C) {
      // (the stored-entry counts are only needed when the implicit zeros qualify)
      const size_t hist = 2 * (size_t)G * sizeof(unsigned int);
      const int use_smem = hist <= 200 * 1024;
      if (use_smem && hist > 48 * 1024 &&
          (e = cudaFuncSetAttribute(k_detect, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)hist)) != cudaSuccess)
        break;
      const int blocks = use_smem ? (int)std::min<long long>((C + 15) / 16, (long long)ctx->sm_count)
                                  : (int)std::min<long long>((C + 15) / 16, (long long)ctx->sm_count * 4);
      k_detect<<<blocks, 512, use_smem ? hist : 0, ctx->stream>>>(m->p, m->i, m->x, C, G, min_expr, min_expr < 0.0 ? 1 : 0,
                                                                   use_smem, d_gt, d_nnz, d_cell);
      count_launch(ctx);
      if ((e = cudaGetLastError()) != cudaSuccess) break;
    }
    if ((st = comm_allreduce_i64(ctx, reinterpret_cast<long long*>(d_gt), (size_t)2 * G)) != M3B_OK) break;
    std::vector<long long> h((size_t)std::max(2 * G, 1));
    if (G && (e = cudaMemcpyAsync(h.data(), d_gt, (size_t)2 * G * sizeof(long long), cudaMemcpyDeviceToHost, ctx->stream)) != cudaSuccess) break;
    if (C && (e = cudaMemcpyAsync(num_genes_expressed, d_cell, C * sizeof(long long), cudaMemcpyDeviceToHost, ctx->stream)) != cudaSuccess) break;
    if ((e = cudaStreamSynchronize(ctx->stream)) != cudaSuccess) break;
    ctx->stats.d2h_bytes += (2LL * G + C) * 8;
    // a negative threshold makes the implicit zeros count as expressed (0 > min_expr)
    for (int g = 0; g < G; ++g)
      num_cells_expressed[g] = h[g] + (min_expr < 0.0 ? (m->n_cells_total - h[(size_t)G + g]) : 0);
    if (min_expr < 0.0) {
      std::vector<long long> hp((size_t)C + 1);
      if ((e = cudaMemcpy(hp.data(), m->p, (C + 1) * sizeof(long long), cudaMemcpyDeviceToHost)) != cudaSuccess) break;
      for (long long c = 0; c < C; ++c) num_genes_expressed[c] += G - (hp[c + 1] - hp[c]);
    }
  } while (0);
  if (e != cudaSuccess) {
    m3b_set_error(ctx, "m3b_detect_genes: %s", cudaGetErrorString(e));
    return M3B_E_CUDA;
  }
  return st;
}

// acc[group][cell] = sum over the cell's entries whose gene is in the group, in entry order.
// One warp per cell.  Entries are loaded 32 at a time and broadcast; lane l owns the groups
// g == l (mod 32), so every group is summed by one lane in a fixed order (bit-reproducible, no
// floating-point atomics).
#define AG_WARPS 8
__global__ void __launch_bounds__(32 * AG_WARPS) k_agg_cells(const long long* __restrict__ p, const int* __restrict__ gi,
                                                             const double* __restrict__ y, long long n_cells,
                                                             const int* __restrict__ gene_group, int n_groups,
                                                             double f0, const int* __restrict__ group_size,
                                                             int gene_mean, double* __restrict__ out, long long ld) {
  extern __shared__ double ag_acc[];  // [AG_WARPS][n_groups]
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  double* acc = ag_acc + (size_t)w * n_groups;
  long long warp = blockIdx.x * (long long)AG_WARPS + w;
  const long long nwarps = (long long)gridDim.x * AG_WARPS;
  for (long long c = warp; c < n_cells; c += nwarps) {
    for (int g = lane; g < n_groups; g += 32) acc[g] = 0.0;
    __syncwarp();
    const long long lo = p[c], hi = p[c + 1];
    for (long long k0 = lo; k0 < hi; k0 += 32) {
      const long long k = k0 + lane;
      int grp = -1;
      double v = 0.0;
      if (k < hi) {
        grp = gene_group[gi[k]];
        if (grp >= 0) v = ag_ld_f64(y + k);
      }
      const unsigned any = __ballot_sync(0xffffffffu, grp >= 0);
      for (unsigned mask = any; mask; mask &= mask - 1) {
        const int src = __ffs(mask) - 1;
        const int g2 = __shfl_sync(0xffffffffu, grp, src);
        const double v2 = __shfl_sync(0xffffffffu, v, src);
        if ((g2 & 31) == lane) acc[g2] += v2;
      }
    }
    __syncwarp();
    // dense constant of the pseudo-count paths (size_only with pseudocount != 0): every member
    // gene of the group contributes f0, stored or not
    for (int g = lane; g < n_groups; g += 32) {
      double v = acc[g] + f0 * (double)group_size[g];
      if (gene_mean) v = v / (double)group_size[g];
      out[(size_t)c * ld + g] = v;
    }
    __syncwarp();
  }
}

// out[g, cg] = sum over the cells of cell group cg.  Block = 32 gene groups x 16 slices: slice s
// adds the cells q == s (mod 16) of the group's list in list order, the 16 slice sums are then
// added in slice order -- a fixed summation order.  grid = (ceil(n_groups/32), n_cell_groups).
__global__ void __launch_bounds__(512) k_agg_groups(const double* __restrict__ cells, long long ld, int n_groups,
                                                    const long long* __restrict__ cg_ptr,
                                                    const long long* __restrict__ cg_cells, double* __restrict__ out) {
  __shared__ double part[16][33];
  const int gl = threadIdx.x & 31, sl = threadIdx.x >> 5;
  const int g = blockIdx.x * 32 + gl, cg = blockIdx.y;
  double a = 0.0;
  if (g < n_groups)
    for (long long q = cg_ptr[cg] + sl; q < cg_ptr[cg + 1]; q += 16) a += cells[(size_t)cg_cells[q] * ld + g];
  part[sl][gl] = a;
  __syncthreads();
  if (sl == 0 && g < n_groups) {
    double t = 0.0;
#pragma unroll
    for (int s2 = 0; s2 < 16; ++s2) t += part[s2][gl];
    out[(size_t)cg * n_groups + g] = t;
  }
}

int run_aggregate(m3b_mat* m, const int* gene_group, int n_groups, int gene_mean, const int* cell_group, int n_cell_groups,
                  int cell_mean, double* out, long long ld_out) {
  m3b_ctx* ctx = m->ctx;
  const int G = m->n_genes;
  const long long C = m->n_cells;
  if (!m->normalized) {
    m3b_set_error(ctx, "m3b_aggregate_expression before m3b_normalize");
    return M3B_E_STATE;
  }
  if (n_groups < 1 || n_groups > 1024 || (cell_group && n_cell_groups < 1) || ld_out < n_groups) {
    m3b_set_error(ctx, "m3b_aggregate_expression: bad group counts / leading dimension");
    return M3B_E_DIMS;
  }
  std::vector<int> gsize(n_groups, 0);
  for (int g = 0; g < G; ++g) {
    if (gene_group[g] >= n_groups) {
      m3b_set_error(ctx, "gene group id %d out of range", gene_group[g]);
      return M3B_E_DIMS;
    }
    if (gene_group[g] >= 0) gsize[gene_group[g]]++;
  }
  // cell lists per cell group (host): cells in index order
  std::vector<long long> cg_ptr, cg_cells;
  std::vector<long long> cg_count;
  if (cell_group) {
    cg_count.assign(n_cell_groups, 0);
    for (long long c = 0; c < C; ++c) {
      if (cell_group[c] >= n_cell_groups) {
        m3b_set_error(ctx, "cell group id %d out of range", cell_group[c]);
        return M3B_E_DIMS;
      }
      if (cell_group[c] >= 0) cg_count[cell_group[c]]++;
    }
    cg_ptr.assign((size_t)n_cell_groups + 1, 0);
    for (int q = 0; q < n_cell_groups; ++q) cg_ptr[q + 1] = cg_ptr[q] + cg_count[q];
    cg_cells.resize((size_t)cg_ptr[n_cell_groups]);
    std::vector<long long> fill(cg_ptr.begin(), cg_ptr.end() - 1);
    for (long long c = 0; c < C; ++c)
      if (cell_group[c] >= 0) cg_cells[(size_t)fill[cell_group[c]]++] = c;
  }
  int *d_gg = nullptr, *d_gs = nullptr;
  double *d_cells = nullptr, *d_out = nullptr;
  long long *d_ptr = nullptr, *d_list = nullptr;
  DevScope guard;
  guard.own(&d_gg);
  guard.own(&d_gs);
  guard.own(&d_cells);
  guard.own(&d_out);
  guard.own(&d_ptr);
  guard.own(&d_list);
  RET(dev_alloc(ctx, &d_gg, (size_t)std::max(G, 1)));
  RET(dev_alloc(ctx, &d_gs, (size_t)n_groups));
  RET(dev_alloc(ctx, &d_cells, (size_t)std::max<long long>(C, 1) * n_groups));
  int st = M3B_OK;
  cudaError_t e = cudaSuccess;
  do {
    if (G && (e = cudaMemcpyAsync(d_gg, gene_group, (size_t)G * sizeof(int), cudaMemcpyHostToDevice, ctx->stream)) != cudaSuccess) break;
    if ((e = cudaMemcpyAsync(d_gs, gsize.data(), n_groups * sizeof(int), cudaMemcpyHostToDevice, ctx->stream)) != cudaSuccess) break;
    if (C) {
      const size_t smem = (size_t)AG_WARPS * n_groups * sizeof(double);
      if (smem > 48 * 1024) {
        if ((e = cudaFuncSetAttribute(k_agg_cells, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) break;
      }
      const int blocks = (int)std::min<long long>((C + AG_WARPS - 1) / AG_WARPS, (long long)ctx->sm_count * 8);
      k_agg_cells<<<blocks, 32 * AG_WARPS, smem, ctx->stream>>>(m->p, m->i, m->y, C, d_gg, n_groups, m->f0, d_gs, gene_mean,
                                                              d_cells, n_groups);
      count_launch(ctx);
      if ((e = cudaGetLastError()) != cudaSuccess) break;
    }
    if (!cell_group) {
      if (C && (e = cudaMemcpy2DAsync(out, ld_out * sizeof(double), d_cells, n_groups * sizeof(double), n_groups * sizeof(double), C,
                                      cudaMemcpyDeviceToHost, ctx->stream)) != cudaSuccess)
        break;
      ctx->stats.d2h_bytes += C * n_groups * 8;
    } else {
      if ((st = dev_alloc(ctx, &d_out, (size_t)n_groups * n_cell_groups + n_cell_groups)) != M3B_OK) break;
      if ((st = dev_alloc(ctx, &d_ptr, cg_ptr.size())) != M3B_OK) break;
      if ((st = dev_alloc(ctx, &d_list, std::max<size_t>(cg_cells.size(), 1))) != M3B_OK) break;
      if ((e = cudaMemcpyAsync(d_ptr, cg_ptr.data(), cg_ptr.size() * sizeof(long long), cudaMemcpyHostToDevice, ctx->stream)) != cudaSuccess) break;
      if (!cg_cells.empty() &&
          (e = cudaMemcpyAsync(d_list, cg_cells.data(), cg_cells.size() * sizeof(long long), cudaMemcpyHostToDevice, ctx->stream)) != cudaSuccess)
        break;
      const long long ne = (long long)n_groups * n_cell_groups;
      k_agg_groups<<<dim3((n_groups + 31) / 32, n_cell_groups), 512, 0, ctx->stream>>>(d_cells, n_groups, n_groups, d_ptr, d_list, d_out);
      count_launch(ctx);
      if ((e = cudaGetLastError()) != cudaSuccess) break;
      // group sizes ride behind the sums so that one allreduce covers both
      std::vector<double> cnt(n_cell_groups);
      for (int q = 0; q < n_cell_groups; ++q) cnt[q] = (double)cg_count[q];
      if ((e = cudaMemcpyAsync(d_out + ne, cnt.data(), n_cell_groups * sizeof(double), cudaMemcpyHostToDevice, ctx->stream)) != cudaSuccess) break;
      if ((st = comm_allreduce_f64(ctx, d_out, (size_t)ne + n_cell_groups)) != M3B_OK) break;
      std::vector<double> h((size_t)ne + n_cell_groups);
      if ((e = cudaMemcpyAsync(h.data(), d_out, h.size() * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream)) != cudaSuccess) break;
      if ((e = cudaStreamSynchronize(ctx->stream)) != cudaSuccess) break;
      ctx->stats.d2h_bytes += (long long)h.size() * 8;
      for (int q = 0; q < n_cell_groups; ++q)
        for (int g = 0; g < n_groups; ++g) {
          double v = h[(size_t)q * n_groups + g];
          if (cell_mean) v = v / h[(size_t)ne + q];
          out[(size_t)q * ld_out + g] = v;
        }
    }
    e = cudaStreamSynchronize(ctx->stream);
  } while (0);
  if (e != cudaSuccess) {
    m3b_set_error(ctx, "m3b_aggregate_expression: %s", cudaGetErrorString(e));
    return M3B_E_CUDA;
  }
  return st;
}
