// K11: NCCL plumbing for the cell-sharded run (one context per rank).
//
// The reference has no distributed backend at all; the
// collectives exist only because cells are sharded across GPUs: one sum-allreduce
// of the gene-length partial A^T w per Lanczos step, one of the j+1 Gram-Schmidt
// coefficients and one of {||w||^2, sum w}; plus the moment / row-sum reductions
// during preparation.  NCCL is resolved with dlopen at first use so that a host
// process which already carries NCCL (torch) shares that copy.
#include <dlfcn.h>
#include <nccl.h>
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "m3b_internal.h"

struct NcclApi {
  void* handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

static NcclApi* nccl_api(m3b_ctx* ctx) {
  static NcclApi api;
  static bool tried = false;
  if (api.handle) return &api;
  if (tried) return nullptr;
  tried = true;
  const char* cands[4] = {getenv("M3B_NCCL_LIB"), "libnccl.so.2", "libnccl.so", nullptr};
  // prefer a copy that is already mapped into the process
  void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
  for (int k = 0; !h && k < 3; ++k)
    if (cands[k]) h = dlopen(cands[k], RTLD_NOW | RTLD_LOCAL);
  if (!h) {
    m3b_set_error(ctx, "cannot load NCCL: %s", dlerror());
    return nullptr;
  }
  api.GetUniqueId = (decltype(api.GetUniqueId))dlsym(h, "ncclGetUniqueId");
  api.CommInitRank = (decltype(api.CommInitRank))dlsym(h, "ncclCommInitRank");
  api.CommDestroy = (decltype(api.CommDestroy))dlsym(h, "ncclCommDestroy");
  api.AllReduce = (decltype(api.AllReduce))dlsym(h, "ncclAllReduce");
  api.GetErrorString = (decltype(api.GetErrorString))dlsym(h, "ncclGetErrorString");
  if (!api.GetUniqueId || !api.CommInitRank || !api.CommDestroy || !api.AllReduce || !api.GetErrorString) {
    m3b_set_error(ctx, "NCCL library lacks required symbols");
    return nullptr;
  }
  api.handle = h;
  return &api;
}

#define NK(ctx, api, call)                                                              \
  do {                                                                                  \
    ncclResult_t r__ = (call);                                                          \
    if (r__ != ncclSuccess) {                                                           \
      m3b_set_error((ctx), "%s:%d: %s -> %s", __FILE__, __LINE__, #call,                \
                    (api)->GetErrorString(r__));                                        \
      return M3B_E_NCCL;                                                                \
    }                                                                                   \
  } while (0)

int comm_get_unique_id(m3b_ctx* ctx, char* id) {
  NcclApi* api = nccl_api(ctx);
  if (!api) return M3B_E_NCCL;
  static_assert(sizeof(ncclUniqueId) == M3B_COMM_ID_BYTES, "ncclUniqueId size");
  ncclUniqueId uid;
  NK(ctx, api, api->GetUniqueId(&uid));
  memcpy(id, &uid, sizeof(uid));
  return M3B_OK;
}

int comm_init(m3b_ctx* ctx, int rank, int world, const char* id) {
  if (world < 1 || rank < 0 || rank >= world) {
    m3b_set_error(ctx, "m3b_comm_init: bad rank/world %d/%d", rank, world);
    return M3B_E_DIMS;
  }
  comm_destroy(ctx);
  ctx->comm.rank = rank;
  ctx->comm.world = world;
  if (world == 1) return M3B_OK;
  NcclApi* api = nccl_api(ctx);
  if (!api) return M3B_E_NCCL;
  ncclUniqueId uid;
  memcpy(&uid, id, sizeof(uid));
  ncclComm_t c;
  CK(ctx, cudaSetDevice(ctx->device));
  NK(ctx, api, api->CommInitRank(&c, world, uid, rank));
  ctx->comm.comm = (void*)c;
  return M3B_OK;
}

static size_t xchg_bytes(int world) {
  return ((size_t)world * 2 * M3B_XCHG_FLAG_STRIDE + 2 * (size_t)M3B_XCHG_CAP) * sizeof(double) +
         2 * M3B_LL_ENTRIES(world) * 16;
}

// Peer-memory exchange buffer: allocate + export (step 1), import everybody's (step 2).
int comm_p2p_handle(m3b_ctx* ctx, char* out64) {
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t size");
  if (ctx->comm.world < 2 || ctx->comm.world > 16) {
    m3b_set_error(ctx, "peer exchange needs 2..16 ranks");
    return M3B_E_STATE;
  }
  CK(ctx, cudaSetDevice(ctx->device));
  if (!ctx->comm.xchg) {
    CK(ctx, cudaMalloc((void**)&ctx->comm.xchg, xchg_bytes(ctx->comm.world)));
    CK(ctx, cudaMemset(ctx->comm.xchg, 0, xchg_bytes(ctx->comm.world)));
  }
  cudaIpcMemHandle_t h;
  CK(ctx, cudaIpcGetMemHandle(&h, ctx->comm.xchg));
  memcpy(out64, &h, 64);
  return M3B_OK;
}

int comm_p2p_open(m3b_ctx* ctx, const char* handles) {
  Comm& c = ctx->comm;
  if (!c.xchg) {
    m3b_set_error(ctx, "m3b_comm_p2p_open before m3b_comm_p2p_handle");
    return M3B_E_STATE;
  }
  CK(ctx, cudaSetDevice(ctx->device));
  std::vector<double*> ptrs(c.world, nullptr);
  for (int q = 0; q < c.world; ++q) {
    if (q == c.rank) {
      ptrs[q] = c.xchg;
      continue;
    }
    cudaIpcMemHandle_t h;
    memcpy(&h, handles + (size_t)q * 64, 64);
    void* p = nullptr;
    CK(ctx, cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    c.peer_base[q] = p;
    ptrs[q] = (double*)p;
  }
  if (!c.peers_dev) CK(ctx, cudaMalloc((void**)&c.peers_dev, 16 * sizeof(double*)));
  CK(ctx, cudaMemcpy(c.peers_dev, ptrs.data(), c.world * sizeof(double*), cudaMemcpyHostToDevice));
  // A second open on a live buffer must not see flags of the previous life: clear the flag area,
  // then a barrier so that no rank publishes into a buffer its owner has not cleared yet.
  CK(ctx, cudaMemset(c.xchg, 0, xchg_bytes(c.world)));  // flags AND the flag-in-data areas
  CK(ctx, cudaDeviceSynchronize());
  c.xchg_seq = 0;
  c.p2p = true;
  return comm_xchg_resync(ctx);
}

// Every rank leaves with the same exchange counter (max over the ranks + 2: sequence numbers only
// grow, so a flag of an abandoned exchange can never match a future one).  Called when the peer
// buffers are opened and at the start of every IRLBA run: a rank-local error exit (linear
// dependency, abort callback, allocation failure) leaves the counters apart, and without this
// every later exchange would wait for its time-out.  The allreduce doubles as the barrier.
int comm_xchg_resync(m3b_ctx* ctx) {
  Comm& c = ctx->comm;
  if (c.world == 1 || !c.p2p) return M3B_OK;
  long long* d = nullptr;
  CK(ctx, cudaMalloc((void**)&d, sizeof(long long)));
  long long h = (long long)c.xchg_seq;
  cudaError_t e = cudaMemcpyAsync(d, &h, sizeof(h), cudaMemcpyHostToDevice, ctx->stream);
  int st = M3B_OK;
  if (e == cudaSuccess) st = comm_allreduce_max_i64(ctx, d, 1);
  if (e == cudaSuccess && st == M3B_OK) e = cudaMemcpyAsync(&h, d, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream);
  if (e == cudaSuccess && st == M3B_OK) e = cudaStreamSynchronize(ctx->stream);
  cudaFree(d);
  if (st != M3B_OK) return st;
  if (e != cudaSuccess) {
    m3b_set_error(ctx, "comm_xchg_resync: %s", cudaGetErrorString(e));
    return M3B_E_CUDA;
  }
  c.xchg_seq = (unsigned long long)h + 2;
  c.ll_launch[0] = c.ll_launch[1] = 0;  // the allreduce above was a barrier: nobody is inside an exchange
  return M3B_OK;
}

void comm_destroy(m3b_ctx* ctx) {
  for (int q = 0; q < 16; ++q)
    if (ctx->comm.peer_base[q]) {
      cudaIpcCloseMemHandle(ctx->comm.peer_base[q]);
      ctx->comm.peer_base[q] = nullptr;
    }
  if (ctx->comm.peers_dev) cudaFree(ctx->comm.peers_dev);
  if (ctx->comm.xchg) cudaFree(ctx->comm.xchg);
  ctx->comm.peers_dev = nullptr;
  ctx->comm.xchg = nullptr;
  ctx->comm.p2p = false;
  if (ctx->comm.comm) {
    NcclApi* api = nccl_api(ctx);
    if (api) api->CommDestroy((ncclComm_t)ctx->comm.comm);
    ctx->comm.comm = nullptr;
  }
  ctx->comm.world = 1;
  ctx->comm.rank = 0;
}

int comm_allreduce_f64(m3b_ctx* ctx, double* dev, size_t n) {
  if (ctx->comm.world == 1 || n == 0) return M3B_OK;
  NcclApi* api = nccl_api(ctx);
  if (!api) return M3B_E_NCCL;
  NK(ctx, api, api->AllReduce(dev, dev, n, ncclDouble, ncclSum, (ncclComm_t)ctx->comm.comm, ctx->stream));
  return M3B_OK;
}

int comm_allreduce_max_i64(m3b_ctx* ctx, long long* dev, size_t n) {
  if (ctx->comm.world == 1 || n == 0) return M3B_OK;
  NcclApi* api = nccl_api(ctx);
  if (!api) return M3B_E_NCCL;
  NK(ctx, api, api->AllReduce(dev, dev, n, ncclInt64, ncclMax, (ncclComm_t)ctx->comm.comm, ctx->stream));
  return M3B_OK;
}

int comm_allreduce_i64(m3b_ctx* ctx, long long* dev, size_t n) {
  if (ctx->comm.world == 1 || n == 0) return M3B_OK;
  NcclApi* api = nccl_api(ctx);
  if (!api) return M3B_E_NCCL;
  NK(ctx, api, api->AllReduce(dev, dev, n, ncclInt64, ncclSum, (ncclComm_t)ctx->comm.comm, ctx->stream));
  return M3B_OK;
}
