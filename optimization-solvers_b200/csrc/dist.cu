// dist.cu — multi-GPU plumbing: one process per GPU, H row-block sharded, O(n) vectors replicated.
// The only exchange step of a dense quasi-Newton iteration is the all-gather of the row-block
// slices of h = H y and u = H' g (n/P doubles per rank).  NCCL is bound at run time (dlopen) so
// that single-GPU use needs no NCCL at all.
#include <dlfcn.h>
#include <cstring>

#include "engine.cuh"

namespace osb {

typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
enum { ncclFloat64 = 8 };
enum { ncclSum = 0 };

struct NcclApi {
  void* handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

static NcclApi& nccl() {
  static NcclApi api;
  if (!api.handle) {
    api.handle = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!api.handle) api.handle = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!api.handle) throw Error(OSB_ERR_NCCL, std::string("cannot load libnccl.so.2: ") + dlerror());
    api.GetUniqueId = (decltype(api.GetUniqueId))dlsym(api.handle, "ncclGetUniqueId");
    api.CommInitRank = (decltype(api.CommInitRank))dlsym(api.handle, "ncclCommInitRank");
    api.CommDestroy = (decltype(api.CommDestroy))dlsym(api.handle, "ncclCommDestroy");
    api.AllGather = (decltype(api.AllGather))dlsym(api.handle, "ncclAllGather");
    api.AllReduce = (decltype(api.AllReduce))dlsym(api.handle, "ncclAllReduce");
    api.GetErrorString = (decltype(api.GetErrorString))dlsym(api.handle, "ncclGetErrorString");
    if (!api.GetUniqueId || !api.CommInitRank || !api.AllGather) throw Error(OSB_ERR_NCCL, "libnccl is missing required symbols");
  }
  return api;
}

#define OSB_NCCL(expr)                                                                                   \
  do {                                                                                                   \
    ncclResult_t _r = (expr);                                                                            \
    if (_r != 0) throw Error(OSB_ERR_NCCL, std::string(#expr) + " failed: " + (nccl().GetErrorString ? nccl().GetErrorString(_r) : "?")); \
  } while (0)

void nccl_unique_id(void* out128) {
  ncclUniqueId id;
  OSB_NCCL(nccl().GetUniqueId(&id));
  std::memcpy(out128, &id, sizeof(id));
}

void ctx_init_dist(Ctx* ctx, int rank, int world, const void* uid) {
  ncclUniqueId id;
  std::memcpy(&id, uid, sizeof(id));
  ncclComm_t comm = nullptr;
  ctx->use();
  OSB_NCCL(nccl().CommInitRank(&comm, world, id, rank));
  ctx->nccl_comm = comm;
  ctx->rank = rank;
  ctx->world = world;
}

void ctx_destroy_dist(Ctx* ctx) {
  if (ctx->nccl_comm) nccl().CommDestroy((ncclComm_t)ctx->nccl_comm);
  ctx->nccl_comm = nullptr;
}

// buf holds world * count doubles; rank r's slice [r*count, (r+1)*count) is valid on entry
void Ctx::all_gather_inplace(double* buf, int64_t count) {
  if (world <= 1) return;
  OSB_NCCL(nccl().AllGather(buf + (int64_t)rank * count, buf, (size_t)count, ncclFloat64, (ncclComm_t)nccl_comm, stream));
  counters[4]++;
}

// ---- peer-memory exchange region (CUDA IPC) -----------------------------------------------------
static size_t xchg_bytes(int world) {
  // 4 exchanged vectors | flags (64 words reserved) | sharded packed storage: 2 parities x world x {h, w} slots
  // | per-source-rank chunk flags of the fused iteration kernel (world x XFLAG2_LD 64-bit words)
  return sizeof(double) * (size_t)(xflag2_off(world) + (int64_t)world * XFLAG2_LD);
}

void ctx_ipc_export(Ctx* ctx, void* out64) {
  ctx->use();
  if (!ctx->xchg) {
    OSB_CUDA(cudaMalloc(&ctx->xchg, xchg_bytes(ctx->world)));
    OSB_CUDA(cudaMemset(ctx->xchg, 0, xchg_bytes(ctx->world)));
    OSB_CUDA(cudaMalloc(&ctx->d_seq, sizeof(unsigned long long)));
    OSB_CUDA(cudaMemset(ctx->d_seq, 0, sizeof(unsigned long long)));
  }
  cudaIpcMemHandle_t h;
  OSB_CUDA(cudaIpcGetMemHandle(&h, ctx->xchg));
  static_assert(sizeof(h) == 64, "cudaIpcMemHandle_t is 64 bytes");
  std::memcpy(out64, &h, 64);
}

void ctx_ipc_connect(Ctx* ctx, const void* handles) {
  ctx->use();
  OSB_REQUIRE(ctx->xchg != nullptr, OSB_ERROR_INPUT_PARAMS, "call osb_ctx_ipc_handle first");
  OSB_REQUIRE(ctx->world + 8 <= 64, OSB_ERROR_INPUT_PARAMS, "peer-memory exchange supports at most 56 ranks");
  std::vector<double*> ptrs(ctx->world, nullptr);
  for (int r = 0; r < ctx->world; ++r) {
    if (r == ctx->rank) {
      ptrs[r] = ctx->xchg;
      continue;
    }
    cudaIpcMemHandle_t h;
    std::memcpy(&h, (const char*)handles + 64 * r, 64);
    void* p = nullptr;
    OSB_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    ctx->peer_opened.push_back(p);
    ptrs[r] = (double*)p;
  }
  OSB_CUDA(cudaMalloc(&ctx->d_peers, sizeof(double*) * ctx->world));
  OSB_CUDA(cudaMemcpy(ctx->d_peers, ptrs.data(), sizeof(double*) * ctx->world, cudaMemcpyHostToDevice));
  ctx->p2p_ready = true;
}

void ctx_ipc_close(Ctx* ctx) {
  for (void* p : ctx->peer_opened) cudaIpcCloseMemHandle(p);
  ctx->peer_opened.clear();
  if (ctx->d_peers) cudaFree(ctx->d_peers);
  if (ctx->xchg) cudaFree(ctx->xchg);
  if (ctx->d_seq) cudaFree(ctx->d_seq);
  ctx->d_peers = nullptr;
  ctx->xchg = nullptr;
  ctx->d_seq = nullptr;
  ctx->p2p_ready = false;
}

// in-place sum over ranks (sample-sharded logistic regression: loss, gradient, Hessian)
void ctx_all_reduce_sum(Ctx* ctx, double* buf, int64_t count) {
  if (ctx->world <= 1) return;
  OSB_NCCL(nccl().AllReduce(buf, buf, (size_t)count, ncclFloat64, ncclSum, (ncclComm_t)ctx->nccl_comm, ctx->stream));
  ctx->counters[4]++;
}

}  // namespace osb
