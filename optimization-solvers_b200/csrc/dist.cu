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

// in-place sum over ranks (sample-sharded logistic regression: loss, gradient, Hessian)
void ctx_all_reduce_sum(Ctx* ctx, double* buf, int64_t count) {
  if (ctx->world <= 1) return;
  OSB_NCCL(nccl().AllReduce(buf, buf, (size_t)count, ncclFloat64, ncclSum, (ncclComm_t)ctx->nccl_comm, ctx->stream));
  ctx->counters[4]++;
}

}  // namespace osb
