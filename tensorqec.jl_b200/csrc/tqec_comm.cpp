// The one collective of the hot path: sum of the four logical-error counters over the ranks of a job (one process per
// GPU), by ncclAllReduce over NVLink / NVSwitch.  Replaces the result channel of the reference's job farm
// (src/multiprocessing.jl:41-52: `take!(results)` per job).  NCCL is bound at run time (dlopen of libnccl.so.2): inside
// a PyTorch process this resolves to the copy torch already loaded, elsewhere to the system library; a host that never
// creates a communicator never needs NCCL.
#include <dlfcn.h>
#include <nccl.h>

#include <cstring>

#include "tqec_common.h"

using namespace tqec;

struct tqec_comm {
  int device, nranks, rank;
  ncclComm_t comm;
  cudaStream_t stream;
  unsigned long long *d_buf;
};

namespace {
struct NcclApi {
  void *handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
};

NcclApi *nccl() {
  static NcclApi api;
  static bool tried = false;
  if (!tried) {
    tried = true;
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char *n : names) {
      api.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
      if (api.handle) break;
    }
    if (api.handle) {
      api.GetUniqueId = (decltype(api.GetUniqueId))dlsym(api.handle, "ncclGetUniqueId");
      api.CommInitRank = (decltype(api.CommInitRank))dlsym(api.handle, "ncclCommInitRank");
      api.AllReduce = (decltype(api.AllReduce))dlsym(api.handle, "ncclAllReduce");
      api.CommDestroy = (decltype(api.CommDestroy))dlsym(api.handle, "ncclCommDestroy");
      api.GetErrorString = (decltype(api.GetErrorString))dlsym(api.handle, "ncclGetErrorString");
      if (!api.GetUniqueId || !api.CommInitRank || !api.AllReduce || !api.CommDestroy || !api.GetErrorString) {
        dlclose(api.handle);
        api.handle = nullptr;
      }
    }
  }
  return api.handle ? &api : nullptr;
}
}  // namespace

#define TQEC_NCCL(call)                                                                            \
  do {                                                                                             \
    ncclResult_t _r = (call);                                                                      \
    if (_r != ncclSuccess) {                                                                       \
      set_error("%s failed: %s", #call, api->GetErrorString(_r));                                  \
      return TQEC_ERR_CUDA;                                                                        \
    }                                                                                              \
  } while (0)

extern "C" int tqec_comm_unique_id(void *id_out) {
  TQEC_REQUIRE(id_out != nullptr, "tqec_comm_unique_id: NULL argument");
  NcclApi *api = nccl();
  if (!api) { set_error("tqec_comm_unique_id: libnccl.so.2 not found (%s)", dlerror() ? dlerror() : "no loader message"); return TQEC_ERR_UNSUPPORTED; }
  ncclUniqueId id;
  TQEC_NCCL(api->GetUniqueId(&id));
  std::memcpy(id_out, &id, sizeof(id));
  return TQEC_OK;
}

extern "C" int tqec_comm_init(int32_t nranks, int32_t rank, const void *unique_id, int32_t device, tqec_comm **out) {
  TQEC_REQUIRE(out && unique_id && nranks >= 1 && rank >= 0 && rank < nranks, "tqec_comm_init: bad arguments (nranks=%d rank=%d)", nranks, rank);
  *out = nullptr;
  NcclApi *api = nccl();
  if (!api) { set_error("tqec_comm_init: libnccl.so.2 not found"); return TQEC_ERR_UNSUPPORTED; }
  TQEC_CUDA(cudaSetDevice(device));
  ncclUniqueId id;
  std::memcpy(&id, unique_id, sizeof(id));
  tqec_comm *c = new tqec_comm();
  std::memset(c, 0, sizeof(*c));
  c->device = device; c->nranks = nranks; c->rank = rank;
  ncclResult_t r = api->CommInitRank(&c->comm, nranks, id, rank);
  if (r != ncclSuccess) { set_error("ncclCommInitRank failed: %s", api->GetErrorString(r)); delete c; return TQEC_ERR_CUDA; }
  cudaError_t e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaMalloc((void **)&c->d_buf, 32);
  if (e != cudaSuccess) { set_error("tqec_comm_init: %s", cudaGetErrorString(e)); tqec_comm_destroy(c); return TQEC_ERR_CUDA; }
  *out = c;
  return TQEC_OK;
}

extern "C" int tqec_comm_destroy(tqec_comm *c) {
  if (!c) return TQEC_OK;
  cudaSetDevice(c->device);
  NcclApi *api = nccl();
  if (c->comm && api) api->CommDestroy(c->comm);
  if (c->d_buf) cudaFree(c->d_buf);
  if (c->stream) cudaStreamDestroy(c->stream);
  delete c;
  return TQEC_OK;
}

namespace tqec {
// in-place sum of four 64-bit counters on `stream` (device buffer)
int comm_allreduce_dev(tqec_comm *c, unsigned long long *d_counts, cudaStream_t stream) {
  NcclApi *api = nccl();
  if (!api) { set_error("allreduce: libnccl.so.2 not found"); return TQEC_ERR_UNSUPPORTED; }
  TQEC_NCCL(api->AllReduce(d_counts, d_counts, 4, ncclUint64, ncclSum, c->comm, stream));
  return TQEC_OK;
}
}  // namespace tqec

extern "C" int tqec_comm_allreduce_counts(tqec_comm *c, int64_t counts[4]) {
  TQEC_REQUIRE(c && counts, "tqec_comm_allreduce_counts: NULL argument");
  TQEC_CUDA(cudaSetDevice(c->device));
  TQEC_CUDA(cudaMemcpyAsync(c->d_buf, counts, 32, cudaMemcpyHostToDevice, c->stream));
  int rc = comm_allreduce_dev(c, c->d_buf, c->stream);
  if (rc) return rc;
  TQEC_CUDA(cudaMemcpyAsync(counts, c->d_buf, 32, cudaMemcpyDeviceToHost, c->stream));
  TQEC_CUDA(cudaStreamSynchronize(c->stream));
  return TQEC_OK;
}
