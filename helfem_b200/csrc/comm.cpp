#include "comm.h"

#include <dlfcn.h>
#include <nccl.h>

#include <cstdlib>
#include <cstring>
#include <mutex>
#include <stdexcept>
#include <string>

namespace hfq {
namespace {

struct Api {
  ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Broadcast)(const void *, void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
};

const Api &api() {
  static Api a;
  static std::once_flag once;
  std::call_once(once, [] {
    void *h = nullptr;
    if (const char *env = getenv("HFQ_NCCL_LIB")) h = dlopen(env, RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);   // the copy the embedding process already uses
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) throw std::runtime_error(std::string("hfq_comm: cannot load libnccl.so.2: ") + dlerror());
    auto sym = [&](const char *name) {
      void *p = dlsym(h, name);
      if (!p) throw std::runtime_error(std::string("hfq_comm: libnccl has no symbol ") + name);
      return p;
    };
    a.GetUniqueId = reinterpret_cast<decltype(a.GetUniqueId)>(sym("ncclGetUniqueId"));
    a.CommInitRank = reinterpret_cast<decltype(a.CommInitRank)>(sym("ncclCommInitRank"));
    a.CommDestroy = reinterpret_cast<decltype(a.CommDestroy)>(sym("ncclCommDestroy"));
    a.AllGather = reinterpret_cast<decltype(a.AllGather)>(sym("ncclAllGather"));
    a.AllReduce = reinterpret_cast<decltype(a.AllReduce)>(sym("ncclAllReduce"));
    a.Broadcast = reinterpret_cast<decltype(a.Broadcast)>(sym("ncclBroadcast"));
    a.GetErrorString = reinterpret_cast<decltype(a.GetErrorString)>(sym("ncclGetErrorString"));
  });
  return a;
}

void check(ncclResult_t r, const char *what) {
  if (r != ncclSuccess) throw std::runtime_error(std::string("NCCL error in ") + what + ": " + api().GetErrorString(r));
}

}  // namespace

void Comm::unique_id(void *out128) {
  static_assert(sizeof(ncclUniqueId) == kCommIdBytes, "ncclUniqueId size");
  check(api().GetUniqueId(reinterpret_cast<ncclUniqueId *>(out128)), "ncclGetUniqueId");
}

Comm::Comm(const void *id128, int rank, int nranks, int device) : rank_(rank), nranks_(nranks) {
  if (nranks < 1 || rank < 0 || rank >= nranks) throw std::logic_error("hfq_comm_init: bad rank / size");
  if (cudaSetDevice(device) != cudaSuccess) throw std::runtime_error("CUDA error: cudaSetDevice in hfq_comm_init");
  ncclUniqueId id;
  memcpy(&id, id128, sizeof(id));
  ncclComm_t c = nullptr;
  check(api().CommInitRank(&c, nranks, id, rank), "ncclCommInitRank");
  comm_ = c;
}

Comm::~Comm() {
  if (comm_) api().CommDestroy(reinterpret_cast<ncclComm_t>(comm_));
}

void Comm::all_gather_inplace(double *buf, size_t count, cudaStream_t st) {
  check(api().AllGather(buf + (size_t)rank_ * count, buf, count, ncclDouble, reinterpret_cast<ncclComm_t>(comm_), st),
        "ncclAllGather");
}
void Comm::all_reduce_sum(double *buf, size_t count, cudaStream_t st) {
  check(api().AllReduce(buf, buf, count, ncclDouble, ncclSum, reinterpret_cast<ncclComm_t>(comm_), st), "ncclAllReduce");
}
void Comm::all_reduce_max_u64(unsigned long long *buf, size_t count, cudaStream_t st) {
  check(api().AllReduce(buf, buf, count, ncclUint64, ncclMax, reinterpret_cast<ncclComm_t>(comm_), st), "ncclAllReduce");
}
void Comm::broadcast(void *buf, size_t bytes, int root, cudaStream_t st) {
  check(api().Broadcast(buf, buf, bytes, ncclUint8, root, reinterpret_cast<ncclComm_t>(comm_), st), "ncclBroadcast");
}

}  // namespace hfq
