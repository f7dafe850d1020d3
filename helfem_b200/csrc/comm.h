// NCCL communicator of a multi-GPU build (one process per GPU).  The library is bound at run time with
// dlopen -- an embedding process (e.g. PyTorch) usually has its own libnccl.so.2 loaded already, and that copy is the
// one used -- so libhelfemqc_b200.so has no link-time dependency on NCCL and single-GPU users never load it.
#pragma once
#include <cuda_runtime.h>

#include <cstddef>
#include <cstdint>

namespace hfq {

constexpr int kCommIdBytes = 128;   // sizeof(ncclUniqueId)

class Comm {
 public:
  // rank 0 creates the id (ncclGetUniqueId); the caller distributes it to all ranks (MPI_Bcast, torch.distributed ...)
  static void unique_id(void *out128);
  Comm(const void *id128, int rank, int nranks, int device);
  ~Comm();
  Comm(const Comm &) = delete;
  Comm &operator=(const Comm &) = delete;
  int rank() const { return rank_; }
  int size() const { return nranks_; }
  // in place: every rank contributes buf[rank*count .. +count) and receives all nranks*count doubles
  void all_gather_inplace(double *buf, size_t count, cudaStream_t st);
  void all_reduce_sum(double *buf, size_t count, cudaStream_t st);
  void all_reduce_max_u64(unsigned long long *buf, size_t count, cudaStream_t st);
  void broadcast(void *buf, size_t bytes, int root, cudaStream_t st);

 private:
  void *comm_ = nullptr;
  int rank_ = 0, nranks_ = 1;
};

}  // namespace hfq
