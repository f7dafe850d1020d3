// J/K engine: device-resident integral caches + the CUDA Fock-build path.
//
// One Engine per (basis, GPU).  It owns the device copies of the integral
// caches and coupling tables, and evaluates
//     J = coulomb(P),  K = exchange(P)
// for dense column-major FP64 density matrices, with the semantics of the
// reference's TwoDBasis::coulomb / TwoDBasis::exchange
// (src/atomic/TwoDBasis.cpp:773-999, src/diatomic/basis.cpp:1627-2089):
// exchange() returns -K, boundary functions of m != 0 shells are removed
// (diatomic), +-m mirroring is available through set_absm_symmetric().
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <functional>
#include <memory>
#include <string>
#include <vector>

#include "tables.h"

namespace hfq {

class Comm;

struct EngineTimings {      // milliseconds of the last call, CUDA events
  float pack = 0, fold = 0, tgemm = 0, offdiag = 0, unpack = 0, total = 0;
  double flops_fold = 0, flops_tgemm = 0, flops_offdiag = 0;  // executed (padded) flops
  double alg_fold = 0, alg_tgemm = 0, alg_offdiag = 0;        // algorithmic (unpadded) flops, DESIGN.md
  int launches = 0;           // kernel launches of the call
  int launches_fold = 0, launches_tgemm = 0, launches_offdiag = 0;
  double h2d_bytes = 0, d2h_bytes = 0;   // host-pointer entry points: bytes moved over PCIe by the call
};

// Longest-processing-time assignment of work units to ranks (owner-computes sharding of the exchange build);
// deterministic, so every rank derives the same ownership from the same plan.
void assign_units(const std::vector<double> &cost, int nranks, std::vector<int> &owner);

class Engine {
 public:
  Engine(const BasisTables &t, int device);
  ~Engine();
  Engine(const Engine &) = delete;
  Engine &operator=(const Engine &) = delete;

  int Nbf() const { return nbf_; }
  const BasisTables &tables() const;   // host description (integral blocks released after upload)
  int device() const { return device_; }
  // Multi-GPU build, one process per GPU: bind this engine to an NCCL communicator (id from Comm::unique_id on rank
  // 0, distributed by the caller).  From then on exchange_dev / jk_dev shard the exchange by output block over the
  // ranks (owner computes) and complete K with one in-place all-gather; J is built in full on every rank.
  void set_comm(const void *id128, int rank, int nranks);
  int comm_size() const;
  void set_absm_symmetric(bool s) { absm_symmetric_ = s; }
  bool absm_symmetric() const { return absm_symmetric_; }

  // Device-pointer entry points: P, J, K are Nbf x Nbf column-major on this
  // engine's device.  shard/nshards partition the exchange output blocks over
  // ranks (blocks owned by other shards are written as zero so that a sum
  // over ranks gives the full matrix).
  void coulomb_dev(const double *dP, int64_t ldP, double *dJ, int64_t ldJ, int shard, int nshards, cudaStream_t stream);
  // Fused build: J = coulomb(P), K = exchange(kscale * P) from one packed copy of P.  With nshards > 1
  // both outputs are partial sums (J over multipoles, K over tasks) to be all-reduced by the caller.
  void jk_dev(const double *dP, int64_t ldP, double kscale, double *dJ, int64_t ldJ, double *dK, int64_t ldK, int shard,
              int nshards, cudaStream_t stream);
  void exchange_dev(const double *dP, int64_t ldP, double *dK, int64_t ldK, int shard, int nshards,
                    cudaStream_t stream);
  // batched radial Coulomb (SAP workload): J_b = fac * prefactor(L=0) * J_0(P_b), b < nb, device-resident
  void coulomb_radial_batch(const double *dP, double *dJ, int nb, int64_t stride, double fac, cudaStream_t stream);
  // Host-pointer entry points (copies in/out on the engine's stream).
  void coulomb(const double *P, int64_t ldP, double *J, int64_t ldJ);
  void exchange(const double *P, int64_t ldP, double *K, int64_t ldK);
  void coulomb_exchange(const double *P, int64_t ldP, double kscale, double *J, int64_t ldJ, double *K, int64_t ldK);
  // multi-GPU, host matrices shared by all ranks: every rank moves its column slice (see engine.cu)
  void jk_spmd_host(const double *P, int64_t ldP, double kscale, double *J, int64_t ldJ, double *K, int64_t ldK);
  // device copy of the density of the last host-pointer call (Nbf x Nbf, ld = Nbf)
  const double *device_density() const;
  int speculative_hits() const { return spec_hits_; }   // fused host calls that ran on a predicted sparse upload

  // Non-zero structure of the last exchange result: sector id of every dense basis function and
  // the (row sector, column sector) pairs that were written; all other blocks of K are zero.
  void output_pattern(std::vector<int> &bf_sector, std::vector<int> &pairs, bool coulomb = false) const;

  // make `waiter` wait for everything queued on `on` so far
  void fence_stream(cudaStream_t on, cudaStream_t waiter);
  const EngineTimings &timings() const { return tm_; }
  cudaStream_t stream() const { return stream_; }
  size_t device_bytes() const { return dev_bytes_; }

 private:
  void pack_density(const double *dP, int64_t ldP, cudaStream_t stream);
  void coulomb_run(const double *dP, int64_t ldP, double *dJ, int64_t ldJ, int shard, int nshards, cudaStream_t stream,
                   bool async);
  // Host-pointer results: only the bounding row range of the blocks that can be non-zero is copied
  // back per column (grouped into rectangles); the rest of the caller's matrix is zero-filled by
  // host threads while the GPU is still computing.
  struct HostRanges {
    std::vector<int> r0, r1;   // per dense column: rows [r0, r1) are copied, the rest is zero
  };
  HostRanges host_ranges(bool coulomb) const;
  // columns [cb, ce) only (ce < 0: all)
  double copy_ranges_async(double *H, int64_t ldH, const double *D, const HostRanges &hr, cudaStream_t st, int cb = 0,
                           int ce = -1) const;   // returns bytes
  static void zero_outside(double *H, int64_t ldH, int n, const HostRanges &hr, int cb = 0, int ce = -1);
  static bool nonzero_outside(const double *H, int64_t ldH, int n, const std::vector<int> &r0, const std::vector<int> &r1);
  bool fused_host(const double *P, int64_t ldP, double kscale, double *J, int64_t ldJ, double *K, int64_t ldK, bool spec);
  HostRanges density_ranges() const;   // bounding non-zero row range per column of the density packed last
  int spec_hits_ = 0;
  std::function<void()> plan_hook_;   // called by exchange_dev once the output pattern is known
  struct Impl;
  struct PlanCache;
  std::unique_ptr<Impl> p_;
  std::unique_ptr<PlanCache> plans_;
  std::unique_ptr<Comm> comm_;
  std::vector<int> last_active_ops_, last_active_j_;
  int device_ = 0, nbf_ = 0;
  bool absm_symmetric_ = false;
  cudaStream_t stream_ = nullptr;
  EngineTimings tm_;
  size_t dev_bytes_ = 0;
};

// host threads for the zero-fill / copy helpers (hfq_set_host_threads); 0 = OpenMP default
void set_host_threads(int n);
int host_threads();

}  // namespace hfq
