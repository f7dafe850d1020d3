// Batched symmetric eigensolver for the many small per-l Fock matrices of the batched atomic SCFs (SAP workload:
// 86 atoms x 4 angular momenta x 69 radial functions) -- caller-side infrastructure (the
// reference leaves the eigenproblems to OpenOrbitalOptimizer / Eigen).  cuSOLVER's batched Jacobi path stops at
// n = 32 and the generic path solves the matrices one after another; here one CTA owns one matrix: the matrix and the
// accumulated rotations live in shared memory, and each step of a cyclic two-sided Jacobi sweep applies the
// floor(n/2) disjoint rotations of one round-robin round in parallel.
#include <cuda_runtime.h>

#include <cstdint>
#include <stdexcept>
#include <string>

namespace hfq {
namespace {

// A: [nb][n][n] symmetric (overwritten: V[b][i][j] = component i of eigenvector j, row-major), W: [nb][n] (unsorted)
__global__ void __launch_bounds__(256) k_jacobi_batch(double *__restrict__ A, double *__restrict__ W, int n, int max_sweeps) {
  extern __shared__ double sm[];
  const int ld = n | 1;                       // odd leading dimension: conflict-free column walks
  double *sA = sm, *sV = sA + (size_t)n * ld;
  double *sc = sV + (size_t)n * ld;           // per pair: c, s
  int *spq = reinterpret_cast<int *>(sc + 2 * ((n + 1) / 2));
  __shared__ double red[8];
  __shared__ double s_off, s_diag;
  const int tid = threadIdx.x, nt = blockDim.x;
  double *Ab = A + (size_t)blockIdx.x * n * n;
  for (int idx = tid; idx < n * n; idx += nt) {
    const int i = idx / n, j = idx % n;
    sA[i * ld + j] = 0.5 * (Ab[idx] + Ab[(size_t)j * n + i]);   // symmetrise the input
    sV[i * ld + j] = i == j ? 1.0 : 0.0;
  }
  __syncthreads();
  const int np = (n + 1) / 2, N2 = 2 * np;    // players of the round-robin tournament (a dummy one if n is odd)
  for (int sweep = 0; sweep < max_sweeps; sweep++) {
    // convergence: off-diagonal norm against the diagonal
    double off = 0.0, dg = 0.0;
    for (int idx = tid; idx < n * n; idx += nt) {
      const int i = idx / n, j = idx % n;
      const double v = sA[i * ld + j];
      if (i == j) dg += v * v; else off += v * v;
    }
    for (int o = 16; o > 0; o >>= 1) {
      off += __shfl_down_sync(0xffffffffu, off, o);
      dg += __shfl_down_sync(0xffffffffu, dg, o);
    }
    if ((tid & 31) == 0) red[tid >> 5] = off;
    __syncthreads();
    if (tid == 0) {
      double s = 0.0;
      for (int w = 0; w < (nt >> 5); w++) s += red[w];
      s_off = s;
    }
    __syncthreads();
    if ((tid & 31) == 0) red[tid >> 5] = dg;
    __syncthreads();
    if (tid == 0) {
      double s = 0.0;
      for (int w = 0; w < (nt >> 5); w++) s += red[w];
      s_diag = s;
    }
    __syncthreads();
    if (s_off <= 1e-30 * s_diag || s_off == 0.0) break;
    for (int r = 0; r < N2 - 1; r++) {
      if (tid < np) {
        int a = (r + tid) % (N2 - 1), b = tid == 0 ? N2 - 1 : (r - tid + N2 - 1) % (N2 - 1);
        int p = min(a, b), q = max(a, b);
        double c = 1.0, s = 0.0;
        if (q < n) {
          const double apq = sA[p * ld + q];
          if (apq != 0.0) {
            const double theta = (sA[q * ld + q] - sA[p * ld + p]) / (2.0 * apq);
            const double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(1.0 + theta * theta));
            c = 1.0 / sqrt(1.0 + t * t);
            s = t * c;
          }
        } else {
          q = -1;
        }
        sc[2 * tid] = c;
        sc[2 * tid + 1] = s;
        spq[2 * tid] = p;
        spq[2 * tid + 1] = q;
      }
      __syncthreads();
      // A <- A J, V <- V J  (columns p, q of every row)
      for (int idx = tid; idx < np * n; idx += nt) {
        const int k = idx / n, row = idx % n, p = spq[2 * k], q = spq[2 * k + 1];
        if (q < 0) continue;
        const double c = sc[2 * k], s = sc[2 * k + 1];
        double *a = sA + row * ld, *v = sV + row * ld;
        const double ap = a[p], aq = a[q], vp = v[p], vq = v[q];
        a[p] = c * ap - s * aq;
        a[q] = s * ap + c * aq;
        v[p] = c * vp - s * vq;
        v[q] = s * vp + c * vq;
      }
      __syncthreads();
      // A <- J^T A  (rows p, q)
      for (int idx = tid; idx < np * n; idx += nt) {
        const int k = idx / n, col = idx % n, p = spq[2 * k], q = spq[2 * k + 1];
        if (q < 0) continue;
        const double c = sc[2 * k], s = sc[2 * k + 1];
        const double ap = sA[p * ld + col], aq = sA[q * ld + col];
        sA[p * ld + col] = c * ap - s * aq;
        sA[q * ld + col] = s * ap + c * aq;
      }
      __syncthreads();
    }
  }
  for (int i = tid; i < n; i += nt) W[(size_t)blockIdx.x * n + i] = sA[i * ld + i];
  for (int idx = tid; idx < n * n; idx += nt) Ab[idx] = sV[(idx / n) * ld + idx % n];
}

}  // namespace

void syev_batch(double *dA, double *dW, int n, int64_t nb, cudaStream_t st) {
  if (n < 1 || nb < 1) return;
  const int ld = n | 1;
  const size_t smem = ((size_t)2 * n * ld + 2 * ((n + 1) / 2)) * sizeof(double) + 2 * ((n + 1) / 2) * sizeof(int) + 16;
  if (smem > 227 * 1024) throw std::logic_error("hfq_syev_batch: matrix too large for the shared-memory Jacobi solver (n <= 118)");
  cudaError_t e = cudaFuncSetAttribute(k_jacobi_batch, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) throw std::runtime_error(std::string("CUDA error: ") + cudaGetErrorString(e));
  k_jacobi_batch<<<(unsigned)nb, 256, smem, st>>>(dA, dW, n, 30);
  e = cudaGetLastError();
  if (e != cudaSuccess) throw std::runtime_error(std::string("CUDA error: ") + cudaGetErrorString(e));
}

}  // namespace hfq
