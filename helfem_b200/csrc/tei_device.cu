// Device-side compute_tei of the diatomic basis: see tei_device.h.  The quadratures are GEMM-shaped
// (weights x basis-function products) and run on the FP64 tensor pipe through dev::k_gemm; the pivoted Cholesky runs
// one CTA per channel.
//
// Reference behaviour: src/diatomic/quadrature.cpp:188-257 (twoe_integral: outer Chebyshev rule, one inner rule per
// outer sub-interval, cumulative inner integrals), src/diatomic/basis.cpp:1382-1483 (the four cosh^k-weighted
// kernels, symmetrised 2-channel matrix), :1483-1537 (sign-aware pivoted Cholesky, relative threshold 1e-12).
#include "tei_device.h"

#include <cuda_runtime.h>

#include <cmath>
#include <stdexcept>
#include <string>

#include "kernels.cuh"

namespace hfq {

namespace {

#define CKT(x)                                                                                       \
  do {                                                                                               \
    cudaError_t e_ = (x);                                                                            \
    if (e_ != cudaSuccess) throw std::runtime_error(std::string("CUDA error: ") + cudaGetErrorString(e_) + " (" #x ")"); \
  } while (0)

template <typename T>
struct TBuf {
  T *p = nullptr;
  size_t n = 0;
  void alloc(size_t count) {
    if (count <= n) return;
    if (p) cudaFree(p);
    p = nullptr;
    CKT(cudaMalloc(&p, count * sizeof(T)));
    n = count;
  }
  void upload(const T *h, size_t count, cudaStream_t st) {
    alloc(count);
    CKT(cudaMemcpyAsync(p, h, count * sizeof(T), cudaMemcpyHostToDevice, st));
  }
  ~TBuf() {
    if (p) cudaFree(p);
  }
};

__device__ __forceinline__ double keep_normal_dev(double v) {
  const double a = fabs(v);
  return (v != 0.0 && !(a >= 2.2250738585072014e-308 && a <= 1.7976931348623157e308)) ? 0.0 : v;
}

// Associated Legendre functions P_L^M(x) for x >= 1 (Hobson convention), L = 0 .. Lhi, at npts points: the host
// recurrence (special.cpp legendre_p, src/legendre/Legendre.h) operation by operation in round-to-nearest arithmetic
// without contraction, so that the table agrees bit for bit with the host's; values that are not normal numbers are
// stored as zero (src/diatomic/quadrature.h:47-84).  P[L * npts + pt].
__global__ void k_legendre_p(int Lhi, int M, const double *__restrict__ x, int npts, double *__restrict__ P) {
  const int pt = blockIdx.x * blockDim.x + threadIdx.x;
  if (pt >= npts) return;
  for (int L = 0; L < M && L <= Lhi; L++) P[(size_t)L * npts + pt] = 0.0;
  if (M > Lhi) return;
  const double xv = x[pt];
  const double w = __dsqrt_rn(fmax(__dsub_rn(__dmul_rn(xv, xv), 1.0), 0.0));
  double pmm = 1.0;
  for (int k = 1; k <= M; k++) pmm = __dmul_rn(pmm, __dmul_rn((double)(2 * k - 1), w));
  double lo = pmm, hi = 0.0;   // out[l - 1], out[l]
  P[(size_t)M * npts + pt] = keep_normal_dev(lo);
  if (M + 1 > Lhi) return;
  hi = __dmul_rn(__dmul_rn((double)(2 * M + 1), xv), pmm);
  P[(size_t)(M + 1) * npts + pt] = keep_normal_dev(hi);
  for (int l = M + 1; l < Lhi; l++) {
    const double nx = __ddiv_rn(__dsub_rn(__dmul_rn(__dmul_rn((double)(2 * l + 1), xv), hi), __dmul_rn((double)(l + M), lo)),
                                (double)(l + 1 - M));
    lo = hi;
    hi = nx;
    P[(size_t)(l + 1) * npts + pt] = keep_normal_dev(hi);
  }
}

// BB[pt][p] = B_i(pt) B_j(pt) at the n*n sub-interval points; subbf[ip][k][q]
__global__ void k_pair_products(int n, int nbf, int np, const double *__restrict__ subbf, const int *__restrict__ pi,
                                const int *__restrict__ pj, double *__restrict__ BB) {
  const int64_t tot = (int64_t)n * n * np;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < tot; idx += (int64_t)gridDim.x * blockDim.x) {
    const int p = (int)(idx % np);
    const int64_t pt = idx / np;
    const int ip = (int)(pt / n), q = (int)(pt % n);
    const double *b = subbf + (size_t)ip * nbf * n;
    BB[idx] = b[(size_t)pi[p] * n + q] * b[(size_t)pj[p] * n + q];
  }
}

// inner-rule weights W0[c][pt] = cw[pt] P_{L_c}(pt), W2 = W0 cosh^2
__global__ void k_inner_weights(int nchan, const int *__restrict__ Lvals, int64_t npts, const double *__restrict__ cw,
                                const double *__restrict__ ch, const double *__restrict__ P, double *__restrict__ W0,
                                double *__restrict__ W2) {
  const int64_t tot = (int64_t)nchan * npts;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < tot; idx += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(idx / npts);
    const int64_t pt = idx % npts;
    const double w0 = cw[pt] * P[(size_t)Lvals[c] * npts + pt];
    W0[idx] = w0;
    W2[idx] = w0 * ch[pt] * ch[pt];
  }
}

// cumulative inner integrals: S[(ch * n + ip)][c][p]  ->  inner[(c * 2 + ch)][ip][p] = sum_{ip' <= ip} S
__global__ void k_inner_prefix(int n, int nchan, int np, const double *__restrict__ S, double *__restrict__ inner) {
  const int tot = nchan * 2 * np;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= tot) return;
  const int p = idx % np, cch = idx / np, ch = cch % 2, c = cch / 2;
  double acc = 0.0;
  for (int ip = 0; ip < n; ip++) {
    acc += S[(((size_t)ch * n + ip) * nchan + c) * np + p];
    inner[((size_t)(c * 2 + ch) * n + ip) * np + p] = acc;
  }
}

// outer-rule weights wb[(c * 2 + k)][r][q] = uw[q] Q_{L_c}(q) cosh(mu_q)^(2k) * bfprod[r][q]
__global__ void k_outer_weights(int nchan, const int *__restrict__ Lvals, int n, int np, const double *__restrict__ uw,
                                const double *__restrict__ chmu, const double *__restrict__ Q,
                                const double *__restrict__ bfprod, double *__restrict__ wb) {
  const int64_t tot = (int64_t)nchan * 2 * np * n;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < tot; idx += (int64_t)gridDim.x * blockDim.x) {
    const int q = (int)(idx % n);
    const int r = (int)((idx / n) % np);
    const int ck = (int)(idx / ((int64_t)n * np)), k = ck % 2, c = ck / 2;
    double v = uw[q] * Q[(size_t)Lvals[c] * n + q] * bfprod[(size_t)r * n + q];
    if (k) v *= chmu[q] * chmu[q];
    wb[idx] = v;
  }
}

// W[c] (2 nn x 2 nn, symmetric) from the four outer products O[(c * 4 + k * 2 + l)][pr][pc]:
//   T00 = O00 + O00^T, T02 = O02 + O20^T, T22 = O22 + O22^T in the pair indices; W = [[T00, -T02], [-T02^T, T22]]
__global__ void k_assemble_W(int nbf, int np, const int *__restrict__ pairof, const double *__restrict__ O,
                             double *__restrict__ W) {
  const int nn = nbf * nbf, N = 2 * nn, c = blockIdx.y;
  const double *Oc = O + (size_t)c * 4 * np * np;
  double *Wc = W + (size_t)c * N * N;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < N * N; idx += gridDim.x * blockDim.x) {
    const int row = idx / N, col = idx % N;
    const int br = row / nn, bc = col / nn, r = row % nn, cc = col % nn;
    const int pr = pairof[r], pc = pairof[cc];
    double v;
    if (br == 0 && bc == 0)
      v = Oc[(size_t)0 * np * np + pr * np + pc] + Oc[(size_t)0 * np * np + pc * np + pr];
    else if (br == 1 && bc == 1)
      v = Oc[(size_t)3 * np * np + pr * np + pc] + Oc[(size_t)3 * np * np + pc * np + pr];
    else if (br == 0)   // (r, nn + cc): -T02(r, cc) = -(O02[pr, pc] + O20[pc, pr])
      v = -(Oc[(size_t)1 * np * np + pr * np + pc] + Oc[(size_t)2 * np * np + pc * np + pr]);
    else                // (nn + r, cc): -T02(cc, r)
      v = -(Oc[(size_t)1 * np * np + pc * np + pr] + Oc[(size_t)2 * np * np + pr * np + pc]);
    Wc[idx] = v;
  }
}

// Sign-aware pivoted Cholesky W ~= B diag(sigma) B^T, one CTA per channel: pivot on the largest |residual diagonal|
// (lowest index on ties, like the host's strict comparison), stop at thresh * initial maximum.
// Bout[c][p][i] (column p), sig[c][p], rank[c]; maxrank columns are available.
__global__ void __launch_bounds__(256) k_sign_cholesky(int N, int maxrank, double thresh, const double *__restrict__ W,
                                                        double *__restrict__ Bout, double *__restrict__ sig,
                                                        int *__restrict__ rank) {
  extern __shared__ double sm[];
  double *d = sm;                 // [N] residual diagonal
  double *bpiv = sm + N;          // [maxrank] sigma_q * B_q[piv]
  __shared__ double red_v[8];
  __shared__ int red_i[8];
  __shared__ double s_dmax0, s_dpiv;
  __shared__ int s_piv;
  const int c = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const double *Wc = W + (size_t)c * N * N;
  double *Bc = Bout + (size_t)c * maxrank * N;
  for (int i = tid; i < N; i += 256) d[i] = Wc[(size_t)i * N + i];
  __syncthreads();
  auto argmax = [&]() {
    double bv = -1.0;
    int bi = 0x7fffffff;
    for (int i = tid; i < N; i += 256) {
      const double a = fabs(d[i]);
      if (a > bv) {
        bv = a;
        bi = i;
      }
    }
    for (int o = 16; o > 0; o >>= 1) {
      const double ov = __shfl_down_sync(0xffffffffu, bv, o);
      const int oi = __shfl_down_sync(0xffffffffu, bi, o);
      if (ov > bv || (ov == bv && oi < bi)) {
        bv = ov;
        bi = oi;
      }
    }
    if (lane == 0) {
      red_v[warp] = bv;
      red_i[warp] = bi;
    }
    __syncthreads();
    if (tid == 0) {
      for (int w = 1; w < 8; w++)
        if (red_v[w] > bv || (red_v[w] == bv && red_i[w] < bi)) {
          bv = red_v[w];
          bi = red_i[w];
        }
      s_dpiv = bv;
      s_piv = bi;
    }
    __syncthreads();
  };
  argmax();
  if (tid == 0) s_dmax0 = s_dpiv;
  __syncthreads();
  const double dmax0 = s_dmax0;
  int r = 0;
  for (; r < N && r < maxrank; r++) {
    if (r > 0) argmax();
    const double dpiv = s_dpiv;
    const int piv = s_piv;
    if (dmax0 <= 0.0 || dpiv <= thresh * dmax0) break;
    const double dp = d[piv], s = dp >= 0.0 ? 1.0 : -1.0;
    for (int q = tid; q < r; q += 256) bpiv[q] = __dmul_rn(sig[(size_t)c * maxrank + q], Bc[(size_t)q * N + piv]);
    __syncthreads();
    const double inv = __ddiv_rn(1.0, __dsqrt_rn(fabs(dp)));
    double *col = Bc + (size_t)r * N;
    for (int i = tid; i < N; i += 256) {
      double v = Wc[(size_t)i * N + piv];   // W is symmetric: row i of column piv
      for (int q = 0; q < r; q++) v = __dsub_rn(v, __dmul_rn(bpiv[q], Bc[(size_t)q * N + i]));
      v = __dmul_rn(v, inv);
      col[i] = v;
      d[i] = (i == piv) ? 0.0 : __dsub_rn(d[i], __dmul_rn(__dmul_rn(s, v), v));
    }
    if (tid == 0) sig[(size_t)c * maxrank + r] = s;
    __syncthreads();
  }
  if (tid == 0) rank[c] = r;
}

}  // namespace

struct TeiDevice::Impl {
  int device = 0;
  cudaStream_t st = nullptr;
  int n = 0, nbf = 0, np = 0, maxrank = 0;
  TBuf<double> cw, subch, subbf, uw, chmu, bfprod, BB, P, W0, W2, S, inner, wb, O, W, Q, Bout, sig;
  TBuf<int> pi, pj, pairof, Lvals, rank, bo_np;
  TBuf<dev::GemmItem> items;
  TBuf<dev::GemmEntry> entries;
  std::vector<double> hB, hsig;
  std::vector<int> hrank;
};

TeiDevice::TeiDevice(int device) : p_(new Impl) {
  p_->device = device;
  CKT(cudaSetDevice(device));
  CKT(cudaStreamCreateWithFlags(&p_->st, cudaStreamNonBlocking));
}

TeiDevice::~TeiDevice() {
  if (p_) {
    cudaSetDevice(p_->device);
    if (p_->st) cudaStreamDestroy(p_->st);
    delete p_;
  }
}

void TeiDevice::set_rule(const TeiRuleView &r) {
  Impl &s = *p_;
  CKT(cudaSetDevice(s.device));
  s.n = r.n;
  s.nbf = r.nbf;
  s.np = r.npair;
  const size_t n = r.n, npts = n * n;
  s.cw.upload(r.cw, npts, s.st);
  s.subch.upload(r.subch, npts, s.st);
  s.subbf.upload(r.subbf, npts * r.nbf, s.st);
  s.uw.upload(r.uw, n, s.st);
  s.chmu.upload(r.chmu, n, s.st);
  s.bfprod.upload(r.bfprod, n * r.npair, s.st);
  s.pi.upload(r.pi, r.npair, s.st);
  s.pj.upload(r.pj, r.npair, s.st);
  std::vector<int> pairof((size_t)r.nbf * r.nbf), bo(n);
  for (int p = 0; p < r.npair; p++) {
    pairof[r.pi[p] + r.pj[p] * r.nbf] = p;
    pairof[r.pj[p] + r.pi[p] * r.nbf] = p;
  }
  for (size_t q = 0; q < n; q++) bo[q] = (int)(q * r.npair);
  s.pairof.upload(pairof.data(), pairof.size(), s.st);
  s.bo_np.upload(bo.data(), bo.size(), s.st);
  s.BB.alloc(npts * r.npair);
  k_pair_products<<<1184, 256, 0, s.st>>>(r.n, r.nbf, r.npair, s.subbf.p, s.pi.p, s.pj.p, s.BB.p);
  CKT(cudaGetLastError());
  CKT(cudaStreamSynchronize(s.st));   // the host arrays of the caller may go away
}

void TeiDevice::run(int Mabs, int Lhi, const double *Qout, const int *Lvals, int nchan, double thresh,
                    TeiChannelResult *out) {
  Impl &s = *p_;
  CKT(cudaSetDevice(s.device));
  if (nchan <= 0) return;
  const int n = s.n, np = s.np, nbf = s.nbf, nn = nbf * nbf, N = 2 * nn;
  const size_t npts = (size_t)n * n;
  const int maxrank = std::min(N, 128);
  s.Lvals.upload(Lvals, nchan, s.st);
  s.Q.upload(Qout, (size_t)(Lhi + 1) * n, s.st);
  // 1. Legendre P at the sub-interval points
  s.P.alloc((size_t)(Lhi + 1) * npts);
  k_legendre_p<<<(unsigned)((npts + 127) / 128), 128, 0, s.st>>>(Lhi, Mabs, s.subch.p, (int)npts, s.P.p);
  CKT(cudaGetLastError());
  // 2. inner weights and the inner integrals per outer sub-interval: S[(ch, ip)][c][p] = sum_q W_ch[c][(ip, q)] BB[(ip, q)][p]
  s.W0.alloc((size_t)nchan * npts);
  s.W2.alloc((size_t)nchan * npts);
  k_inner_weights<<<1184, 256, 0, s.st>>>(nchan, s.Lvals.p, (int64_t)npts, s.cw.p, s.subch.p, s.P.p, s.W0.p, s.W2.p);
  CKT(cudaGetLastError());
  s.S.alloc((size_t)2 * n * nchan * np);
  std::vector<dev::GemmItem> items;
  std::vector<dev::GemmEntry> entries;
  for (int ch = 0; ch < 2; ch++)
    for (int ip = 0; ip < n; ip++) {
      dev::GemmItem it{};
      it.C = s.S.p + ((size_t)ch * n + ip) * nchan * np;
      it.browoff = s.bo_np.p;
      it.M = nchan;
      it.N = np;
      it.K = n;
      it.ent0 = (int)entries.size();
      entries.push_back(dev::GemmEntry{(ch ? s.W2.p : s.W0.p) + (size_t)ip * n, s.BB.p + (size_t)ip * n * np, (int64_t)npts});
      it.ent1 = (int)entries.size();
      it.accumulate = 0;
      it.ldc = np;
      it.alpha = 1.0;
      items.push_back(it);
    }
  auto launch = [&](int maxM, int maxN) {
    s.items.upload(items.data(), items.size(), s.st);
    s.entries.upload(entries.data(), entries.size(), s.st);
    const dim3 grid((maxN + 63) / 64, (maxM + 63) / 64, (unsigned)items.size());
    dev::k_gemm<64, 64, 2, 2, false><<<grid, 128, 0, s.st>>>(s.items.p, s.entries.p);
    CKT(cudaGetLastError());
    CKT(cudaStreamSynchronize(s.st));   // the descriptor vectors are rebuilt by the next stage
  };
  launch(nchan, np);
  // 3. cumulative sums over the sub-intervals
  s.inner.alloc((size_t)nchan * 2 * n * np);
  k_inner_prefix<<<(nchan * 2 * np + 127) / 128, 128, 0, s.st>>>(n, nchan, np, s.S.p, s.inner.p);
  CKT(cudaGetLastError());
  // 4. outer rule: O[(c, k, l)][pr][pc] = sum_q wb_k[c][pr][q] inner_l[c][q][pc]
  s.wb.alloc((size_t)nchan * 2 * np * n);
  k_outer_weights<<<1184, 256, 0, s.st>>>(nchan, s.Lvals.p, n, np, s.uw.p, s.chmu.p, s.Q.p, s.bfprod.p, s.wb.p);
  CKT(cudaGetLastError());
  s.O.alloc((size_t)nchan * 4 * np * np);
  items.clear();
  entries.clear();
  for (int c = 0; c < nchan; c++)
    for (int k = 0; k < 2; k++)
      for (int l = 0; l < 2; l++) {
        dev::GemmItem it{};
        it.C = s.O.p + ((size_t)c * 4 + k * 2 + l) * np * np;
        it.browoff = s.bo_np.p;
        it.M = np;
        it.N = np;
        it.K = n;
        it.ent0 = (int)entries.size();
        entries.push_back(dev::GemmEntry{s.wb.p + (size_t)(c * 2 + k) * np * n, s.inner.p + (size_t)(c * 2 + l) * n * np, (int64_t)n});
        it.ent1 = (int)entries.size();
        it.accumulate = 0;
        it.ldc = np;
        it.alpha = 1.0;
        items.push_back(it);
      }
  launch(np, np);
  // 5. the symmetric 2-channel matrices and their factors
  s.W.alloc((size_t)nchan * N * N);
  k_assemble_W<<<dim3(148, nchan), 256, 0, s.st>>>(nbf, np, s.pairof.p, s.O.p, s.W.p);
  CKT(cudaGetLastError());
  s.Bout.alloc((size_t)nchan * maxrank * N);
  s.sig.alloc((size_t)nchan * maxrank);
  s.rank.alloc(nchan);
  k_sign_cholesky<<<nchan, 256, (size_t)(N + maxrank) * sizeof(double), s.st>>>(N, maxrank, thresh, s.W.p, s.Bout.p, s.sig.p,
                                                                                s.rank.p);
  CKT(cudaGetLastError());
  s.hrank.resize(nchan);
  s.hsig.resize((size_t)nchan * maxrank);
  CKT(cudaMemcpyAsync(s.hrank.data(), s.rank.p, nchan * sizeof(int), cudaMemcpyDeviceToHost, s.st));
  CKT(cudaMemcpyAsync(s.hsig.data(), s.sig.p, s.hsig.size() * sizeof(double), cudaMemcpyDeviceToHost, s.st));
  CKT(cudaStreamSynchronize(s.st));
  for (int c = 0; c < nchan; c++) {
    const int r = s.hrank[c];
    if (r >= maxrank) throw std::runtime_error("device compute_tei: Cholesky rank exceeds the buffer");
    out[c].rank = r;
    out[c].sigma.assign(s.hsig.begin() + (size_t)c * maxrank, s.hsig.begin() + (size_t)c * maxrank + r);
    out[c].B.resize((size_t)r * N);
    CKT(cudaMemcpyAsync(out[c].B.data(), s.Bout.p + (size_t)c * maxrank * N, (size_t)r * N * sizeof(double),
                        cudaMemcpyDeviceToHost, s.st));
  }
  CKT(cudaStreamSynchronize(s.st));
}

}  // namespace hfq
