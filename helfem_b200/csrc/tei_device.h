// Device-side compute_tei of the diatomic basis: the in-element two-electron kernels
// W = [[T00, -T02], [-T02^T, T22]] of every (L, |M|) channel by nested Gauss-Chebyshev quadrature
// (src/diatomic/quadrature.cpp:188-257, src/diatomic/basis.cpp:1382-1483) and their sign-aware pivoted Cholesky
// factors (basis.cpp:1483-1537), for one radial element and one |M| run per call.
#pragma once
#include <vector>

namespace hfq {

// Host-prepared data of the nested rule of one element at the converged order n (plain arrays; the device never
// evaluates the basis functions themselves).
struct TeiRuleView {
  int n = 0, nbf = 0, npair = 0;
  const double *cw = nullptr;      // [n*n]  sublen[ip] * w[q] * sinh(mu_sub(ip, q))
  const double *subch = nullptr;   // [n*n]  cosh(mu_sub(ip, q))
  const double *subbf = nullptr;   // [ip][k][q]: basis function k at sub-interval point (ip, q)
  const double *uw = nullptr;      // [n]    mulen * w[q] * sinh(mu[q])
  const double *chmu = nullptr;    // [n]    cosh(mu[q])
  const double *bfprod = nullptr;  // [p][q] B_i B_j at the outer points, pair p
  const int *pi = nullptr, *pj = nullptr;   // [npair] pair -> (i <= j)
};

struct TeiChannelResult {
  std::vector<double> B, sigma;   // B[(size_t)p * N + i], N = 2 nbf^2 (column p of the factor)
  int rank = 0;
};

class TeiDevice {
 public:
  explicit TeiDevice(int device);
  ~TeiDevice();
  // element data (uploaded once per element)
  void set_rule(const TeiRuleView &r);
  // one |M| run: Legendre P of degree L <= Lhi at the sub-interval points on the device, Q at the outer points from the
  // host (Qout[L * n + q]); channels c = 0 .. nchan-1 with multipole order Lvals[c].  thresh: relative Cholesky threshold.
  void run(int Mabs, int Lhi, const double *Qout, const int *Lvals, int nchan, double thresh, TeiChannelResult *out);

 private:
  struct Impl;
  Impl *p_;
};

}  // namespace hfq
