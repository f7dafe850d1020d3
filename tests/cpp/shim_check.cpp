// Compile/link/run check of include/helfem_b200.hpp against libhelfemqc_b200.so with a minimal
// column-major matrix type (stand-in for Eigen::MatrixXd / arma::mat).  Runs the host-only part of the
// interface: basis construction, compute_tei(), one-electron matrices, error behaviour.  With a CUDA
// device present (argv[1] == "gpu") it also runs one J/K build and checks K = K^T and linearity.
#include <cmath>
#include <cstdio>
#include <cstring>
#include <vector>

#include "helfem_b200.hpp"

struct Mat {
  long r = 0, c = 0;
  std::vector<double> a;
  Mat() = default;
  Mat(long rows, long cols) : r(rows), c(cols), a((size_t)rows * cols, 0.0) {}
  double *data() { return a.data(); }
  const double *data() const { return a.data(); }
  long rows() const { return r; }
  long cols() const { return c; }
  double &operator()(long i, long j) { return a[(size_t)i + (size_t)j * r]; }
  double operator()(long i, long j) const { return a[(size_t)i + (size_t)j * r]; }
};

int main(int argc, char **argv) {
  const bool gpu = argc > 1 && !std::strcmp(argv[1], "gpu");
  using Basis = helfem_b200::atomic::TwoDBasis<Mat>;
  Basis basis(/*Z=*/2, /*lmax=*/1, /*mmax=*/1, /*nelem=*/2);
  // reference behaviour: using the basis before compute_tei() is a logic_error (TwoDBasis.cpp:775)
  try {
    Mat P(4, 4);
    basis.coulomb(P);
    std::printf("FAIL: no exception before compute_tei\n");
    return 1;
  } catch (const std::logic_error &) {
  }
  basis.compute_tei();
  const long n = (long)basis.Nbf();
  Mat S, T, V;
  basis.one_electron(S, T, V);
  double asym = 0.0, tr = 0.0;
  for (long i = 0; i < n; i++) {
    tr += S(i, i);
    for (long j = 0; j < n; j++) asym = std::fmax(asym, std::fabs(S(i, j) - S(j, i)));
  }
  if (!(tr > 0.0) || asym > 1e-14) {
    std::printf("FAIL: overlap tr %g asym %g\n", tr, asym);
    return 1;
  }
  // wrong-sized matrix: logic_error like remove_boundaries (basis.cpp:2093-2097)
  try {
    Mat P(n + 1, n + 1);
    basis.exchange(P);
    std::printf("FAIL: no exception for a wrong-sized matrix\n");
    return 1;
  } catch (const std::logic_error &) {
  }
  helfem_b200::diatomic::TwoDBasis<Mat> dia(1, 1, 1.4, {2, 1}, 2);
  dia.set_absm_symmetric(true);   // before compute_tei: remembered, applied when the context exists
  if (!gpu) {
    // no CUDA device: the first J/K call must fail loudly (no CPU fallback), as a runtime_error
    try {
      Mat P(n, n);
      basis.coulomb(P);
      std::printf("FAIL: coulomb succeeded without a GPU\n");
      return 1;
    } catch (const std::logic_error &) {
      std::printf("FAIL: missing device reported as logic_error\n");
      return 1;
    } catch (const std::runtime_error &) {
    }
    std::printf("OK host-only Nbf=%ld\n", n);
    return 0;
  }
  Mat P(n, n);
  for (long i = 0; i < n; i++)
    for (long j = 0; j < n; j++) P(i, j) = 1.0 / (1.0 + i + j);
  const Mat J = basis.coulomb(P), K = basis.exchange(P);
  Mat J2, K2;
  basis.coulomb_exchange(P, 0.5, J2, K2);
  double e = 0.0, kmax = 0.0;
  for (long i = 0; i < n; i++)
    for (long j = 0; j < n; j++) {
      kmax = std::fmax(kmax, std::fabs(K(i, j)));
      e = std::fmax(e, std::fabs(K(i, j) - K(j, i)));
      e = std::fmax(e, std::fabs(0.5 * K(i, j) - K2(i, j)));
      e = std::fmax(e, std::fabs(J(i, j) - J2(i, j)));
    }
  if (!(kmax > 0.0) || e > 1e-12 * kmax) {
    std::printf("FAIL: K symmetry / linearity %g of %g\n", e, kmax);
    return 1;
  }
  std::printf("OK gpu Nbf=%ld\n", n);
  return 0;
}
