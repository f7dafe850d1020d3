// Finite-element primitives for the host-side basis setup: quadrature rules,
// element grids, Lagrange interpolating polynomial (LIP) shape functions and the
// order-doubling convergence loop.  Plain C++17, column-major dense matrices in
// std::vector<double>; no Eigen/Armadillo dependency.
//
// Behaviour follows the reference (susilehtola/HelFEM):
//   libhelfem/include/lobatto.h:36-109, chebyshev.h:32-57, grid.h:38-103,
//   LIPBasis.h:32-116, LIPBasis_eval.h:29-88,
//   libhelfem/src/FiniteElementBasis.cpp:52-64,297-313,441-479,500-604,
//   libhelfem/src/RadialBasis.cpp:114-162, src/diatomic/basis.cpp:125-200.
#pragma once
#include <cmath>
#include <cstddef>
#include <functional>
#include <limits>
#include <stdexcept>
#include <vector>

namespace hfq {

// Column-major dense matrix (rows x cols), element (i,j) at a[i + j*rows].
struct Mat {
  int rows = 0, cols = 0;
  std::vector<double> a;
  Mat() = default;
  Mat(int r, int c) : rows(r), cols(c), a((size_t)r * c, 0.0) {}
  double &operator()(int i, int j) { return a[(size_t)i + (size_t)j * rows]; }
  double operator()(int i, int j) const { return a[(size_t)i + (size_t)j * rows]; }
  bool empty() const { return a.empty(); }
};

double max_abs(const Mat &m);
double max_abs_diff(const Mat &a, const Mat &b);
// C = A^T diag(w) B for (npts x na), (npts x nb)
Mat weighted_gram(const Mat &A, const std::vector<double> &w, const Mat &B);

void lobatto_rule(int n, std::vector<double> &x, std::vector<double> &w);
void chebyshev_rule(int n, std::vector<double> &x, std::vector<double> &w);
std::vector<double> element_grid(double rmax, int num_el, int igrid, double zexp);

// d^n L_i / dx^n at points x for the Lagrange polynomials on nodes x0.
// Output (npts x nnodes), n in {0,1,2}.
Mat lip_eval(const std::vector<double> &x, const std::vector<double> &x0, int n);

// One-dimensional finite-element basis of LIPs with one shared function between
// neighbouring elements (noverlap = 1).
class FEBasis {
 public:
  FEBasis() = default;
  FEBasis(int nnodes, const std::vector<double> &bval, bool zero_func_left, bool zero_func_right);
  int nel() const { return (int)bval_.size() - 1; }
  int nbf() const { return nbf_; }
  int nnodes() const { return (int)x0_.size(); }
  const std::vector<double> &nodes() const { return x0_; }
  std::vector<int> enabled(int iel) const;
  int nprim(int iel) const { return (int)enabled(iel).size(); }
  int first(int iel) const { return first_[iel]; }
  int last(int iel) const { return last_[iel]; }
  double begin(int iel) const { return bval_[iel]; }
  double end(int iel) const { return bval_[iel + 1]; }
  double mid(int iel) const { return 0.5 * (bval_[iel + 1] + bval_[iel]); }
  double scale(int iel) const { return 0.5 * (bval_[iel + 1] - bval_[iel]); }
  std::vector<double> coord(const std::vector<double> &x, int iel) const;
  // n-th derivative w.r.t. the real coordinate of the enabled functions.
  Mat eval_dnf(const std::vector<double> &x, int n, int iel) const;
  // B(r)/r (and derivatives) on the element touching r = 0.
  Mat eval_over_r(const std::vector<double> &x, int n, int iel) const;

 private:
  std::vector<double> x0_, bval_;
  std::vector<int> first_, last_;
  bool zl_ = false, zr_ = false;
  int nbf_ = 0;
};

// Order-doubling refinement of probe(n) until the block is stable.
//  floor_rel < 0 : three exits (8 eps, 2x stall, two-doubling stall).
//  floor_rel >= 0: two exits (8 eps, sqrt-eps regime with floor or 2x stall),
//                  optional fall-back to the seed order at the cap.
Mat converge_block(const std::function<Mat(int)> &probe, int nstart, int nmax, double floor_rel = -1.0,
                   bool seed_fallback = false, int *nconv = nullptr);

}  // namespace hfq
