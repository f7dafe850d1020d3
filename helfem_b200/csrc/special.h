// Special functions for the host-side setup: associated Legendre functions
// P_l^m(x), Q_l^m(x) for x >= 1 (prolate spheroidal radial coordinate) and
// Gaunt coefficients.
//
// Conventions follow the reference so that tables are interchangeable:
//   Legendre: Hobson form, no Condon-Shortley phase outside the cut
//             (src/legendre/Legendre.h:16-22); three regimes for Q
//             (upward / Christoffel / Miller, Legendre.h:335-378).
//   Gaunt:    coeff(L,M,l,m,lp) = int Y_L^M* Y_l^m Y_lp^(M-m) dOmega
//             (src/general/gaunt.cpp:77-84,202-252); the reference takes the
//             values from the external library wignernj -- here they are
//             computed by exact-degree Gauss-Legendre quadrature of normalised
//             associated Legendre functions in extended precision.
#pragma once
#include <vector>

namespace hfq {

// P_l^m(x) for l = 0..lmax at fixed m (x >= 1); out[l].
void legendre_p(int lmax, int m, double x, double *out);
// Q_l^m(x) for l = 0..lmax at fixed m (x > 1); out[l].
void legendre_q(int lmax, int m, double x, double *out);

class GauntTable {
 public:
  // lmax: largest l on any of the three positions that will be queried.
  explicit GauntTable(int lmax);
  // int Y_L^M* Y_l^m Y_lp^(M-m)
  double coeff(int L, int M, int l, int m, int lp) const;
  // int Y_lj^mj* cos^2(theta) Y_L^M Y_li^mi  (src/general/gaunt.cpp:254-272)
  double mod_coeff(int lj, int mj, int L, int M, int li, int mi) const;

 private:
  int lmax_, nq_;
  std::vector<long double> xq_, wq_;
  // theta_[ (l*(l+1)/2 + m) * nq + q ] = normalised Theta_l^m(x_q), m >= 0
  std::vector<long double> theta_;
  long double theta(int l, int m, int q) const;
  // int Y_l1^m1 Y_l2^m2 Y_l3^m3 (no conjugation)
  double gaunt3(int l1, int m1, int l2, int m2, int l3, int m3) const;
};

}  // namespace hfq
