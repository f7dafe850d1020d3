// Density functionals evaluated on the device: the step of DFTGridWorkerBase::compute_xc
// (src/general/dftgrid_common.cpp:98-255) that the reference hands to libxc, for the functionals its recorded
// test energies use: XC_LDA_X (1), XC_LDA_C_VWN (7, VWN5), XC_GGA_X_PBE (101), XC_GGA_C_PBE (130), XC_MGGA_X_TPSS (202),
// XC_MGGA_C_TPSS (231).  Each functional is written once as the energy per particle e(n, sigma, tau) of the
// spin-unpolarised density in forward-mode dual numbers (value, d/dn, d/dsigma, d/dtau), so exc, vrho = d(n e)/dn,
// vsigma = d(n e)/dsigma and vtau = d(n e)/dtau come from one expression and cannot drift apart.  Constants are libxc's (lda_c_vwn.c paramagnetic VWN5 set; gga_x_pbe.c kappa = 0.8040,
// mu = beta pi^2 / 3; gga_c_pbe.c beta = 0.06672455060314922, gamma = (1 - ln 2) / pi^2 on the "modified" PW92 of
// lda_c_pw.c; mgga_x_tpss.c b = 0.40, c = 1.59096, e = 1.537, kappa = 0.804, mu = 0.21951; mgga_c_tpss.c d = 2.8,
// C(0, 0) = 0.53, z = min(tau_W / tau, 1)).  Spin-polarised densities: exchange through the exact spin-scaling relation
// E_x[na, nb] = (E_x[2 na] + E_x[2 nb]) / 2; polarised correlation is not built in (callers use hfq_grid_density +
// libxc + hfq_grid_fxc for it).
#pragma once
#ifdef __CUDACC__
#include <cuda_runtime.h>
#else   // host-only compilation (tests/cpp/xc_check.cpp): the same expressions under a plain C++ compiler
#include <cmath>
#ifndef __host__
#define __host__
#define __device__
#endif
using std::atan; using std::cbrt; using std::exp; using std::expm1; using std::log; using std::log1p; using std::sqrt;
#endif

namespace hfq {
namespace xc {

struct D2 {
  double v, n, s, t;   // value, d/dn, d/dsigma, d/dtau
};
__host__ __device__ inline D2 mk(double v, double n = 0.0, double s = 0.0, double t = 0.0) { return D2{v, n, s, t}; }
__host__ __device__ inline D2 operator+(D2 a, D2 b) { return mk(a.v + b.v, a.n + b.n, a.s + b.s, a.t + b.t); }
__host__ __device__ inline D2 operator-(D2 a, D2 b) { return mk(a.v - b.v, a.n - b.n, a.s - b.s, a.t - b.t); }
__host__ __device__ inline D2 operator*(D2 a, D2 b) {
  return mk(a.v * b.v, a.n * b.v + a.v * b.n, a.s * b.v + a.v * b.s, a.t * b.v + a.v * b.t);
}
__host__ __device__ inline D2 operator/(D2 a, D2 b) {
  const double q = a.v / b.v;
  return mk(q, (a.n - q * b.n) / b.v, (a.s - q * b.s) / b.v, (a.t - q * b.t) / b.v);
}
__host__ __device__ inline D2 operator+(D2 a, double c) { return mk(a.v + c, a.n, a.s, a.t); }
__host__ __device__ inline D2 operator+(double c, D2 a) { return mk(a.v + c, a.n, a.s, a.t); }
__host__ __device__ inline D2 operator-(D2 a, double c) { return mk(a.v - c, a.n, a.s, a.t); }
__host__ __device__ inline D2 operator-(double c, D2 a) { return mk(c - a.v, -a.n, -a.s, -a.t); }
__host__ __device__ inline D2 operator*(D2 a, double c) { return mk(a.v * c, a.n * c, a.s * c, a.t * c); }
__host__ __device__ inline D2 operator*(double c, D2 a) { return mk(a.v * c, a.n * c, a.s * c, a.t * c); }
__host__ __device__ inline D2 operator/(D2 a, double c) { return mk(a.v / c, a.n / c, a.s / c, a.t / c); }
__host__ __device__ inline D2 operator/(double c, D2 a) { return mk(c) / a; }
__host__ __device__ inline D2 chain(D2 a, double f, double df) { return mk(f, df * a.n, df * a.s, df * a.t); }
__host__ __device__ inline D2 dlog(D2 a) { return chain(a, log(a.v), 1.0 / a.v); }
__host__ __device__ inline D2 dlog1p(D2 a) { return chain(a, log1p(a.v), 1.0 / (1.0 + a.v)); }
__host__ __device__ inline D2 dexpm1(D2 a) { return chain(a, expm1(a.v), exp(a.v)); }
__host__ __device__ inline D2 dsqrt(D2 a) {
  const double r = sqrt(a.v);
  return chain(a, r, 0.5 / r);
}
__host__ __device__ inline D2 dcbrt(D2 a) {
  const double r = cbrt(a.v);
  return chain(a, r, r / (3.0 * a.v));
}
__host__ __device__ inline D2 datan(D2 a) { return chain(a, atan(a.v), 1.0 / (1.0 + a.v * a.v)); }
__host__ __device__ inline D2 dmax(D2 a, D2 b) { return a.v >= b.v ? a : b; }
__host__ __device__ inline D2 dmin1(D2 a) { return a.v < 1.0 ? a : mk(1.0); }   // min(a, 1)

constexpr double kPi = 3.14159265358979323846;

// Wigner-Seitz radius rs = (3 / (4 pi n))^(1/3)
__host__ __device__ inline D2 wigner_seitz(D2 n) { return dcbrt(3.0 / (4.0 * kPi) / n); }
// exchange energy per particle of the uniform gas
__host__ __device__ inline D2 ex_uniform(D2 n) { return (-0.75 * cbrt(3.0 / kPi)) * dcbrt(n); }

__host__ __device__ inline D2 lda_x(D2 n, D2) { return ex_uniform(n); }

__host__ __device__ inline D2 lda_c_vwn(D2 n, D2) {
  const double A = 0.0310907, b = 3.72744, c = 12.9352, x0 = -0.10498;
  const double Q = sqrt(4.0 * c - b * b), X0 = x0 * x0 + b * x0 + c;
  const D2 x = dsqrt(wigner_seitz(n));
  const D2 X = x * x + b * x + c;
  const D2 at = datan(Q / (2.0 * x + b));
  const D2 xm = x - x0;
  return A * (dlog(x * x / X) + (2.0 * b / Q) * at -
              (b * x0 / X0) * (dlog(xm * xm / X) + (2.0 * (b + 2.0 * x0) / Q) * at));
}

__host__ __device__ inline D2 gga_x_pbe(D2 n, D2 sigma) {
  const double kappa = 0.8040, mu = 0.06672455060314922 * kPi * kPi / 3.0;
  const D2 kF = dcbrt(3.0 * kPi * kPi * n);
  const D2 d = 2.0 * kF * n;
  const D2 s2 = sigma / (d * d);
  return ex_uniform(n) * (1.0 + kappa - kappa / (1.0 + (mu / kappa) * s2));
}

// PW92 correlation energy per particle ("modified" constants of lda_c_pw.c), paramagnetic or fully polarised branch
__host__ __device__ inline D2 pw92(D2 rs, bool polarised) {
  const double a = polarised ? 0.01554535 : 0.0310907, a1 = polarised ? 0.20548 : 0.21370;
  const double b1 = polarised ? 14.1189 : 7.5957, b2 = polarised ? 6.1977 : 3.5876;
  const double b3 = polarised ? 3.3662 : 1.6382, b4 = polarised ? 0.62517 : 0.49294;
  const D2 srs = dsqrt(rs);
  return (-2.0 * a) * (1.0 + a1 * rs) * dlog1p(1.0 / ((2.0 * a) * (b1 * srs + b2 * rs + b3 * rs * srs + b4 * rs * rs)));
}

// PBE correlation energy per particle of a density n with |grad n|^2 = sigma at zeta = 0, or at zeta = 1
// (phi = 2^(-1/3): the one-spin term of the TPSS correlation)
__host__ __device__ inline D2 pbe_correlation(D2 n, D2 sigma, bool polarised) {
  const double beta = 0.06672455060314922, gamma = (1.0 - 0.69314718055994530942) / (kPi * kPi);
  const double phi = polarised ? 0.79370052598409973738 : 1.0, gphi3 = gamma * phi * phi * phi;
  const D2 ec = pw92(wigner_seitz(n), polarised);
  const D2 kF = dcbrt(3.0 * kPi * kPi * n);
  const D2 d = (2.0 * phi) * dsqrt((4.0 / kPi) * kF) * n;     // 2 phi ks n
  const D2 t2 = sigma / (d * d);
  const D2 Aa = (beta / gamma) / dexpm1((-1.0 / gphi3) * ec);
  const D2 At2 = Aa * t2;
  return ec + gphi3 * dlog1p((beta / gamma) * t2 * (1.0 + At2) / (1.0 + At2 + At2 * At2));
}

__host__ __device__ inline D2 gga_c_pbe(D2 n, D2 sigma) { return pbe_correlation(n, sigma, false); }

// TPSS exchange (Tao, Perdew, Staroverov, Scuseria, PRL 91, 146401 (2003), eqs. 5-10)
__host__ __device__ inline D2 mgga_x_tpss(D2 n, D2 sigma, D2 tau) {
  const double kappa = 0.804, b = 0.40, c = 1.59096, e = 1.537, mu = 0.21951, mu_ge = 10.0 / 81.0, se = sqrt(e);
  const D2 n13 = dcbrt(n), n53 = n * n13 * n13, n83 = n53 * n;
  const double c23 = cbrt(3.0 * kPi * kPi) * cbrt(3.0 * kPi * kPi);   // (3 pi^2)^(2/3)
  const D2 p = sigma / ((4.0 * c23) * n83);
  const D2 tw = sigma / (8.0 * n);
  const D2 z = tw / tau;
  const D2 alpha = (tau - tw) / ((0.3 * c23) * n53);
  const D2 qb = (9.0 / 20.0) * (alpha - 1.0) / dsqrt(1.0 + b * alpha * (alpha - 1.0)) + (2.0 / 3.0) * p;
  const D2 z2 = z * z, opz2 = 1.0 + z2;
  const D2 den = 1.0 + se * p;
  const D2 x = ((mu_ge + c * z2 / (opz2 * opz2)) * p + (146.0 / 2025.0) * qb * qb -
                (73.0 / 405.0) * qb * dsqrt(0.5 * ((9.0 / 25.0) * z2 + p * p)) + (mu_ge * mu_ge / kappa) * p * p +
                (2.0 * se * mu_ge * 9.0 / 25.0) * z2 + (e * mu) * p * p * p) /
               (den * den);
  return ex_uniform(n) * (1.0 + kappa - kappa / (1.0 + x / kappa));
}

// TPSS correlation (ibid., eqs. 11-14), zeta = 0
__host__ __device__ inline D2 mgga_c_tpss(D2 n, D2 sigma, D2 tau) {
  const double d = 2.8, C = 0.53;
  const D2 z = dmin1(sigma / (8.0 * n * tau));
  const D2 e_pbe = pbe_correlation(n, sigma, false);
  const D2 e_one = dmax(pbe_correlation(0.5 * n, 0.25 * sigma, true), e_pbe);
  const D2 z2 = z * z;
  const D2 rev = e_pbe * (1.0 + C * z2) - (1.0 + C) * z2 * e_one;
  return rev * (1.0 + d * rev * z2 * z);
}

__host__ __device__ inline bool known(int id) { return id == 1 || id == 7 || id == 101 || id == 130 || id == 202 || id == 231; }
__host__ __device__ inline bool is_gga(int id) { return id == 101 || id == 130 || id == 202 || id == 231; }   // needs the gradient
__host__ __device__ inline bool is_mgga(int id) { return id == 202 || id == 231; }                              // needs tau
__host__ __device__ inline bool is_exchange(int id) { return id == 1 || id == 101 || id == 202; }

// energy per particle of functional id at (n, sigma, tau) with derivative seeds d/dn = 1, d/dsigma = 1, d/dtau = 1
__host__ __device__ inline D2 energy(int id, double n, double sigma, double tau = 1.0) {
  const D2 dn = mk(n, 1.0, 0.0, 0.0), ds = mk(sigma, 0.0, 1.0, 0.0), dt = mk(tau, 0.0, 0.0, 1.0);
  switch (id) {
    case 1: return lda_x(dn, ds);
    case 7: return lda_c_vwn(dn, ds);
    case 101: return gga_x_pbe(dn, ds);
    case 130: return gga_c_pbe(dn, ds);
    case 202: return mgga_x_tpss(dn, ds, dt);
    case 231: return mgga_c_tpss(dn, ds, dt);
  }
  return mk(0.0);
}

}  // namespace xc
}  // namespace hfq
