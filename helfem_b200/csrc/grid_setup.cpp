// Host tables of the atomic DFT grid: radial functions (B/r and its derivatives) at the
// modified Gauss-Chebyshev nodes of every element, the theta x phi compound rule and the
// spherical harmonics with their theta derivatives.
// Reference behaviour: libhelfem/src/RadialBasis.cpp:868-926 (get_bf/get_df/get_lf),
// src/general/angular.cpp:21-69, src/general/spherical_harmonics.cpp:20-35,
// src/atomic/TwoDBasis.cpp:1133-1235 (eval_bf/eval_df/eval_lf).
#include <cmath>
#include <stdexcept>

#include "fem.h"
#include "grid.h"

namespace hfq {

static std::complex<double> ylm(int l, int m, double cth, double phi) {
  if (m < 0) {
    const std::complex<double> v = std::conj(ylm(l, -m, cth, phi));
    return ((-m) & 1) ? -v : v;
  }
  if (m > l) return 0.0;
  // std::sph_legendre: normalised associated Legendre function incl. the Condon-Shortley phase
  return std::polar(1.0, m * phi) * std::sph_legendre((unsigned)l, (unsigned)m, std::acos(cth));
}

// radial tables of the atomic / sadatom basis at the Chebyshev nodes (RadialBasis.cpp:868-926): B/r, its first
// derivative, radial Laplacian f'' + 2 f'/r, and f/r^2
static void fill_atomic_radial(const BasisTables &t, GridTables &g) {
  // radial functions
  const FEBasis fe(t.nnodes, t.bval, true, true);
  std::vector<double> xq, wq;
  chebyshev_rule(g.nrad, xq, wq);
  const size_t tsz = (size_t)g.Nel * g.NI * g.nrad;
  g.F.assign(tsz, 0.0);
  g.D.assign(tsz, 0.0);
  g.L1.assign(tsz, 0.0);
  g.F2.assign(tsz, 0.0);
  g.F3.assign(tsz, 0.0);
  for (int e = 0; e < g.Nel; e++) {
    const std::vector<double> r = fe.coord(xq, e);
    for (int q = 0; q < g.nrad; q++) {
      g.r.push_back(r[q]);
      g.wrad.push_back(wq[q] * fe.scale(e));
    }
    Mat f, d, l2;
    if (e == 0) {
      f = fe.eval_over_r(xq, 0, e);
      d = fe.eval_over_r(xq, 1, e);
      l2 = fe.eval_over_r(xq, 2, e);
    } else {
      const Mat B0 = fe.eval_dnf(xq, 0, e), B1 = fe.eval_dnf(xq, 1, e), B2 = fe.eval_dnf(xq, 2, e);
      f = B0;
      d = B0;
      l2 = B0;
      for (int j = 0; j < B0.cols; j++)
        for (int q = 0; q < g.nrad; q++) {
          const double ir = 1.0 / r[q];
          f(q, j) = B0(q, j) * ir;
          d(q, j) = (-B0(q, j) * ir + B1(q, j)) * ir;
          l2(q, j) = ((2.0 * B0(q, j) * ir - 2.0 * B1(q, j)) * ir + B2(q, j)) * ir;
        }
    }
    for (int j = 0; j < f.cols; j++)
      for (int q = 0; q < g.nrad; q++) {
        const size_t o = ((size_t)e * g.NI + j) * g.nrad + q;
        g.F[o] = f(q, j);
        g.D[o] = d(q, j);
        g.L1[o] = l2(q, j) + 2.0 * d(q, j) / r[q];
        g.F2[o] = f(q, j) / (r[q] * r[q]);
      }
  }
}

GridTables build_atomic_grid(const BasisTables &t, int lang, int mang) {
  if (t.kind != BasisKind::Atomic) throw std::logic_error("build_atomic_grid: atomic basis required");
  if (t.bval.empty()) throw std::logic_error("build_atomic_grid: basis was not built by this library");
  GridTables g;
  g.lang = lang;
  g.mang = mang;
  g.nang = lang * mang;
  g.nrad = t.nquad;
  g.Nel = t.Nel;
  g.Nang = t.Nang();
  g.NI = 0;
  for (int n : t.en) g.NI = std::max(g.NI, n);
  // angular rule: Chebyshev in cos(theta) x uniform phi
  std::vector<double> xl, wl;
  chebyshev_rule(lang, xl, wl);
  const double dphi = 2.0 * std::acos(-1.0) / mang;
  for (int i = 0; i < lang; i++)
    for (int j = 0; j < mang; j++) {
      g.cth.push_back(xl[i]);
      g.phi.push_back(j * dphi);
      g.wang.push_back(wl[i] * dphi);
    }
  // angular functions
  g.Y.assign((size_t)g.Nang * g.nang, 0.0);
  g.Th.assign((size_t)g.Nang * g.nang, 0.0);
  for (int a = 0; a < g.Nang; a++) {
    const int l = t.lval[a], m = t.mval[a];
    for (int ia = 0; ia < g.nang; ia++) {
      const double c = g.cth[ia], p = g.phi[ia];
      const double sinth = std::sqrt(std::max((1.0 - c) * (1.0 + c), 0.0));
      const double cot = sinth > 0.0 ? c / sinth : 0.0;
      const std::complex<double> y = ylm(l, m, c, p);
      std::complex<double> ang = (double)m * cot * y;
      if (m < l) ang += std::sqrt((double)(l - m) * (l + m + 1)) * std::polar(1.0, -p) * ylm(l, m + 1, c, p);
      g.Y[(size_t)a * g.nang + ia] = y;
      g.Th[(size_t)a * g.nang + ia] = ang;
    }
  }
  fill_atomic_radial(t, g);
  // per-point weights and scale factors (1, r, r sin(theta)); src/atomic/dftgrid.cpp:486-512
  const size_t N = (size_t)g.Nel * g.nang * g.nrad;
  g.wtot.assign(N, 0.0);
  g.lfac.assign(N, 1.0);
  for (auto &sc : g.scale) sc.assign(N, 1.0);
  for (int e = 0; e < g.Nel; e++)
    for (int ia = 0; ia < g.nang; ia++) {
      const double sth = std::sqrt(1.0 - g.cth[ia] * g.cth[ia]);
      for (int q = 0; q < g.nrad; q++) {
        const size_t p = ((size_t)e * g.nang + ia) * g.nrad + q;
        const double r = g.r[(size_t)e * g.nrad + q];
        g.wtot[p] = g.wang[ia] * g.wrad[(size_t)e * g.nrad + q] * r * r;
        g.scale[1][p] = r;
        g.scale[2][p] = r * sth;
      }
    }
  return g;
}

// Pure-m diatomic grid: real Y_l^m(cos nu) at phi = 0, Gauss-Chebyshev in cos(nu), the FE
// functions B, B', B'' at the mu quadrature nodes; weight 2 pi w_nu w_mu Rh^3 sinh(mu)
// (sinh^2 mu + sin^2 nu); scale factors h_mu = h_nu = Rh sqrt(sinh^2 mu + sin^2 nu),
// h_phi = Rh sinh(mu) sin(nu); Laplacian = (1/h^2) [R'' + coth(mu) R' - (l(l+1) + m^2/sinh^2 mu) R] Y.
// Reference: src/diatomic/dftgrid_purem.cpp:30-200.
// mang <= 1: pure-m grid (phi analytic, real harmonics at phi = 0, weight 2 pi w_nu);
// mang >= 2: the general 3D grid of src/diatomic/dftgrid.cpp:414-518 (complex harmonics on the
// theta x phi compound rule, every (a,b) pair couples, no Laplacian in the reference).
GridTables build_sadatom_grid(const BasisTables &t) {
  if (t.kind != BasisKind::Sadatom) throw std::logic_error("build_sadatom_grid: sadatom basis required");
  if (t.bval.empty()) throw std::logic_error("build_sadatom_grid: basis was not built by this library");
  GridTables g;
  // a batch of atoms (BasisTables::batch) is laid out along the "angular point" axis: point (element, atom, radial
  // node), and the angular table of function a = (atom, l) is the indicator of its atom -- the separable engine then
  // produces per-atom densities and per-atom Fock blocks in one pass
  const int nb = std::max(1, t.batch), nl = t.Nang() / nb;
  g.lang = g.mang = 1;
  g.nang = nb;
  g.nrad = t.nquad;
  g.Nel = t.Nel;
  g.Nang = t.Nang();
  g.NI = 0;
  for (int n : t.en) g.NI = std::max(g.NI, n);
  g.pure_m = true;           // both kinetic potentials enter the same term (src/sadatom/dftgrid.cpp:303, :406)
  g.same_l_only = true;
  g.clamp_theta_kin = true;
  const double pi = std::acos(-1.0);
  g.cth.assign(nb, 1.0);
  g.phi.assign(nb, 0.0);
  g.wang.assign(nb, 4.0 * pi);
  g.Y.assign((size_t)g.Nang * nb, 0.0);
  g.Th.assign((size_t)g.Nang * nb, 0.0);
  for (int a = 0; a < g.Nang; a++) {
    const int atom = a / nl;
    g.Y[(size_t)a * nb + atom] = 1.0;
    g.Th[(size_t)a * nb + atom] = std::complex<double>(0.0, std::sqrt((double)t.lval[a] * (t.lval[a] + 1)));
  }
  fill_atomic_radial(t, g);
  const size_t N = (size_t)g.Nel * nb * g.nrad;
  g.wtot.resize(N);
  for (int c = 0; c < 3; c++) g.scale[c].resize(N);
  g.lfac.assign(N, 1.0);
  for (int e = 0; e < g.Nel; e++)
    for (int ia = 0; ia < nb; ia++)
      for (int q = 0; q < g.nrad; q++) {
        const size_t p = ((size_t)e * nb + ia) * g.nrad + q, pr = (size_t)e * g.nrad + q;
        g.wtot[p] = 4.0 * pi * g.wrad[pr] * g.r[pr] * g.r[pr];
        g.scale[0][p] = 1.0;
        g.scale[1][p] = g.r[pr];
        g.scale[2][p] = g.r[pr];
      }
  return g;
}

GridTables build_diatomic_grid(const BasisTables &t, int lang, int mang) {
  if (t.kind != BasisKind::Diatomic) throw std::logic_error("build_diatomic_grid: diatomic basis required");
  if (t.bval.empty()) throw std::logic_error("build_diatomic_grid: basis was not built by this library");
  GridTables g;
  g.pure_m = mang <= 1;
  g.lang = lang;
  g.mang = g.pure_m ? 1 : mang;
  g.nang = lang * g.mang;
  g.nrad = t.nquad;
  g.Nel = t.Nel;
  g.Nang = t.Nang();
  for (int n : t.en) g.NI = std::max(g.NI, n);
  {
    std::vector<double> xl, wl;
    chebyshev_rule(lang, xl, wl);
    const double dphi = 2.0 * std::acos(-1.0) / g.mang;
    for (int i = 0; i < lang; i++)
      for (int j = 0; j < g.mang; j++) {
        g.cth.push_back(xl[i]);
        g.phi.push_back(g.pure_m ? 0.0 : j * dphi);
        g.wang.push_back(g.pure_m ? 2.0 * std::acos(-1.0) * wl[i] : wl[i] * dphi);
      }
  }
  g.Y.assign((size_t)g.Nang * g.nang, 0.0);
  g.Th.assign((size_t)g.Nang * g.nang, 0.0);
  for (int a = 0; a < g.Nang; a++) {
    const int l = t.lval[a], m = t.mval[a];
    for (int ia = 0; ia < g.nang; ia++) {
      const double c = g.cth[ia], ph = g.phi[ia];
      const double sth = std::sqrt(std::max((1.0 - c) * (1.0 + c), 0.0));
      const double cot = sth > 0.0 ? c / sth : 0.0;
      if (g.pure_m) {
        const double y = ylm(l, m, c, 0.0).real();
        double dy = m * cot * y;
        if (m < l) dy += std::sqrt((double)(l - m) * (double)(l + m + 1)) * ylm(l, m + 1, c, 0.0).real();
        g.Y[(size_t)a * g.nang + ia] = y;
        g.Th[(size_t)a * g.nang + ia] = dy;
      } else {
        const std::complex<double> y = ylm(l, m, c, ph);
        std::complex<double> ang = (double)m * cot * y;
        if (m < l) ang += std::sqrt((double)(l - m) * (l + m + 1)) * std::polar(1.0, -ph) * ylm(l, m + 1, c, ph);
        g.Y[(size_t)a * g.nang + ia] = y;
        g.Th[(size_t)a * g.nang + ia] = ang;
      }
    }
  }
  const FEBasis fe(t.nnodes, t.bval, false, true);
  std::vector<double> xq, wq;
  chebyshev_rule(g.nrad, xq, wq);
  const size_t tsz = (size_t)g.Nel * g.NI * g.nrad;
  g.F.assign(tsz, 0.0);
  g.D.assign(tsz, 0.0);
  g.L1.assign(tsz, 0.0);
  g.F2.assign(tsz, 0.0);
  g.F3.assign(tsz, 0.0);
  for (int e = 0; e < g.Nel; e++) {
    const std::vector<double> mu = fe.coord(xq, e);
    const Mat B0 = fe.eval_dnf(xq, 0, e), B1 = fe.eval_dnf(xq, 1, e), B2 = fe.eval_dnf(xq, 2, e);
    for (int q = 0; q < g.nrad; q++) {
      g.r.push_back(mu[q]);
      g.wrad.push_back(wq[q] * fe.scale(e));
    }
    for (int j = 0; j < B0.cols; j++)
      for (int q = 0; q < g.nrad; q++) {
        const double sh = std::sinh(mu[q]), coth = sh > 0.0 ? std::cosh(mu[q]) / sh : 0.0;
        const size_t o = ((size_t)e * g.NI + j) * g.nrad + q;
        g.F[o] = B0(q, j);
        g.D[o] = B1(q, j);
        g.L1[o] = B2(q, j) + coth * B1(q, j);
        g.F2[o] = B0(q, j);
        g.F3[o] = sh > 0.0 ? B0(q, j) / (sh * sh) : 0.0;
      }
  }
  const size_t N = (size_t)g.Nel * g.nang * g.nrad;
  g.wtot.assign(N, 0.0);
  g.lfac.assign(N, 0.0);
  for (auto &sc : g.scale) sc.assign(N, 1.0);
  const double Rh = t.Rhalf;
  for (int e = 0; e < g.Nel; e++)
    for (int ia = 0; ia < g.nang; ia++) {
      const double c = g.cth[ia];
      const double sth = std::sqrt(std::max((1.0 - c) * (1.0 + c), 0.0));
      for (int q = 0; q < g.nrad; q++) {
        const size_t p = ((size_t)e * g.nang + ia) * g.nrad + q;
        const double sh = std::sinh(g.r[(size_t)e * g.nrad + q]);
        const double h = Rh * std::sqrt(sh * sh + sth * sth), hphi = Rh * sh * sth;
        g.wtot[p] = g.wang[ia] * g.wrad[(size_t)e * g.nrad + q] * Rh * Rh * Rh * sh * (sh * sh + sth * sth);
        g.scale[0][p] = h;
        g.scale[1][p] = h;
        g.scale[2][p] = hphi;
        g.lfac[p] = h > 0.0 ? 1.0 / (h * h) : 0.0;
      }
    }
  return g;
}

}  // namespace hfq
