#include "special.h"

#include <algorithm>
#include <cmath>
#include <limits>
#include <stdexcept>

namespace hfq {

// ---------------------------------------------------------------------------
// Legendre functions, x >= 1
// ---------------------------------------------------------------------------

void legendre_p(int lmax, int m, double x, double *out) {
  for (int l = 0; l <= lmax; l++) out[l] = 0.0;
  if (m > lmax) return;
  const double w = std::sqrt(std::max(x * x - 1.0, 0.0));
  double pmm = 1.0;
  for (int k = 1; k <= m; k++) pmm *= (2 * k - 1) * w;  // Hobson: no (-1)^m
  out[m] = pmm;
  if (m + 1 <= lmax) out[m + 1] = (2 * m + 1) * x * pmm;
  for (int l = m + 1; l < lmax; l++) out[l + 1] = ((2 * l + 1) * x * out[l] - (l + m) * out[l - 1]) / (l + 1 - m);
}

static inline double q00(double x) { return 0.5 * (std::log(std::fabs(x + 1.0)) - std::log(std::fabs(x - 1.0))); }

// Modified Lentz evaluation of Q_n^m / Q_{n-1}^m from the three-term recurrence.
static double q_ratio_cf(double x, int n0, int m) {
  const double tiny = std::numeric_limits<double>::min() * 1e4;
  const double tol = 8.0 * std::numeric_limits<double>::epsilon();
  double f = tiny, C = f, D = 0.0, a = 1.0;
  int n = n0;
  for (int it = 0; it < 1000000; it++) {
    const double b = (2 * n + 1) * x / (n + m);
    D = b + a * D;
    if (D == 0.0) D = tiny;
    C = b + a / C;
    if (C == 0.0) C = tiny;
    D = 1.0 / D;
    const double delta = C * D;
    f *= delta;
    if (std::fabs(delta - 1.0) < tol) return f;
    a = -(double)(n - m + 1) / (n + m);
    n++;
  }
  throw std::runtime_error("legendre_q: continued fraction did not converge");
}

void legendre_q(int lmax, int M, double x, double *out) {
  if (!(x > 1.0)) throw std::domain_error("legendre_q: need x > 1");
  const int n = lmax + 1;
  std::vector<double> q0(n, 0.0), q1(n, 0.0);
  const double w = std::sqrt(std::max(x * x - 1.0, 0.0));
  const double Q00 = q00(x);
  const double xswitch = (lmax > 0) ? std::cosh(std::log(1e3) / (2.0 * lmax)) : std::numeric_limits<double>::infinity();
  if (x < xswitch) {
    // Christoffel form Q_l = P_l Q_0 - W_l (and its derivative for m = 1):
    // accurate close to x = 1 where Miller's continued fraction stalls.
    std::vector<double> P(n, 0.0), W(n, 0.0), Pp(n, 0.0), Wp(n, 0.0);
    P[0] = 1.0;
    if (lmax >= 1) {
      P[1] = x;
      W[1] = 1.0;
      Pp[1] = 1.0;
    }
    for (int k = 1; k < lmax; k++) {
      const double inv = 1.0 / (k + 1), a = 2 * k + 1, b = k;
      P[k + 1] = (a * x * P[k] - b * P[k - 1]) * inv;
      W[k + 1] = (a * x * W[k] - b * W[k - 1]) * inv;
      Pp[k + 1] = (a * (P[k] + x * Pp[k]) - b * Pp[k - 1]) * inv;
      Wp[k + 1] = (a * (W[k] + x * Wp[k]) - b * Wp[k - 1]) * inv;
    }
    for (int l = 0; l <= lmax; l++) q0[l] = P[l] * Q00 - W[l];
    q1[0] = -1.0 / w;
    for (int l = 1; l <= lmax; l++) q1[l] = w * Pp[l] * Q00 - P[l] / w - w * Wp[l];
  } else {
    // Miller: downward recurrence from a continued-fraction ratio, normalised
    // on the closed forms Q_0^0 and Q_0^1.
    q0[0] = Q00;
    q1[0] = -1.0 / w;
    if (lmax >= 1) {
      const double low = std::numeric_limits<double>::min() * 1e4;
      for (int m = 0; m <= 1; m++) {
        std::vector<double> &q = m ? q1 : q0;
        const double norm = m ? (-1.0 / w) : Q00;
        const double ratio = q_ratio_cf(x, lmax, m);
        q[lmax - 1] = low;
        q[lmax] = low * ratio;
        for (int l = lmax - 1; l >= 1; l--) q[l - 1] = ((2 * l + 1) * x * q[l] - (l + 1 - m) * q[l + 1]) / (l + m);
        const double sc = norm / q[0];
        for (int l = 0; l <= lmax; l++) q[l] *= sc;
      }
    }
  }
  if (M == 0) {
    std::copy(q0.begin(), q0.end(), out);
    return;
  }
  if (M == 1) {
    std::copy(q1.begin(), q1.end(), out);
    return;
  }
  // upward in m: Q_l^{m+1} = -2 m x / w Q_l^m + (l+m)(l-m+1) Q_l^{m-1}   (x > 1)
  for (int l = 0; l <= lmax; l++) {
    double a = q0[l], b = q1[l];
    for (int m = 1; m < M; m++) {
      const double c = -(2.0 * m) * x / w * b + (double)(l + m) * (double)(l - m + 1) * a;
      a = b;
      b = c;
    }
    out[l] = b;
  }
}

// ---------------------------------------------------------------------------
// Gaunt coefficients by quadrature
// ---------------------------------------------------------------------------

static void gauss_legendre(int n, std::vector<long double> &x, std::vector<long double> &w) {
  x.assign(n, 0.0L);
  w.assign(n, 0.0L);
  const long double pi = std::acos(-1.0L);
  for (int i = 0; i < (n + 1) / 2; i++) {
    long double z = std::cos(pi * (i + 0.75L) / (n + 0.5L)), pp = 0.0L;
    for (int it = 0; it < 100; it++) {
      long double p1 = 1.0L, p2 = 0.0L;
      for (int j = 1; j <= n; j++) {
        const long double p3 = p2;
        p2 = p1;
        p1 = ((2 * j - 1) * z * p2 - (j - 1) * p3) / j;
      }
      pp = n * (z * p1 - p2) / (z * z - 1.0L);
      const long double z1 = z;
      z = z1 - p1 / pp;
      if (std::fabs(z - z1) < 1e-19L) break;
    }
    // recompute derivative at the converged node
    long double p1 = 1.0L, p2 = 0.0L;
    for (int j = 1; j <= n; j++) {
      const long double p3 = p2;
      p2 = p1;
      p1 = ((2 * j - 1) * z * p2 - (j - 1) * p3) / j;
    }
    pp = n * (z * p1 - p2) / (z * z - 1.0L);
    x[i] = -z;
    x[n - 1 - i] = z;
    w[i] = w[n - 1 - i] = 2.0L / ((1.0L - z * z) * pp * pp);
  }
}

GauntTable::GauntTable(int lmax) : lmax_(lmax) {
  // integrand degree <= 3 lmax  ->  exact with nq >= (3 lmax)/2 + 1 points
  nq_ = (3 * lmax) / 2 + 2;
  gauss_legendre(nq_, xq_, wq_);
  const size_t nlm = (size_t)(lmax + 1) * (lmax + 2) / 2;
  theta_.assign(nlm * nq_, 0.0L);
  // Fully normalised Theta_l^m with Condon-Shortley phase:
  //   Y_l^m = Theta_l^m(cos th) e^{i m phi} / sqrt(2 pi),  int Theta^2 dx = 1
  for (int q = 0; q < nq_; q++) {
    const long double x = xq_[q], s = std::sqrt((1.0L - x) * (1.0L + x));
    long double pmm = std::sqrt(0.5L);  // Theta_0^0
    for (int m = 0; m <= lmax; m++) {
      if (m > 0) pmm = -pmm * s * std::sqrt((2.0L * m + 1.0L) / (2.0L * m));
      long double pl2 = 0.0L, pl1 = pmm;
      theta_[((size_t)m * (m + 1) / 2 + m) * nq_ + q] = pmm;
      for (int l = m + 1; l <= lmax; l++) {
        const long double a = std::sqrt((4.0L * l * l - 1.0L) / ((long double)l * l - (long double)m * m));
        const long double b = std::sqrt((((long double)l - 1) * (l - 1) - (long double)m * m) / (4.0L * (l - 1) * (l - 1) - 1.0L));
        const long double pl = a * (x * pl1 - b * pl2);
        theta_[((size_t)l * (l + 1) / 2 + m) * nq_ + q] = pl;
        pl2 = pl1;
        pl1 = pl;
      }
    }
  }
}

long double GauntTable::theta(int l, int m, int q) const {
  const int ma = m < 0 ? -m : m;
  const long double v = theta_[((size_t)l * (l + 1) / 2 + ma) * nq_ + q];
  return (m < 0 && (ma & 1)) ? -v : v;
}

double GauntTable::gaunt3(int l1, int m1, int l2, int m2, int l3, int m3) const {
  if (m1 + m2 + m3 != 0) return 0.0;
  if (l1 < 0 || l2 < 0 || l3 < 0) return 0.0;
  if (std::abs(m1) > l1 || std::abs(m2) > l2 || std::abs(m3) > l3) return 0.0;
  if ((l1 + l2 + l3) & 1) return 0.0;
  if (l3 < std::abs(l1 - l2) || l3 > l1 + l2) return 0.0;
  if (l1 > lmax_ || l2 > lmax_ || l3 > lmax_) throw std::logic_error("GauntTable: l outside table");
  long double s = 0.0L;
  for (int q = 0; q < nq_; q++) s += wq_[q] * theta(l1, m1, q) * theta(l2, m2, q) * theta(l3, m3, q);
  // (1/sqrt(2 pi))^3 * 2 pi
  return (double)(s / std::sqrt(2.0L * std::acos(-1.0L)));
}

double GauntTable::coeff(int L, int M, int l, int m, int lp) const {
  if (L < 0 || l < 0 || lp < 0) return 0.0;
  if (std::abs(M) > L || std::abs(m) > l) return 0.0;
  const int mp = M - m;
  if (std::abs(mp) > lp) return 0.0;
  const double v = gaunt3(L, -M, l, m, lp, mp);
  return (M & 1) ? -v : v;
}

double GauntTable::mod_coeff(int lj, int mj, int L, int M, int li, int mi) const {
  if (mj != M + mi) return 0.0;
  const double pi = std::acos(-1.0);
  const double c0 = 2.0 / 3.0 * std::sqrt(pi), c2 = 4.0 / 15.0 * std::sqrt(5.0 * pi);
  const double cpl0 = coeff(L, M, 0, 0, L) * coeff(lj, mj, li, mi, L);
  double cpl2 = 0.0;
  for (int Lp = std::max(std::max(L - 2, 0), std::abs(M)); Lp <= L + 2; Lp++)
    cpl2 += coeff(Lp, M, 2, 0, L) * coeff(lj, mj, li, mi, Lp);
  return c0 * cpl0 + c2 * cpl2;
}

}  // namespace hfq
