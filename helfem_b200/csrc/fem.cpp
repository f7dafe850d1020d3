#include "fem.h"

#include <algorithm>

namespace hfq {

double max_abs(const Mat &m) {
  double v = 0.0;
  for (double x : m.a) v = std::max(v, std::fabs(x));
  return v;
}

double max_abs_diff(const Mat &a, const Mat &b) {
  double v = 0.0;
  for (size_t i = 0; i < a.a.size(); i++) v = std::max(v, std::fabs(a.a[i] - b.a[i]));
  return v;
}

Mat weighted_gram(const Mat &A, const std::vector<double> &w, const Mat &B) {
  const int np = A.rows, na = A.cols, nb = B.cols;
  Mat C(na, nb);
  std::vector<double> wa((size_t)np);
  for (int i = 0; i < na; i++) {
    for (int q = 0; q < np; q++) wa[q] = A(q, i) * w[q];
    for (int j = 0; j < nb; j++) {
      double s = 0.0;
      const double *bj = &B.a[(size_t)j * np];
      for (int q = 0; q < np; q++) s += wa[q] * bj[q];
      C(i, j) = s;
    }
  }
  return C;
}

// Gauss-Lobatto: interior nodes are the roots of P'_{n-1}; Newton from the
// Chebyshev-Gauss-Lobatto guess (lobatto.h:36-109).
void lobatto_rule(int n, std::vector<double> &x, std::vector<double> &w) {
  if (n < 2) throw std::runtime_error("Lobatto rule needs n >= 2");
  x.assign(n, 0.0);
  w.assign(n, 0.0);
  const double pi = std::acos(-1.0);
  const double tol = 100.0 * std::numeric_limits<double>::epsilon();
  for (int i = 0; i < n; i++) x[i] = std::cos(pi * i / (n - 1));
  std::vector<double> pm1(n), pm2(n), xold(n);
  for (;;) {
    xold = x;
    double err = 0.0;
    for (int i = 0; i < n; i++) {
      double a = 1.0, b = x[i];  // P_0, P_1
      for (int j = 2; j <= n - 1; j++) {
        const double c = ((2 * j - 1) * x[i] * b - (j - 1) * a) / j;
        a = b;
        b = c;
      }
      // b = P_{n-1}, a = P_{n-2}
      const double xn = xold[i] - (x[i] * b - a) / (n * b);
      err = std::max(err, std::fabs(xn - xold[i]));
      x[i] = xn;
    }
    if (err <= tol) break;
  }
  std::reverse(x.begin(), x.end());
  for (int i = 0; i < n; i++) {
    double a = 1.0, b = x[i];
    for (int j = 2; j <= n - 1; j++) {
      const double c = ((2 * j - 1) * x[i] * b - (j - 1) * a) / j;
      a = b;
      b = c;
    }
    w[i] = 2.0 / ((double)(n - 1) * n * b * b);
  }
}

// Modified Gauss-Chebyshev rule of the second kind (chebyshev.h:32-57).
void chebyshev_rule(int n, std::vector<double> &x, std::vector<double> &w) {
  x.assign(n, 0.0);
  w.assign(n, 0.0);
  const double pi = std::acos(-1.0);
  for (int i = 1; i <= n; i++) {
    const double ang = i * pi / (n + 1);
    const double s = std::sin(ang), c = std::cos(ang);
    w[n - i] = 16.0 / 3.0 / (n + 1) * s * s * s * s;
    x[n - i] = 1.0 - 2.0 * i / (n + 1) + 2.0 / pi * (1.0 + 2.0 / 3.0 * s * s) * c * s;
  }
}

std::vector<double> element_grid(double rmax, int num_el, int igrid, double zexp) {
  std::vector<double> b(num_el + 1, 0.0);
  switch (igrid) {
    case 1:
      for (int i = 0; i <= num_el; i++) b[i] = rmax * i / num_el;
      break;
    case 2:
      for (int i = 0; i <= num_el; i++) b[i] = (double)i * i * rmax / ((double)num_el * num_el);
      break;
    case 3:
      for (int i = 0; i <= num_el; i++) b[i] = rmax * std::pow((double)i / num_el, zexp);
      break;
    case 4: {
      const double upper = std::pow(std::log(rmax + 1.0), 1.0 / zexp);
      for (int i = 0; i <= num_el; i++) {
        // same arithmetic as an evenly spaced (LinSpaced) parameter
        const double t = (i == num_el) ? upper : (upper / num_el) * i;
        b[i] = std::exp(std::pow(t, zexp)) - 1.0;
      }
      break;
    }
    default:
      throw std::logic_error("element_grid: grid type not supported");
  }
  b[0] = 0.0;
  b[num_el] = rmax;
  return b;
}

Mat lip_eval(const std::vector<double> &x, const std::vector<double> &x0, int n) {
  const int np = (int)x.size(), N = (int)x0.size();
  Mat out(np, N);
  if (n < 0 || n > 2) throw std::logic_error("lip_eval: derivative order not supported");
  for (int fi = 0; fi < N; fi++) {
    for (int q = 0; q < np; q++) {
      const double xv = x[q];
      double val = 0.0;
      if (n == 0) {
        double v = 1.0;
        for (int p = 0; p < N; p++)
          if (p != fi) v *= (xv - x0[p]) / (x0[fi] - x0[p]);
        val = v;
      } else if (n == 1) {
        for (int d1 = 0; d1 < N; d1++) {
          if (d1 == fi) continue;
          double v = 1.0;
          for (int p = 0; p < N; p++)
            if (p != fi && p != d1) v *= (xv - x0[p]) / (x0[fi] - x0[p]);
          val += v / (x0[fi] - x0[d1]);
        }
      } else {
        for (int d1 = 0; d1 < N; d1++) {
          if (d1 == fi) continue;
          for (int d2 = 0; d2 < d1; d2++) {
            if (d2 == fi) continue;
            double v = 1.0;
            for (int p = 0; p < N; p++)
              if (p != fi && p != d1 && p != d2) v *= (xv - x0[p]) / (x0[fi] - x0[p]);
            val += v / ((x0[fi] - x0[d1]) * (x0[fi] - x0[d2]));
          }
        }
        val *= 2.0;
      }
      out(q, fi) = val;
    }
  }
  return out;
}

FEBasis::FEBasis(int nnodes, const std::vector<double> &bval, bool zl, bool zr) : bval_(bval), zl_(zl), zr_(zr) {
  std::vector<double> w;
  lobatto_rule(nnodes, x0_, w);
  const int ne = nel();
  first_.assign(ne, 0);
  last_.assign(ne, 0);
  for (int iel = 0; iel < ne; iel++) {
    first_[iel] = (iel == 0) ? 0 : last_[iel - 1];  // one shared function
    last_[iel] = first_[iel] + nprim(iel) - 1;
  }
  nbf_ = last_[ne - 1] + 1;
}

std::vector<int> FEBasis::enabled(int iel) const {
  std::vector<int> en;
  const int N = nnodes();
  for (int i = 0; i < N; i++) {
    if (iel == 0 && zl_ && i == 0) continue;
    if (iel == nel() - 1 && zr_ && i == N - 1) continue;
    en.push_back(i);
  }
  return en;
}

std::vector<double> FEBasis::coord(const std::vector<double> &x, int iel) const {
  std::vector<double> r(x.size());
  const double m = mid(iel), s = scale(iel);
  for (size_t i = 0; i < x.size(); i++) r[i] = m + s * x[i];
  return r;
}

Mat FEBasis::eval_dnf(const std::vector<double> &x, int n, int iel) const {
  const Mat prim = lip_eval(x, x0_, n);
  const std::vector<int> en = enabled(iel);
  Mat out((int)x.size(), (int)en.size());
  const double sc = std::pow(scale(iel), n);
  for (size_t k = 0; k < en.size(); k++)
    for (size_t q = 0; q < x.size(); q++) out((int)q, (int)k) = prim((int)q, en[k]) / sc;
  return out;
}

// LIPBasis.h:87-112: divide out the (x+1) factor analytically.
Mat FEBasis::eval_over_r(const std::vector<double> &x, int n, int iel) const {
  if (std::fabs(begin(iel)) > 1e-14) throw std::logic_error("eval_over_r: element does not start at r=0");
  const std::vector<int> en = enabled(iel);
  if (en.empty() || en[0] == 0) throw std::logic_error("eval_over_r needs the first function dropped");
  std::vector<double> xr(x0_.begin() + 1, x0_.end());
  const Mat red = lip_eval(x, xr, n);
  const double sc = 1.0 / std::pow(scale(iel), n + 1);
  Mat out((int)x.size(), (int)en.size());
  for (size_t k = 0; k < en.size(); k++) {
    const double den = x0_[en[k]] + 1.0;
    for (size_t q = 0; q < x.size(); q++) out((int)q, (int)k) = sc * red((int)q, en[k] - 1) / den;
  }
  return out;
}

Mat converge_block(const std::function<Mat(int)> &probe, int nstart, int nmax, double floor_rel, bool seed_fallback,
                   int *nconv) {
  const double eps = std::numeric_limits<double>::epsilon();
  const double tol = 8.0 * eps, sqrteps = std::sqrt(eps);
  Mat prev, cur, seed;
  bool have = false;
  double prevdiff = -1.0, prevprevdiff = -1.0;
  int n = std::max(nstart, 2);
  for (;;) {
    cur = probe(n);
    if (!have) seed = cur;
    if (have) {
      const double diff = max_abs_diff(cur, prev), scale = max_abs(cur);
      bool done = diff <= tol * (scale + tol);
      if (!done && floor_rel < 0.0) {
        if (prevdiff >= 0.0 && diff <= sqrteps * (scale + tol) && diff > 0.5 * prevdiff) done = true;
        if (prevprevdiff >= 0.0 && diff <= sqrteps * (scale + tol) && diff > 0.125 * prevprevdiff) done = true;
      } else if (!done) {
        if (diff <= sqrteps * (scale + tol) &&
            (diff <= floor_rel * (scale + tol) || (prevdiff >= 0.0 && diff > 0.5 * prevdiff)))
          done = true;
      }
      if (done) {
        if (nconv) *nconv = n;
        return cur;
      }
      prevprevdiff = prevdiff;
      prevdiff = diff;
    }
    prev = cur;
    have = true;
    if (n >= nmax) {
      if (seed_fallback) {
        if (nconv) *nconv = std::max(nstart, 2);
        return seed;
      }
      if (nconv) *nconv = n;
      return cur;
    }
    n = std::min(2 * n, nmax);
  }
}

}  // namespace hfq
