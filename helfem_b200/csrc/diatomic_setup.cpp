// Host-side setup of the diatomic prolate-spheroidal basis: multipole channel
// list, cross-element Legendre-weighted integrals and the in-element 2-channel
// two-electron kernel in sign-aware Cholesky form.
//
// Reference behaviour: src/diatomic/basis.cpp:505-520 (angular ordering),
// :525-647 (channel maps), :352-376 (P/Q integrals), :1334-1380 (order probe),
// :1382-1547 (compute_tei), :1549-1560 (prefactor);
// src/diatomic/quadrature.cpp:133-257 (nested Gauss-Chebyshev quadrature),
// src/diatomic/quadrature.h:47-84 (Legendre value filtering).
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <set>

#include "fem.h"
#include "special.h"
#include "tables.h"
#include "tei_device.h"

namespace hfq {

namespace {

constexpr int kOrderCap = 512;
constexpr double kCdThresh = 1e-12;  // relative, basis.cpp:1462

inline double keep_normal(double v) { return (v != 0.0 && !std::isnormal(v)) ? 0.0 : v; }

// Legendre values at a point set for one |M| and all L <= Lhi: tab[L*npts + q].
struct LegTable {
  int npts = 0;
  std::vector<double> P, Q;
};

LegTable legendre_table(int Mabs, int Lhi, const std::vector<double> &chmu, bool wantP, bool wantQ) {
  LegTable t;
  t.npts = (int)chmu.size();
  std::vector<double> buf(Lhi + 1);
  if (wantP) {
    t.P.assign((size_t)(Lhi + 1) * t.npts, 0.0);
    for (int q = 0; q < t.npts; q++) {
      legendre_p(Lhi, Mabs, chmu[q], buf.data());
      for (int L = 0; L <= Lhi; L++) t.P[(size_t)L * t.npts + q] = keep_normal(buf[L]);
    }
  }
  if (wantQ) {
    t.Q.assign((size_t)(Lhi + 1) * t.npts, 0.0);
    for (int q = 0; q < t.npts; q++) {
      if (chmu[q] == 1.0) continue;  // logarithmic singularity: reported as zero
      legendre_q(Lhi, Mabs, chmu[q], buf.data());
      for (int L = 0; L <= Lhi; L++) t.Q[(size_t)L * t.npts + q] = keep_normal(buf[L]);
    }
  }
  return t;
}

// Quantities of the nested rule that depend only on (element, order).
struct NestedRule {
  int n = 0, nbf = 0, npair = 0;
  double mulen = 0.0;
  std::vector<double> x, w, chmu, shmu;
  std::vector<int> pi, pj;           // unique pairs i <= j
  Mat bfprod;                        // (n x npair) outer B_i B_j
  std::vector<Mat> subbf;            // per sub-interval (n x nbf)
  std::vector<double> sublen;
  std::vector<double> subch, subsh;  // n*n, sub-interval ip at [ip*n, (ip+1)*n)
};

NestedRule make_nested_rule(const FEBasis &fe, int iel, int n) {
  NestedRule nr;
  nr.n = n;
  chebyshev_rule(n, nr.x, nr.w);
  const std::vector<int> en = fe.enabled(iel);
  nr.nbf = (int)en.size();
  for (int i = 0; i < nr.nbf; i++)
    for (int j = i; j < nr.nbf; j++) {
      nr.pi.push_back(i);
      nr.pj.push_back(j);
    }
  nr.npair = (int)nr.pi.size();
  const double mumin = fe.begin(iel), mumax = fe.end(iel);
  const double mumid = 0.5 * (mumax + mumin);
  nr.mulen = 0.5 * (mumax - mumin);
  std::vector<double> mu(n);
  nr.chmu.resize(n);
  nr.shmu.resize(n);
  for (int q = 0; q < n; q++) {
    mu[q] = mumid + nr.mulen * nr.x[q];
    nr.chmu[q] = std::cosh(mu[q]);
    nr.shmu[q] = std::sinh(mu[q]);
  }
  auto eval = [&](const std::vector<double> &xp) {
    const Mat prim = lip_eval(xp, fe.nodes(), 0);
    Mat out((int)xp.size(), nr.nbf);
    for (int k = 0; k < nr.nbf; k++)
      for (size_t q = 0; q < xp.size(); q++) out((int)q, k) = prim((int)q, en[k]);
    return out;
  };
  const Mat bf = eval(nr.x);
  nr.bfprod = Mat(n, nr.npair);
  for (int p = 0; p < nr.npair; p++)
    for (int q = 0; q < n; q++) nr.bfprod(q, p) = bf(q, nr.pi[p]) * bf(q, nr.pj[p]);
  nr.subbf.resize(n);
  nr.sublen.resize(n);
  nr.subch.resize((size_t)n * n);
  nr.subsh.resize((size_t)n * n);
  for (int ip = 0; ip < n; ip++) {
    const double a = (ip == 0) ? mumin : mu[ip - 1], b = mu[ip];
    const double smid = 0.5 * (b + a), slen = 0.5 * (b - a);
    std::vector<double> xp(n);
    for (int q = 0; q < n; q++) {
      const double smu = smid + slen * nr.x[q];
      nr.subch[(size_t)ip * n + q] = std::cosh(smu);
      nr.subsh[(size_t)ip * n + q] = std::sinh(smu);
      xp[q] = (smu - mumid) / nr.mulen;
    }
    nr.subbf[ip] = eval(xp);
    nr.sublen[ip] = slen;
  }
  return nr;
}

// The symmetrised 2-channel kernel W = [[T00,-T02],[-T02^T,T22]] (2 nbf^2 square)
// for one (L,|M|), from Legendre values Psub (at the n*n sub-interval points)
// and Qout (at the n outer points).  Pair symmetry (i<->j) is exploited: the
// integrals are evaluated on the nbf(nbf+1)/2 unique products and expanded.
Mat kernel_W(const NestedRule &nr, const double *Psub, const double *Qout) {
  const int n = nr.n, nbf = nr.nbf, np = nr.npair;
  // cumulative inner integrals for cosh^0 and cosh^2 weights: inner[l][(ip, pair)]
  std::vector<double> inner0((size_t)n * np, 0.0), inner2((size_t)n * np, 0.0);
  std::vector<double> acc0(np, 0.0), acc2(np, 0.0), w0(n), w2(n);
  for (int ip = 0; ip < n; ip++) {
    const Mat &bf = nr.subbf[ip];
    for (int q = 0; q < n; q++) {
      const double ch = nr.subch[(size_t)ip * n + q];
      w0[q] = nr.sublen[ip] * nr.w[q] * nr.subsh[(size_t)ip * n + q] * Psub[(size_t)ip * n + q];
      w2[q] = w0[q] * ch * ch;
    }
    for (int p = 0; p < np; p++) {
      const double *bi = &bf.a[(size_t)nr.pi[p] * n], *bj = &bf.a[(size_t)nr.pj[p] * n];
      double s0 = 0.0, s2 = 0.0;
      for (int q = 0; q < n; q++) {
        const double bb = bi[q] * bj[q];
        s0 += w0[q] * bb;
        s2 += w2[q] * bb;
      }
      acc0[p] += s0;
      acc2[p] += s2;
      inner0[(size_t)p * n + ip] = acc0[p];
      inner2[(size_t)p * n + ip] = acc2[p];
    }
  }
  // outer weights with Q_L^M and cosh^k
  std::vector<double> wb0((size_t)n * np), wb2((size_t)n * np);
  for (int p = 0; p < np; p++)
    for (int q = 0; q < n; q++) {
      const double wq = nr.mulen * nr.w[q] * nr.shmu[q] * Qout[q] * nr.bfprod(q, p);
      wb0[(size_t)p * n + q] = wq;
      wb2[(size_t)p * n + q] = wq * nr.chmu[q] * nr.chmu[q];
    }
  // O_kl[(pair r),(pair c)] = sum_q wb_k[r][q] inner_l[c][q]
  auto outer = [&](const std::vector<double> &wb, const std::vector<double> &in, std::vector<double> &O) {
    O.assign((size_t)np * np, 0.0);
    for (int c = 0; c < np; c++) {
      const double *ic = &in[(size_t)c * n];
      for (int r = 0; r < np; r++) {
        const double *wr = &wb[(size_t)r * n];
        double s = 0.0;
        for (int q = 0; q < n; q++) s += wr[q] * ic[q];
        O[(size_t)r + (size_t)c * np] = s;
      }
    }
  };
  std::vector<double> O00, O02, O20, O22;
  outer(wb0, inner0, O00);
  outer(wb0, inner2, O02);
  outer(wb2, inner0, O20);
  outer(wb2, inner2, O22);
  // T_kl = O_kl + O_lk^T
  const int nn = nbf * nbf;
  std::vector<int> pairof((size_t)nn);
  for (int p = 0; p < np; p++) {
    pairof[nr.pi[p] + nr.pj[p] * nbf] = p;
    pairof[nr.pj[p] + nr.pi[p] * nbf] = p;
  }
  Mat W(2 * nn, 2 * nn);
  for (int c = 0; c < nn; c++) {
    const int pc = pairof[c];
    for (int r = 0; r < nn; r++) {
      const int pr = pairof[r];
      const double t00 = O00[pr + (size_t)pc * np] + O00[pc + (size_t)pr * np];
      const double t02 = O02[pr + (size_t)pc * np] + O20[pc + (size_t)pr * np];
      const double t22 = O22[pr + (size_t)pc * np] + O22[pc + (size_t)pr * np];
      W(r, c) = t00;
      W(r, nn + c) = -t02;
      W(nn + c, r) = -t02;
      W(nn + r, nn + c) = t22;
    }
  }
  // remove roundoff asymmetry exactly as the reference does
  for (int c = 0; c < 2 * nn; c++)
    for (int r = 0; r < c; r++) {
      const double v = 0.5 * (W(r, c) + W(c, r));
      W(r, c) = v;
      W(c, r) = v;
    }
  return W;
}

// W ~= B diag(sigma) B^T, pivot on the largest |residual diagonal|, stop at
// thresh * initial max (basis.cpp:1498-1537).
void sign_cholesky(const Mat &W, double thresh, std::vector<double> &B, std::vector<double> &sigma, int &rank) {
  const int N = W.rows;
  std::vector<double> d(N);
  double dmax0 = 0.0;
  for (int i = 0; i < N; i++) {
    d[i] = W(i, i);
    dmax0 = std::max(dmax0, std::fabs(d[i]));
  }
  B.clear();
  sigma.clear();
  rank = 0;
  for (int p = 0; p < N; p++) {
    int piv = 0;
    double dpiv = -1.0;
    for (int i = 0; i < N; i++)
      if (std::fabs(d[i]) > dpiv) {
        dpiv = std::fabs(d[i]);
        piv = i;
      }
    if (dmax0 <= 0.0 || dpiv <= thresh * dmax0) break;
    const double dp = d[piv], s = dp >= 0.0 ? 1.0 : -1.0;
    B.resize((size_t)(rank + 1) * N);
    double *col = &B[(size_t)rank * N];
    for (int i = 0; i < N; i++) col[i] = W(i, piv);
    for (int q = 0; q < rank; q++) {
      const double *bq = &B[(size_t)q * N];
      const double f = sigma[q] * bq[piv];
      for (int i = 0; i < N; i++) col[i] -= f * bq[i];
    }
    const double inv = 1.0 / std::sqrt(std::fabs(dp));
    for (int i = 0; i < N; i++) col[i] *= inv;
    for (int i = 0; i < N; i++) d[i] -= s * col[i] * col[i];
    d[piv] = 0.0;
    sigma.push_back(s);
    rank++;
  }
}

Mat weighted_element(const FEBasis &fe, int iel, int der, const std::vector<double> &x, const std::vector<double> &w,
                     const std::function<double(double)> &f) {
  const std::vector<double> mu = fe.coord(x, iel);
  std::vector<double> wp(x.size());
  for (size_t q = 0; q < x.size(); q++) {
    const double fv = f ? f(mu[q]) : 1.0;
    wp[q] = std::isfinite(fv) ? w[q] * fe.scale(iel) * fv : 0.0;
  }
  const Mat bf = fe.eval_dnf(x, der, iel);
  return weighted_gram(bf, wp, bf);
}

}  // namespace

BasisTables build_diatomic_tables(int Z1, int Z2, double Rbond, const std::vector<int> &lmax_per_m, int nelem,
                                  int nnodes, double Rmax, int igrid, double zexp, int nquad, int device) {
  BasisTables t;
  t.kind = BasisKind::Diatomic;
  t.nch = 2;
  t.Z1 = Z1;
  t.Z2 = Z2;
  t.Rhalf = 0.5 * Rbond;
  t.nnodes = nnodes;
  t.nquad = nquad > 0 ? nquad : 5 * nnodes;
  t.drop_first_m_nonzero = true;
  t.sign_by_M = true;
  t.Lext = 2;
  const double mumax = std::acosh(Rmax / t.Rhalf);
  t.bval = element_grid(mumax, nelem, igrid, zexp);
  const FEBasis fe(nnodes, t.bval, false, true);
  t.Nrad = fe.nbf();
  t.Nel = fe.nel();
  for (int e = 0; e < t.Nel; e++) {
    t.efirst.push_back(fe.first(e));
    t.en.push_back(fe.nprim(e));
  }
  for (int mabs = 0; mabs < (int)lmax_per_m.size(); mabs++)
    for (int l = mabs; l <= lmax_per_m[mabs]; l++) {
      t.lval.push_back(l);
      t.mval.push_back(mabs);
      if (mabs > 0) {
        t.lval.push_back(l);
        t.mval.push_back(-mabs);
      }
    }
  // channel list: every (L,|M|) reachable from a pair of basis functions, |M|-major
  std::set<std::pair<int, int>> lm;  // (|M|, L)
  const int na = t.Nang();
  for (int i = 0; i < na; i++)
    for (int j = 0; j < na; j++) {
      const int M = std::abs(t.mval[j] - t.mval[i]);
      const int L0 = std::max(std::abs(t.lval[j] - t.lval[i]) - 2, M), L1 = t.lval[j] + t.lval[i] + 2;
      for (int L = L0; L <= L1; L++) lm.insert({M, L});
    }
  const double pi = std::acos(-1.0);
  for (const auto &p : lm) {
    t.lmL.push_back(p.second);
    t.lmM.push_back(p.first);
    double fr = 1.0;
    for (int k = p.second + p.first; k > p.second - p.first; k--) fr *= k;
    t.pref.push_back(4.0 * pi * std::pow(t.Rhalf, 5) / fr);
  }
  const int nlm = (int)t.lmL.size();
  t.blocks.resize((size_t)nlm * t.Nel);

  // |M| runs
  struct Run { int lo, hi, M, Lhi; };
  std::vector<Run> runs;
  for (int i = 0; i < nlm;) {
    int j = i, Lhi = 0;
    while (j < nlm && t.lmM[j] == t.lmM[i]) Lhi = std::max(Lhi, t.lmL[j++]);
    runs.push_back({i, j, t.lmM[i], Lhi});
    i = j;
  }
  const int nseed = std::max(t.nquad, 5);
  const double floor_rel = 256.0 * std::numeric_limits<double>::epsilon();

  const bool timing = getenv("HFQ_SETUP_TIMING") != nullptr;
  auto now = [] { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
  const double t_start = now();
  // ---- cross-element factors int B_i B_j sinh cosh^k {P,Q}_L^|M|(cosh mu) dmu
  const int ntask = (int)runs.size() * t.Nel;
#pragma omp parallel for schedule(dynamic)
  for (int task = 0; task < ntask; task++) {
    const Run &run = runs[task / t.Nel];
    const int iel = task % t.Nel;
    struct Pts { std::vector<double> x, w, ch, sh; LegTable leg; Mat bf; };
    std::map<int, std::shared_ptr<Pts>> cache;  // per quadrature order
    auto pts = [&](int n) {
      auto it = cache.find(n);
      if (it != cache.end()) return it->second;
      auto p = std::make_shared<Pts>();
      chebyshev_rule(n, p->x, p->w);
      const std::vector<double> mu = fe.coord(p->x, iel);
      p->ch.resize(n);
      p->sh.resize(n);
      for (int q = 0; q < n; q++) {
        p->ch[q] = std::cosh(mu[q]);
        p->sh[q] = std::sinh(mu[q]);
      }
      p->leg = legendre_table(run.M, run.Lhi, p->ch, true, true);
      p->bf = fe.eval_dnf(p->x, 0, iel);
      cache[n] = p;
      return p;
    };
    for (int ilm = run.lo; ilm < run.hi; ilm++) {
      const int L = t.lmL[ilm];
      ChannelBlock &b = t.blocks[(size_t)ilm * t.Nel + iel];
      b.n = fe.nprim(iel);
      b.small.assign((size_t)2 * b.n * b.n, 0.0);
      b.big.assign((size_t)2 * b.n * b.n, 0.0);
      for (int fam = 0; fam < 2; fam++)      // 0: P (inner element), 1: Q (outer element)
        for (int ch = 0; ch < 2; ch++) {     // 0: cosh^0, 1: cosh^2
          const Mat blk = converge_block(
              [&](int n) {
                const auto p = pts(n);
                const double *tab = fam ? &p->leg.Q[(size_t)L * n] : &p->leg.P[(size_t)L * n];
                std::vector<double> wp(n);
                for (int q = 0; q < n; q++) {
                  double fv = p->sh[q] * tab[q];
                  if (ch) fv *= p->ch[q] * p->ch[q];
                  wp[q] = std::isfinite(fv) ? p->w[q] * fe.scale(iel) * fv : 0.0;
                }
                return weighted_gram(p->bf, wp, p->bf);
              },
              nseed, kOrderCap, floor_rel, /*seed_fallback=*/true);
          std::vector<double> &dst = fam ? b.big : b.small;
          std::copy(blk.a.begin(), blk.a.end(), dst.begin() + (size_t)ch * b.n * b.n);
        }
    }
  }

  const double t_cross = now();
  // ---- in-element kernel: converge the rule on the hardest multipole, then
  //      build every (L,|M|) once at that order and factorise.
  int Lhard = t.lmL[0], Mhard = t.lmM[0];
  for (int i = 0; i < nlm; i++)
    if (t.lmL[i] > Lhard || (t.lmL[i] == Lhard && t.lmM[i] > Mhard)) {
      Lhard = t.lmL[i];
      Mhard = t.lmM[i];
    }
  std::unique_ptr<TeiDevice> teidev;
  for (int iel = 0; iel < t.Nel; iel++) {
    int nconv = std::min(nseed, kOrderCap);
    std::shared_ptr<NestedRule> last;
    converge_block(
        [&](int n) {
          last = std::make_shared<NestedRule>(make_nested_rule(fe, iel, n));
          const LegTable ps = legendre_table(Mhard, Lhard, last->subch, true, false);
          const LegTable qo = legendre_table(Mhard, Lhard, last->chmu, false, true);
          return kernel_W(*last, &ps.P[(size_t)Lhard * ps.npts], &qo.Q[(size_t)Lhard * qo.npts]);
        },
        std::min(nseed, kOrderCap), kOrderCap, floor_rel, false, &nconv);
    if (!last || last->n != nconv) last = std::make_shared<NestedRule>(make_nested_rule(fe, iel, nconv));
    const NestedRule &nr = *last;
    if (device >= 0) {
      // every channel of the element on the GPU (tei_device.cu): the host hands over the rule's point data and the
      // Legendre Q values at the n outer points; P at the n^2 inner points, the quadratures and the factorisations
      // run on the device
      if (!teidev) teidev.reset(new TeiDevice(device));
      const int n = nr.n;
      std::vector<double> cw((size_t)n * n), sub((size_t)n * n * nr.nbf), uw(n), bfp((size_t)nr.npair * n);
      for (int ip = 0; ip < n; ip++) {
        for (int q = 0; q < n; q++) cw[(size_t)ip * n + q] = nr.sublen[ip] * nr.w[q] * nr.subsh[(size_t)ip * n + q];
        std::copy(nr.subbf[ip].a.begin(), nr.subbf[ip].a.end(), sub.begin() + (size_t)ip * n * nr.nbf);
      }
      for (int q = 0; q < n; q++) uw[q] = nr.mulen * nr.w[q] * nr.shmu[q];
      for (int p = 0; p < nr.npair; p++)
        for (int q = 0; q < n; q++) bfp[(size_t)p * n + q] = nr.bfprod(q, p);
      TeiRuleView view;
      view.n = n;
      view.nbf = nr.nbf;
      view.npair = nr.npair;
      view.cw = cw.data();
      view.subch = nr.subch.data();
      view.subbf = sub.data();
      view.uw = uw.data();
      view.chmu = nr.chmu.data();
      view.bfprod = bfp.data();
      view.pi = nr.pi.data();
      view.pj = nr.pj.data();
      teidev->set_rule(view);
      for (const Run &run : runs) {
        const LegTable qo = legendre_table(run.M, run.Lhi, nr.chmu, false, true);
        std::vector<int> Ls;
        for (int ilm = run.lo; ilm < run.hi; ilm++) Ls.push_back(t.lmL[ilm]);
        std::vector<TeiChannelResult> res(Ls.size());
        teidev->run(run.M, run.Lhi, qo.Q.data(), Ls.data(), (int)Ls.size(), kCdThresh, res.data());
        for (int ilm = run.lo; ilm < run.hi; ilm++) {
          ChannelBlock &b = t.blocks[(size_t)ilm * t.Nel + iel];
          TeiChannelResult &r = res[ilm - run.lo];
          b.B = std::move(r.B);
          b.sigma = std::move(r.sigma);
          b.rank = r.rank;
        }
      }
      if (timing) std::fprintf(stderr, "[hfq setup] element %d: nested rule order %d (device)\n", iel, nconv);
      continue;
    }
    for (const Run &run : runs) {
      const LegTable ps = legendre_table(run.M, run.Lhi, nr.subch, true, false);
      const LegTable qo = legendre_table(run.M, run.Lhi, nr.chmu, false, true);
#pragma omp parallel for schedule(dynamic)
      for (int ilm = run.lo; ilm < run.hi; ilm++) {
        const int L = t.lmL[ilm];
        const Mat W = kernel_W(nr, &ps.P[(size_t)L * ps.npts], &qo.Q[(size_t)L * qo.npts]);
        ChannelBlock &b = t.blocks[(size_t)ilm * t.Nel + iel];
        sign_cholesky(W, kCdThresh, b.B, b.sigma, b.rank);
      }
    }
    if (timing) std::fprintf(stderr, "[hfq setup] element %d: nested rule order %d\n", iel, nconv);
  }
  if (timing)
    std::fprintf(stderr, "[hfq setup] cross-element factors %.3f s, in-element kernels + factorisation %.3f s\n",
                 t_cross - t_start, now() - t_cross);
  return t;
}

// One-electron matrices for the diatomic basis (basis.cpp:1032-1166).
void diatomic_one_electron_into(const BasisTables &t, double *S, double *T, double *V);

void diatomic_one_electron(const BasisTables &t, std::vector<double> &S, std::vector<double> &T,
                           std::vector<double> &V) {
  const size_t nb = (size_t)t.Nbf();
  S.resize(nb * nb);
  T.resize(nb * nb);
  V.resize(nb * nb);
  diatomic_one_electron_into(t, S.data(), T.data(), V.data());
}

// S, T, V: caller-owned Nbf x Nbf column-major matrices; every element is defined (zero-filled in parallel first)
void diatomic_one_electron_into(const BasisTables &t, double *S, double *T, double *V) {
  const FEBasis fe(t.nnodes, t.bval, false, true);
  const int N = t.Nrad, na = t.Nang();
  const int nseed = std::max(t.nquad, 5);
  const double floor_rel = 256.0 * std::numeric_limits<double>::epsilon();
  auto radial = [&](int der, const std::function<double(double)> &f) {
    return converge_block(
        [&](int n) {
          std::vector<double> x, w;
          lobatto_rule(n, x, w);
          Mat M(N, N);
          for (int e = 0; e < t.Nel; e++) {
            const Mat b = weighted_element(fe, e, der, x, w, f);
            const int f0 = fe.first(e), ne = fe.nprim(e);
            for (int j = 0; j < ne; j++)
              for (int i = 0; i < ne; i++) M(f0 + i, f0 + j) += b(i, j);
          }
          return M;
        },
        nseed, kOrderCap, floor_rel);
  };
  const Mat I10 = radial(0, [](double mu) { return std::sinh(mu); });
  const Mat I11 = radial(0, [](double mu) { return std::sinh(mu) * std::cosh(mu); });
  const Mat I12 = radial(0, [](double mu) { return std::sinh(mu) * std::cosh(mu) * std::cosh(mu); });
  const Mat Im1 = radial(0, [](double mu) { return 1.0 / std::sinh(mu); });
  const Mat Trad = radial(1, [](double mu) { return std::sinh(mu); });
  int lmax = 0;
  for (int l : t.lval) lmax = std::max(lmax, l);
  const GauntTable g(lmax + 2);
  const double pi = std::acos(-1.0);
  const double c0 = 2.0 / 3.0 * std::sqrt(pi), c2 = 4.0 / 15.0 * std::sqrt(5.0 * pi), c1 = 2.0 * std::sqrt(pi / 3.0);
  // The blocks are written straight into the boundary-free matrices (the reference assembles them with the dummy
  // functions and removes the boundaries afterwards, basis.cpp:1032-1166; for N2 that is 3 x 1.8 GB of intermediates
  // of which 0.4 GB are non-zero): dense index of (angular i, radial a) = first[i] + a - skip[i], where the m != 0
  // shells drop radial function 0.
  const std::vector<int64_t> pidx = t.pure_idx();
  const int nd = t.Ndummy(), nb = (int)pidx.size();
  std::vector<int64_t> pure_of((size_t)nd, -1);
  for (int k = 0; k < nb; k++) pure_of[pidx[k]] = k;
#pragma omp parallel for schedule(static)
  for (int c = 0; c < nb; c++) {
    std::memset(S + (size_t)c * nb, 0, (size_t)nb * sizeof(double));
    std::memset(T + (size_t)c * nb, 0, (size_t)nb * sizeof(double));
    std::memset(V + (size_t)c * nb, 0, (size_t)nb * sizeof(double));
  }
  const double R3 = std::pow(t.Rhalf, 3), R2 = t.Rhalf * t.Rhalf;
#pragma omp parallel for collapse(2) schedule(dynamic, 8)
  for (int i = 0; i < na; i++)
    for (int j = 0; j < na; j++) {
      if (t.mval[i] != t.mval[j]) continue;
      const int li = t.lval[i], lj = t.lval[j], m = t.mval[i];
      const double cos2 = c0 * g.coeff(lj, m, 0, 0, li) + c2 * g.coeff(lj, m, 2, 0, li);
      const double cos1 = c1 * g.coeff(lj, m, 1, 0, li);
      for (int b = 0; b < N; b++) {
        const int64_t pc = pure_of[(size_t)j * N + b];
        if (pc < 0) continue;
        for (int a = 0; a < N; a++) {
          const int64_t pr = pure_of[(size_t)i * N + a];
          if (pr < 0) continue;
          const size_t o = (size_t)pr + (size_t)pc * nb;
          double s = -cos2 * I10(a, b), v = 0.0;
          if (li == lj) {
            s += I12(a, b);
            v += (double)(t.Z1 + t.Z2) * I11(a, b);
          }
          if (t.Z1 != t.Z2) v += (double)(t.Z2 - t.Z1) * cos1 * I10(a, b);
          S[o] = R3 * s;
          V[o] = -R2 * v;
          if (i == j) T[o] = 0.5 * t.Rhalf * (Trad(a, b) + (double)li * (li + 1) * I10(a, b) + (double)m * m * Im1(a, b));
        }
      }
    }
}

void atomic_one_electron(const BasisTables &t, std::vector<double> &S, std::vector<double> &T, std::vector<double> &V);

void one_electron_matrices(const BasisTables &t, std::vector<double> &S, std::vector<double> &T, std::vector<double> &V) {
  if (t.kind != BasisKind::Diatomic)
    atomic_one_electron(t, S, T, V);
  else
    diatomic_one_electron(t, S, T, V);
}

}  // namespace hfq
