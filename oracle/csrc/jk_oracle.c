/* CPU oracle, C restatement of the reference's J/K loops -- TEST INFRASTRUCTURE ONLY.
 *
 * Used (a) as the timed CPU baseline of bench.py (cpu_baseline / --impl reference:
 * the reference's own algorithm and OpenMP axes on the box's host cores) and (b) to
 * cross-check the numpy oracle at sizes where pure Python is too slow.  Nothing in
 * helfem_b200/ links or calls this file.
 *
 * Follows, loop for loop:
 *   diatomic exchange  src/diatomic/basis.cpp:1818-2089  (omp collapse(2) over output blocks)
 *   diatomic coulomb   src/diatomic/basis.cpp:1627-1816  (serial, as in the reference)
 *   atomic exchange    src/atomic/TwoDBasis.cpp:879-999 + CoulombExchangeFE.h:484-530
 *   atomic coulomb     src/atomic/TwoDBasis.cpp:773-877 + CoulombExchangeFE.h:432-482
 * Matrices are column-major.  Coupling coefficients come in as dense tables
 *   g0[(j*Nang+i)*NL+L] = mod_coeff(lj,mj,L,mj-mi,li,mi)   (diatomic channel "0")
 *   g2[(j*Nang+i)*NL+L] = coeff(lj,mj,L,mj-mi,li)          (diatomic channel "2", atomic)
 * and the integral caches as arrays of pointers indexed ilm*Nel+iel (diatomic) or L*Nel+iel (atomic).
 */
#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct {
  int Nang, Nrad, Nel, NL;
  const int *efirst, *en, *lval, *mval;
  const double *g0, *g2;
} basis_t;

static inline int imax(int a, int b) { return a > b ? a : b; }
static inline int imin(int a, int b) { return a < b ? a : b; }

int jk_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* launchers such as torchrun export OMP_NUM_THREADS=1; the timed CPU arm asks for the cores it may use */
void jk_set_num_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

/* C(m x n) += alpha * op(A) * op(B); tiny dense kernels, column-major */
static void gemm_nn(int m, int n, int k, double alpha, const double *A, int lda, const double *B, int ldb, double *C,
                    int ldc) {
  for (int j = 0; j < n; j++)
    for (int p = 0; p < k; p++) {
      const double b = alpha * B[p + (size_t)j * ldb];
      const double *a = A + (size_t)p * lda;
      double *c = C + (size_t)j * ldc;
      for (int i = 0; i < m; i++) c[i] += a[i] * b;
    }
}
static void gemm_tn(int m, int n, int k, double alpha, const double *A, int lda, const double *B, int ldb, double *C,
                    int ldc) { /* A is k x m, used transposed */
  for (int j = 0; j < n; j++)
    for (int i = 0; i < m; i++) {
      const double *a = A + (size_t)i * lda, *b = B + (size_t)j * ldb;
      double s = 0.0;
      for (int p = 0; p < k; p++) s += a[p] * b[p];
      C[i + (size_t)j * ldc] += alpha * s;
    }
}
static void gemm_nt(int m, int n, int k, double alpha, const double *A, int lda, const double *B, int ldb, double *C,
                    int ldc) { /* B is n x k, used transposed */
  for (int p = 0; p < k; p++)
    for (int j = 0; j < n; j++) {
      const double b = alpha * B[j + (size_t)p * ldb];
      const double *a = A + (size_t)p * lda;
      double *c = C + (size_t)j * ldc;
      for (int i = 0; i < m; i++) c[i] += a[i] * b;
    }
}

static double block_norm(const double *P, size_t ld, int N) {
  double s = 0.0;
  for (int j = 0; j < N; j++)
    for (int i = 0; i < N; i++) s += P[i + j * ld] * P[i + j * ld];
  return sqrt(s);
}

static int find_lm(int nlm, const int *lmL, const int *lmM, int L, int Mabs) {
  for (int i = 0; i < nlm; i++)
    if (lmL[i] == L && lmM[i] == Mabs) return i;
  return -1;
}

/* ------------------------------------------------------------------------------------------
 * Diatomic exchange, selected output blocks (basis.cpp:1818-2089).
 * P is the boundary-expanded density (Ndummy x Ndummy).  Kout holds nblocks blocks of
 * Nrad x Nrad (the reference's K.block(jang*Nrad, kang*Nrad), sign included).
 * ---------------------------------------------------------------------------------------- */
void jk_diatomic_exchange_blocks(int Nang, int Nrad, int Nel, int NL, const int *efirst, const int *en,
                                 const int *lval, const int *mval, const double *g0, const double *g2, int nlm,
                                 const int *lmL, const int *lmM, const double *LMfac_abs, const double *const *dP0,
                                 const double *const *dP2, const double *const *dQ0, const double *const *dQ2,
                                 const double *const *cdB, const double *const *cdS, const int *rank, const double *P,
                                 int nblocks, const int *jangs, const int *kangs, double *Kout) {
  const size_t Nd = (size_t)Nang * Nrad, NN = (size_t)Nrad * Nrad;
  int mabs = 0;
  for (int i = 0; i < Nang; i++) mabs = imax(mabs, abs(mval[i]));
  const int dmoff = 2 * mabs, ndm = 4 * mabs + 1;
  /* density blocks bucketed by dm = mi - ml (basis.cpp:1855-1864) */
  int *cnt = (int *)calloc(ndm, sizeof(int));
  int *pi = (int *)malloc(sizeof(int) * (size_t)Nang * Nang), *pl = (int *)malloc(sizeof(int) * (size_t)Nang * Nang);
  int *start = (int *)calloc(ndm + 1, sizeof(int));
  char *nz = (char *)calloc((size_t)Nang * Nang, 1);
  for (int i = 0; i < Nang; i++)
    for (int l = 0; l < Nang; l++) {
      if (block_norm(P + (size_t)i * Nrad + (size_t)l * Nrad * Nd, Nd, Nrad) < 10 * DBL_EPSILON) continue;
      nz[(size_t)i * Nang + l] = 1;
      cnt[mval[i] - mval[l] + dmoff]++;
    }
  for (int d = 0; d < ndm; d++) start[d + 1] = start[d] + cnt[d];
  memset(cnt, 0, sizeof(int) * ndm);
  for (int i = 0; i < Nang; i++)
    for (int l = 0; l < Nang; l++)
      if (nz[(size_t)i * Nang + l]) {
        const int d = mval[i] - mval[l] + dmoff;
        pi[start[d] + cnt[d]] = i;
        pl[start[d] + cnt[d]] = l;
        cnt[d]++;
      }
  int nmax = 0;
  for (int e = 0; e < Nel; e++) nmax = imax(nmax, en[e]);

#pragma omp parallel
  {
    double *R = (double *)malloc(sizeof(double) * 4 * NN * (size_t)nlm);
    char *couple = (char *)malloc(nlm);
    double *t1 = (double *)malloc(sizeof(double) * nmax * nmax), *t2 = (double *)malloc(sizeof(double) * nmax * nmax);
    double *ks = (double *)malloc(sizeof(double) * nmax * nmax);
    double *rb = (double *)malloc(sizeof(double) * 4 * nmax * nmax);
#pragma omp for schedule(dynamic)
    for (int blk = 0; blk < nblocks; blk++) {
      const int jang = jangs[blk], kang = kangs[blk];
      const int lj = lval[jang], mj = mval[jang], lk = lval[kang], mk = mval[kang];
      double *Kb = Kout + (size_t)blk * NN;
      memset(Kb, 0, sizeof(double) * NN);
      const int dmjk = mj - mk;
      if (abs(dmjk) > 2 * mabs) continue;
      memset(couple, 0, nlm);
      int any = 0;
      for (int c = start[dmjk + dmoff]; c < start[dmjk + dmoff + 1]; c++) {
        const int iang = pi[c], lang = pl[c];
        const int li = lval[iang], mi = mval[iang], ll = lval[lang];
        const int M = mj - mi;
        const int Lmin = imax(imax(abs(li - lj), abs(lk - ll)) - 2, abs(M));
        const int Lmax = imin(li + lj, lk + ll) + 2;
        const double *Psub = P + (size_t)iang * Nrad + (size_t)lang * Nrad * Nd;
        for (int L = Lmin; L <= Lmax && L < NL; L++) {
          const double mj0 = g0[((size_t)jang * Nang + iang) * NL + L], mk0 = g0[((size_t)kang * Nang + lang) * NL + L];
          const double cj2 = g2[((size_t)jang * Nang + iang) * NL + L], ck2 = g2[((size_t)kang * Nang + lang) * NL + L];
          const double c00 = mj0 * mk0, c02 = -mj0 * ck2, c20 = -cj2 * mk0, c22 = cj2 * ck2;
          if (c00 == 0.0 && c02 == 0.0 && c20 == 0.0 && c22 == 0.0) continue;
          const int ilm = find_lm(nlm, lmL, lmM, L, abs(M));
          const double LMfac = ((M & 1) ? -1.0 : 1.0) * LMfac_abs[ilm];
          double *R00 = R + (size_t)ilm * 4 * NN, *R02 = R00 + NN, *R20 = R02 + NN, *R22 = R20 + NN;
          if (!couple[ilm]) {
            memset(R00, 0, sizeof(double) * 4 * NN);
            couple[ilm] = 1;
            any = 1;
          }
          const double f00 = LMfac * c00, f02 = LMfac * c02, f20 = LMfac * c20, f22 = LMfac * c22;
          for (int c2 = 0; c2 < Nrad; c2++) {
            const double *ps = Psub + (size_t)c2 * Nd;
            const size_t o = (size_t)c2 * Nrad;
            for (int r = 0; r < Nrad; r++) {
              const double p = ps[r];
              R00[o + r] += f00 * p;
              R02[o + r] += f02 * p;
              R20[o + r] += f20 * p;
              R22[o + r] += f22 * p;
            }
          }
        }
      }
      if (!any) continue;
      for (int iel = 0; iel < Nel; iel++)
        for (int jel = 0; jel < Nel; jel++) {
          const int fi = efirst[iel], fj = efirst[jel], Ni = en[iel], Nj = en[jel];
          memset(ks, 0, sizeof(double) * Ni * Nj);
          for (int ilm = 0; ilm < nlm; ilm++) {
            if (!couple[ilm]) continue;
            const double *Rm = R + (size_t)ilm * 4 * NN;
            for (int ab = 0; ab < 4; ab++) /* materialise the element blocks like the reference does */
              for (int c2 = 0; c2 < Nj; c2++)
                for (int r = 0; r < Ni; r++) rb[ab * Ni * Nj + r + c2 * Ni] = Rm[ab * NN + (fi + r) + (size_t)(fj + c2) * Nrad];
            if (iel == jel) {
              /* RI-K: K -= sum_p sigma_p (M0' R00 M0' - M0' R02 M2' - M2' R20 M0' + M2' R22 M2') */
              const double *B = cdB[ilm * Nel + iel], *sg = cdS[ilm * Nel + iel];
              const int nn = Ni * Ni, r_ = rank[ilm * Nel + iel];
              for (int p = 0; p < r_; p++) {
                const double *M0 = B + (size_t)p * 2 * nn, *M2 = M0 + nn;
                /* t1 = R00 M0' - R02 M2' ; ks += sigma M0' t1 */
                memset(t1, 0, sizeof(double) * nn);
                gemm_nt(Ni, Ni, Ni, 1.0, rb, Ni, M0, Ni, t1, Ni);
                gemm_nt(Ni, Ni, Ni, -1.0, rb + nn, Ni, M2, Ni, t1, Ni);
                gemm_tn(Ni, Ni, Ni, sg[p], M0, Ni, t1, Ni, ks, Ni);
                /* t2 = -R20 M0' + R22 M2' ; ks += sigma M2' t2 */
                memset(t2, 0, sizeof(double) * nn);
                gemm_nt(Ni, Ni, Ni, -1.0, rb + 2 * nn, Ni, M0, Ni, t2, Ni);
                gemm_nt(Ni, Ni, Ni, 1.0, rb + 3 * nn, Ni, M2, Ni, t2, Ni);
                gemm_tn(Ni, Ni, Ni, sg[p], M2, Ni, t2, Ni, ks, Ni);
              }
            } else {
              const double *i0 = (iel > jel) ? dQ0[ilm * Nel + iel] : dP0[ilm * Nel + iel];
              const double *i2 = (iel > jel) ? dQ2[ilm * Nel + iel] : dP2[ilm * Nel + iel];
              const double *j0 = (iel > jel) ? dP0[ilm * Nel + jel] : dQ0[ilm * Nel + jel];
              const double *j2 = (iel > jel) ? dP2[ilm * Nel + jel] : dQ2[ilm * Nel + jel];
              memset(t1, 0, sizeof(double) * Ni * Nj);
              gemm_nt(Ni, Nj, Nj, 1.0, rb, Ni, j0, Nj, t1, Ni);
              gemm_nt(Ni, Nj, Nj, 1.0, rb + Ni * Nj, Ni, j2, Nj, t1, Ni);
              gemm_nn(Ni, Nj, Ni, 1.0, i0, Ni, t1, Ni, ks, Ni);
              memset(t1, 0, sizeof(double) * Ni * Nj);
              gemm_nt(Ni, Nj, Nj, 1.0, rb + 2 * Ni * Nj, Ni, j0, Nj, t1, Ni);
              gemm_nt(Ni, Nj, Nj, 1.0, rb + 3 * Ni * Nj, Ni, j2, Nj, t1, Ni);
              gemm_nn(Ni, Nj, Ni, 1.0, i2, Ni, t1, Ni, ks, Ni);
            }
          }
          for (int c2 = 0; c2 < Nj; c2++)
            for (int r = 0; r < Ni; r++) Kb[(fi + r) + (size_t)(fj + c2) * Nrad] -= ks[r + c2 * Ni];
        }
    }
    free(R); free(couple); free(t1); free(t2); free(ks); free(rb);
  }
  free(cnt); free(pi); free(pl); free(start); free(nz);
}

/* ------------------------------------------------------------------------------------------
 * Diatomic coulomb (basis.cpp:1627-1816), serial like the reference.  nLM channels (L,M)
 * given by LML/LMM; P and J are boundary-expanded Ndummy x Ndummy.
 * ---------------------------------------------------------------------------------------- */
/* stride > 1: TIMING SAMPLE ONLY -- the fold visits the angular rows k = 0, stride, 2 stride .., the radial step the
 * channels iLM = 0, stride, .., the unfold the rows i = 0, stride, ..; every one of the three phases is linear in
 * the number of rows / channels visited, so stride * (time of the sample) estimates the time of the full build.
 * The result of a sampled call is not a Coulomb matrix. */
#define JK_ALL_M (1 << 30)
static void coulomb_impl(int Nang, int Nrad, int Nel, int NL, const int *efirst, const int *en, const int *lval,
                         const int *mval, const double *g0, const double *g2, int nlm, const int *lmL, const int *lmM,
                         const double *LMfac_abs, int nLM, const int *LML, const int *LMM, const double *const *dP0,
                         const double *const *dP2, const double *const *dQ0, const double *const *dQ2,
                         const double *const *cdB, const double *const *cdS, const int *rank, const double *P,
                         double *J, int stride, int Monly) {
  /* Monly != JK_ALL_M: only the channels with M == Monly are built (fold, radial step and unfold skip every
   * other M).  For a density whose blocks all have m_k - m_l == Monly this IS the complete Coulomb matrix (all other
   * Paux vanish identically); used by bench.py to check J at full size without the 12/13 of the fold that adds
   * zeros.  The timed arm always runs the reference's unscreened loops (Monly = JK_ALL_M). */
  const size_t Nd = (size_t)Nang * Nrad, NN = (size_t)Nrad * Nrad;
  double *Paux0 = (double *)calloc(NN * nLM, sizeof(double)), *Paux2 = (double *)calloc(NN * nLM, sizeof(double));
  double *Jaux0 = (double *)calloc(NN * nLM, sizeof(double)), *Jaux2 = (double *)calloc(NN * nLM, sizeof(double));
  int Mlo = 0, Mhi = 0;
  for (int i = 0; i < nLM; i++) {
    Mlo = imin(Mlo, LMM[i]);
    Mhi = imax(Mhi, LMM[i]);
  }
  const int nMv = Mhi - Mlo + 1;
  int *LMidx = (int *)malloc(sizeof(int) * (size_t)NL * nMv);
  for (int i = 0; i < NL * nMv; i++) LMidx[i] = -1;
  for (int i = 0; i < nLM; i++) LMidx[LML[i] * nMv + LMM[i] - Mlo] = i;
  for (int k = 0; k < Nang; k += stride)
    for (int l = 0; l < Nang; l++) {
      const int M = mval[k] - mval[l];
      if (Monly != JK_ALL_M && M != Monly) continue;
      const int Lmin = imax(abs(lval[k] - lval[l]) - 2, abs(M)), Lmax = lval[k] + lval[l] + 2;
      const double *Prad = P + (size_t)k * Nrad + (size_t)l * Nrad * Nd;
      for (int L = Lmin; L <= Lmax; L++) {
        const int iLM = LMidx[L * nMv + M - Mlo];
        const double c0 = g0[((size_t)k * Nang + l) * NL + L], c2 = g2[((size_t)k * Nang + l) * NL + L];
        for (int c = 0; c < Nrad; c++)
          for (int r = 0; r < Nrad; r++) {
            const double p = Prad[r + (size_t)c * Nd];
            if (c0 != 0.0) Paux0[iLM * NN + r + (size_t)c * Nrad] += c0 * p;
            if (c2 != 0.0) Paux2[iLM * NN + r + (size_t)c * Nrad] += c2 * p;
          }
      }
    }
  int nmax = 0;
  for (int e = 0; e < Nel; e++) nmax = imax(nmax, en[e]);
  double *p2 = (double *)malloc(sizeof(double) * 2 * nmax * nmax), *cv = (double *)malloc(sizeof(double) * 2 * nmax * nmax);
  for (int iLM = 0; iLM < nLM; iLM += stride) {
    const int L = LML[iLM], M = LMM[iLM];
    if (Monly != JK_ALL_M && M != Monly) continue;
    const int ilm = find_lm(nlm, lmL, lmM, L, abs(M));
    const double LMfac = ((M & 1) ? -1.0 : 1.0) * LMfac_abs[ilm];
    double *Ja0 = Jaux0 + iLM * NN, *Ja2 = Jaux2 + iLM * NN;
    const double *Pa0 = Paux0 + iLM * NN, *Pa2 = Paux2 + iLM * NN;
    for (int jel = 0; jel < Nel; jel++) {
      const int fj = efirst[jel], Nj = en[jel];
      double js0 = 0, jb0 = 0, js2 = 0, jb2 = 0;
      for (int a = 0; a < Nj; a++)
        for (int b = 0; b < Nj; b++) { /* trace(D * Psub) = sum D(a,b) Psub(b,a) */
          const double ps0 = Pa0[(fj + b) + (size_t)(fj + a) * Nrad], ps2 = Pa2[(fj + b) + (size_t)(fj + a) * Nrad];
          js0 += dP0[ilm * Nel + jel][a + b * Nj] * ps0;
          jb0 += dQ0[ilm * Nel + jel][a + b * Nj] * ps0;
          js2 += dP2[ilm * Nel + jel][a + b * Nj] * ps2;
          jb2 += dQ2[ilm * Nel + jel][a + b * Nj] * ps2;
        }
      js0 *= LMfac; jb0 *= LMfac; js2 *= LMfac; jb2 *= LMfac;
      for (int iel = 0; iel < Nel; iel++) {
        if (iel == jel) continue;
        const int fi = efirst[iel], Ni = en[iel];
        const double f0 = (iel < jel) ? (jb0 - jb2) : (js0 - js2);
        const double *i0 = (iel < jel) ? dP0[ilm * Nel + iel] : dQ0[ilm * Nel + iel];
        const double *i2 = (iel < jel) ? dP2[ilm * Nel + iel] : dQ2[ilm * Nel + iel];
        for (int c = 0; c < Ni; c++)
          for (int r = 0; r < Ni; r++) {
            Ja0[(fi + r) + (size_t)(fi + c) * Nrad] += f0 * i0[r + c * Ni];
            Ja2[(fi + r) + (size_t)(fi + c) * Nrad] -= f0 * i2[r + c * Ni];
          }
      }
      const int nn = Nj * Nj, r_ = rank[ilm * Nel + jel];
      const double *B = cdB[ilm * Nel + jel], *sg = cdS[ilm * Nel + jel];
      for (int c = 0; c < Nj; c++)
        for (int r = 0; r < Nj; r++) {
          p2[r + c * Nj] = Pa0[(fj + r) + (size_t)(fj + c) * Nrad];
          p2[nn + r + c * Nj] = Pa2[(fj + r) + (size_t)(fj + c) * Nrad];
        }
      memset(cv, 0, sizeof(double) * 2 * nn);
      for (int p = 0; p < r_; p++) {
        const double *bp = B + (size_t)p * 2 * nn;
        double s = 0.0;
        for (int q = 0; q < 2 * nn; q++) s += bp[q] * p2[q];
        s *= sg[p] * LMfac;
        for (int q = 0; q < 2 * nn; q++) cv[q] += bp[q] * s;
      }
      for (int c = 0; c < Nj; c++)
        for (int r = 0; r < Nj; r++) {
          Ja0[(fj + r) + (size_t)(fj + c) * Nrad] += cv[r + c * Nj];
          Ja2[(fj + r) + (size_t)(fj + c) * Nrad] += cv[nn + r + c * Nj];
        }
    }
  }
  if (stride == 1) memset(J, 0, sizeof(double) * Nd * Nd);
  for (int i = 0; i < Nang; i += stride)
    for (int j = 0; j < Nang; j++) {
      const int M = mval[j] - mval[i];
      if (Monly != JK_ALL_M && M != Monly) continue;
      const int Lmin = imax(abs(lval[j] - lval[i]) - 2, abs(M)), Lmax = lval[j] + lval[i] + 2;
      double *Jb = J + (size_t)i * Nrad + (size_t)j * Nrad * Nd;
      for (int L = Lmin; L <= Lmax; L++) {
        const int iLM = LMidx[L * nMv + M - Mlo];
        const double c0 = g0[((size_t)j * Nang + i) * NL + L], c2 = g2[((size_t)j * Nang + i) * NL + L];
        if (c0 == 0.0 && c2 == 0.0) continue;
        for (int c = 0; c < Nrad; c++)
          for (int r = 0; r < Nrad; r++)
            Jb[r + (size_t)c * Nd] += c0 * Jaux0[iLM * NN + r + (size_t)c * Nrad] + c2 * Jaux2[iLM * NN + r + (size_t)c * Nrad];
      }
    }
  free(Paux0); free(Paux2); free(Jaux0); free(Jaux2); free(LMidx); free(p2); free(cv);
}

void jk_diatomic_coulomb(int Nang, int Nrad, int Nel, int NL, const int *efirst, const int *en, const int *lval,
                         const int *mval, const double *g0, const double *g2, int nlm, const int *lmL, const int *lmM,
                         const double *LMfac_abs, int nLM, const int *LML, const int *LMM, const double *const *dP0,
                         const double *const *dP2, const double *const *dQ0, const double *const *dQ2,
                         const double *const *cdB, const double *const *cdS, const int *rank, const double *P,
                         double *J) {
  coulomb_impl(Nang, Nrad, Nel, NL, efirst, en, lval, mval, g0, g2, nlm, lmL, lmM, LMfac_abs, nLM, LML, LMM, dP0, dP2, dQ0,
               dQ2, cdB, cdS, rank, P, J, 1, JK_ALL_M);
}

/* Coulomb matrix of a density that only has blocks with m_k - m_l == Monly (see coulomb_impl) */
void jk_diatomic_coulomb_single_M(int Nang, int Nrad, int Nel, int NL, const int *efirst, const int *en, const int *lval,
                                  const int *mval, const double *g0, const double *g2, int nlm, const int *lmL,
                                  const int *lmM, const double *LMfac_abs, int nLM, const int *LML, const int *LMM,
                                  const double *const *dP0, const double *const *dP2, const double *const *dQ0,
                                  const double *const *dQ2, const double *const *cdB, const double *const *cdS,
                                  const int *rank, const double *P, double *J, int Monly) {
  coulomb_impl(Nang, Nrad, Nel, NL, efirst, en, lval, mval, g0, g2, nlm, lmL, lmM, LMfac_abs, nLM, LML, LMM, dP0, dP2, dQ0,
               dQ2, cdB, cdS, rank, P, J, 1, Monly);
}

void jk_diatomic_coulomb_timing_sample(int Nang, int Nrad, int Nel, int NL, const int *efirst, const int *en,
                                       const int *lval, const int *mval, const double *g0, const double *g2, int nlm,
                                       const int *lmL, const int *lmM, const double *LMfac_abs, int nLM, const int *LML,
                                       const int *LMM, const double *const *dP0, const double *const *dP2,
                                       const double *const *dQ0, const double *const *dQ2, const double *const *cdB,
                                       const double *const *cdS, const int *rank, const double *P, double *J,
                                       int stride) {
  coulomb_impl(Nang, Nrad, Nel, NL, efirst, en, lval, mval, g0, g2, nlm, lmL, lmM, LMfac_abs, nLM, LML, LMM, dP0, dP2, dQ0,
               dQ2, cdB, cdS, rank, P, J, stride < 1 ? 1 : stride, JK_ALL_M);
}

/* ------------------------------------------------------------------------------------------
 * Atomic exchange, selected output blocks (TwoDBasis.cpp:879-999, CoulombExchangeFE.h:484-530).
 * caches indexed L*Nel+iel; g2 = coeff table.
 * ---------------------------------------------------------------------------------------- */
void jk_atomic_exchange_blocks(int Nang, int Nrad, int Nel, int NL, const int *efirst, const int *en, const int *lval,
                               const int *mval, const double *g2, const double *const *dsmall,
                               const double *const *dbig, const double *const *chol, const int *rank, const double *P,
                               int nblocks, const int *jangs, const int *kangs, double *Kout) {
  const size_t Nd = (size_t)Nang * Nrad, NN = (size_t)Nrad * Nrad;
  double *norms = (double *)malloc(sizeof(double) * (size_t)Nang * Nang);
  for (int i = 0; i < Nang; i++)
    for (int l = 0; l < Nang; l++) norms[(size_t)i * Nang + l] = block_norm(P + (size_t)i * Nrad + (size_t)l * Nrad * Nd, Nd, Nrad);
  int nmax = 0;
  for (int e = 0; e < Nel; e++) nmax = imax(nmax, en[e]);
  const double pi = acos(-1.0);
#pragma omp parallel
  {
    double *R = (double *)malloc(sizeof(double) * NN * NL);
    char *couple = (char *)malloc(NL);
    double *t1 = (double *)malloc(sizeof(double) * nmax * nmax), *ps = (double *)malloc(sizeof(double) * nmax * nmax);
#pragma omp for schedule(dynamic)
    for (int blk = 0; blk < nblocks; blk++) {
      const int jang = jangs[blk], kang = kangs[blk];
      const int lj = lval[jang], mj = mval[jang], lk = lval[kang], mk = mval[kang];
      double *Kb = Kout + (size_t)blk * NN;
      memset(Kb, 0, sizeof(double) * NN);
      memset(R, 0, sizeof(double) * NN * NL);
      memset(couple, 0, NL);
      for (int iang = 0; iang < Nang; iang++)
        for (int lang = 0; lang < Nang; lang++) {
          const int li = lval[iang], mi = mval[iang], ll = lval[lang], ml = mval[lang];
          const int M = mj - mi;
          if (M != mk - ml) continue;
          if (norms[(size_t)iang * Nang + lang] < 10 * DBL_EPSILON) continue;
          const int Lmin = imax(imax(abs(li - lj), abs(lk - ll)), abs(M)), Lmax = imin(li + lj, lk + ll);
          const double *Psub = P + (size_t)iang * Nrad + (size_t)lang * Nrad * Nd;
          for (int L = Lmin; L <= Lmax; L++) {
            const double cpl = g2[((size_t)jang * Nang + iang) * NL + L] * g2[((size_t)kang * Nang + lang) * NL + L];
            if (cpl == 0.0) continue;
            const double f = 4.0 * pi / (2 * L + 1) * cpl;
            double *RL = R + (size_t)L * NN;
            for (int c = 0; c < Nrad; c++)
              for (int r = 0; r < Nrad; r++) RL[r + (size_t)c * Nrad] += f * Psub[r + (size_t)c * Nd];
            couple[L] = 1;
          }
        }
      for (int L = 0; L < NL; L++) {
        if (!couple[L]) continue;
        const double *RL = R + (size_t)L * NN;
        for (int iel = 0; iel < Nel; iel++)
          for (int jel = 0; jel < Nel; jel++) {
            const int fi = efirst[iel], fj = efirst[jel], Ni = en[iel], Nj = en[jel];
            for (int c = 0; c < Nj; c++)
              for (int r = 0; r < Ni; r++) ps[r + c * Ni] = RL[(fi + r) + (size_t)(fj + c) * Nrad];
            double *Kdst = Kb + fi + (size_t)fj * Nrad;
            if (iel == jel) {
              const double *Lf = chol[L * Nel + iel];
              const int nn = Ni * Ni;
              for (int p = 0; p < rank[L * Nel + iel]; p++) {
                const double *Mp = Lf + (size_t)p * nn;
                memset(t1, 0, sizeof(double) * nn);
                gemm_nt(Ni, Ni, Ni, 1.0, ps, Ni, Mp, Ni, t1, Ni);       /* P Mp' */
                gemm_nn(Ni, Ni, Ni, -1.0, Mp, Ni, t1, Ni, Kdst, Nrad);   /* K -= Mp (P Mp') */
              }
            } else {
              const double *iint = (iel > jel) ? dbig[L * Nel + iel] : dsmall[L * Nel + iel];
              const double *jint = (iel > jel) ? dsmall[L * Nel + jel] : dbig[L * Nel + jel];
              memset(t1, 0, sizeof(double) * Ni * Nj);
              gemm_nt(Ni, Nj, Nj, 1.0, ps, Ni, jint, Nj, t1, Ni);
              gemm_nn(Ni, Nj, Ni, -1.0, iint, Ni, t1, Ni, Kdst, Nrad);
            }
          }
      }
    }
    free(R); free(couple); free(t1); free(ps);
  }
  free(norms);
}
