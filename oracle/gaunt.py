"""Gaunt coefficients for the oracle.  Test infrastructure only.

The reference takes its values from the third-party library wignernj (pinned
v0.8.0, CMakeLists.txt:288-294; call sites src/general/gaunt.cpp:48-57,83,172),
which is not in the tree.  wignernj evaluates the coupling coefficients from
exact prime-factorised rationals; here the published Racah formula for the 3j
symbol is evaluated in exact integer arithmetic (``gaunt_exact``) and the only
rounding is the final square root (50 digits, then to double).  The result is
pinned against the reference's own table src/general/gaunt_test.cpp
(tests/golden/gaunt_ref.json).

Conventions follow src/general/gaunt.cpp:77-84, :202-272:
  coeff(L,M,l,m,lp)           = int Y_L^M* Y_l^m Y_lp^(M-m) dOmega
                              = (-1)^M gaunt(L,-M,l,m,lp,M-m)
  mod_coeff(lj,mj,L,M,li,mi)  = int Y_lj^mj* cos^2(theta) Y_L^M Y_li^mi dOmega
"""
from fractions import Fraction
from functools import lru_cache
from math import factorial, pi, sqrt

import mpmath

mpmath.mp.dps = 50


@lru_cache(maxsize=None)
def _threej_parts(j1, j2, j3, m1, m2, m3):
    """3j symbol as (S, A): value = S*sqrt(A), S and A exact rationals."""
    if m1 + m2 + m3 != 0:
        return Fraction(0), Fraction(0)
    if j3 < abs(j1 - j2) or j3 > j1 + j2:
        return Fraction(0), Fraction(0)
    if abs(m1) > j1 or abs(m2) > j2 or abs(m3) > j3:
        return Fraction(0), Fraction(0)
    f = factorial
    delta = Fraction(f(j1 + j2 - j3) * f(j1 - j2 + j3) * f(-j1 + j2 + j3), f(j1 + j2 + j3 + 1))
    A = delta * f(j1 + m1) * f(j1 - m1) * f(j2 + m2) * f(j2 - m2) * f(j3 + m3) * f(j3 - m3)
    kmin = max(0, j2 - j3 - m1, j1 - j3 + m2)
    kmax = min(j1 + j2 - j3, j1 - m1, j2 + m2)
    S = Fraction(0)
    for k in range(kmin, kmax + 1):
        den = (f(k) * f(j1 + j2 - j3 - k) * f(j1 - m1 - k) * f(j2 + m2 - k)
               * f(j3 - j2 + m1 + k) * f(j3 - j1 - m2 + k))
        S += Fraction((-1) ** k, den)
    if (j1 - j2 - m3) % 2:
        S = -S
    return S, A


@lru_cache(maxsize=None)
def gaunt_exact(l1, m1, l2, m2, l3, m3):
    """int Y_l1^m1 Y_l2^m2 Y_l3^m3 dOmega (no conjugation)."""
    if m1 + m2 + m3 != 0 or (l1 + l2 + l3) % 2:
        return 0.0
    S0, A0 = _threej_parts(l1, l2, l3, 0, 0, 0)
    S1, A1 = _threej_parts(l1, l2, l3, m1, m2, m3)
    if S0 == 0 or S1 == 0:
        return 0.0
    rat = A0 * A1 * (2 * l1 + 1) * (2 * l2 + 1) * (2 * l3 + 1)
    val = (mpmath.mpf(S0.numerator) / S0.denominator) * (mpmath.mpf(S1.numerator) / S1.denominator) \
        * mpmath.sqrt(mpmath.mpf(rat.numerator) / rat.denominator / (4 * mpmath.pi))
    return float(val)


def gaunt_coefficient(L, M, l, m, lp, mp):
    """src/general/gaunt.cpp:77-84."""
    sign = -1.0 if (M & 1) else 1.0
    return sign * gaunt_exact(L, -M, l, m, lp, mp)


def modified_gaunt_coefficient(lj, mj, L, M, li, mi):
    """Free-function form, src/general/gaunt.cpp:90-103."""
    const0 = 2.0 / 3.0 * sqrt(pi)
    const2 = 4.0 / 15.0 * sqrt(5.0 * pi)
    cpl0 = gaunt_coefficient(L, M, 0, 0, L, M) * gaunt_coefficient(lj, mj, li, mi, L, M)
    cpl2 = 0.0
    for Lp in range(max(max(L - 2, 0), abs(M)), L + 3):
        cpl2 += gaunt_coefficient(Lp, M, 2, 0, L, M) * gaunt_coefficient(lj, mj, li, mi, Lp, M)
    return const0 * cpl0 + const2 * cpl2


class Gaunt:
    """Lookup object with the reference's GauntT interface
    (src/general/gaunt.cpp:202-290).  No table limits: values are computed on
    demand and memoised."""

    def coeff(self, L, M, l, m, lp):
        if L < 0 or l < 0 or lp < 0:
            return 0.0
        if abs(M) > L or abs(m) > l:
            return 0.0
        mp = M - m
        if abs(mp) > lp:
            return 0.0
        return gaunt_coefficient(L, M, l, m, lp, mp)

    def mod_coeff(self, lj, mj, L, M, li, mi):
        if mj != M + mi:
            return 0.0
        const0 = 2.0 / 3.0 * sqrt(pi)
        const2 = 4.0 / 15.0 * sqrt(5.0 * pi)
        cpl0 = self.coeff(L, M, 0, 0, L) * self.coeff(lj, mj, li, mi, L)
        cpl2 = 0.0
        for Lp in range(max(max(L - 2, 0), abs(M)), L + 3):
            cpl2 += self.coeff(Lp, M, 2, 0, L) * self.coeff(lj, mj, li, mi, Lp)
        return const0 * cpl0 + const2 * cpl2

    def cosine_coupling(self, lj, mj, li, mi):
        if mi != mj:
            return 0.0
        return 2.0 * sqrt(pi / 3.0) * self.coeff(lj, mj, 1, 0, li)

    def cosine2_coupling(self, lj, mj, li, mi):
        if mi != mj:
            return 0.0
        const0 = 2.0 / 3.0 * sqrt(pi)
        const2 = 4.0 / 15.0 * sqrt(5.0 * pi)
        return const0 * self.coeff(lj, mj, 0, 0, li) + const2 * self.coeff(lj, mj, 2, 0, li)


# ---------------------------------------------------------------------------
# Bulk tables by quadrature (numpy), for sizes where the exact evaluation is too
# slow (N2 lmax=30: ~10^7 coefficients).  Validated against gaunt_exact in
# tests/test_oracle.py.  int Y_l1^m1 Y_l2^m2 Y_l3^m3 = (2 pi)^(-1/2) int Theta Theta Theta dx
# with fully normalised Theta_l^m (Condon-Shortley phase), Gauss-Legendre exact
# for the polynomial integrand.
# ---------------------------------------------------------------------------
import numpy as _np


def _theta_table(lmax, m, x):
    """Theta_l^m(x) for l = 0..lmax (zero for l < |m|), shape (lmax+1, len(x)), long double."""
    ma = abs(m)
    x = x.astype(_np.longdouble)
    s = _np.sqrt((1 - x) * (1 + x))
    out = _np.zeros((lmax + 1, len(x)), dtype=_np.longdouble)
    if ma > lmax:
        return out
    pmm = _np.full(len(x), _np.sqrt(_np.longdouble(0.5)))
    for k in range(1, ma + 1):
        pmm = -pmm * s * _np.sqrt(_np.longdouble(2 * k + 1) / (2 * k))
    out[ma] = pmm
    pl2 = _np.zeros_like(pmm)
    pl1 = pmm
    for l in range(ma + 1, lmax + 1):
        a = _np.sqrt(_np.longdouble(4 * l * l - 1) / (l * l - ma * ma))
        b = _np.sqrt(_np.longdouble((l - 1) ** 2 - ma * ma) / (4 * (l - 1) ** 2 - 1))
        pl = a * (x * pl1 - b * pl2)
        out[l] = pl
        pl2, pl1 = pl1, pl
    if m < 0 and (ma & 1):
        out = -out
    return out


def coupling_tables(lval, mval, NL, diatomic):
    """Dense tables for the C oracle: g2[(j,i,L)] = coeff(lj,mj,L,mj-mi,li) and (diatomic)
    g0[(j,i,L)] = mod_coeff(lj,mj,L,mj-mi,li,mi), L = 0..NL-1."""
    lval = _np.asarray(lval); mval = _np.asarray(mval)
    na = len(lval)
    lmax = int(max(lval.max(), NL + 2)) + 1
    nq = (3 * lmax) // 2 + 2
    xq, wq = _np.polynomial.legendre.leggauss(nq)
    # refine nodes/weights to long double with two Newton steps
    xl = xq.astype(_np.longdouble)
    for _ in range(3):
        p0 = _np.ones_like(xl); p1 = xl.copy()
        for k in range(2, nq + 1):
            p0, p1 = p1, ((2 * k - 1) * xl * p1 - (k - 1) * p0) / k
        dp = nq * (xl * p1 - p0) / (xl * xl - 1)
        xl = xl - p1 / dp
    p0 = _np.ones_like(xl); p1 = xl.copy()
    for k in range(2, nq + 1):
        p0, p1 = p1, ((2 * k - 1) * xl * p1 - (k - 1) * p0) / k
    dp = nq * (xl * p1 - p0) / (xl * xl - 1)
    wl = 2 / ((1 - xl * xl) * dp * dp)
    inv = 1 / _np.sqrt(2 * _np.longdouble(_np.pi))
    Lcap = NL + 2
    ms = sorted(set(int(m) for m in mval))
    th = {m: _theta_table(lmax, m, xl) for m in range(-2 * max(map(abs, ms)) - 1, 2 * max(map(abs, ms)) + 2)}
    g2 = _np.zeros((na, na, NL)); g0 = _np.zeros((na, na, NL)) if diatomic else None
    c0 = 2.0 / 3.0 * _np.sqrt(_np.pi); c2 = 4.0 / 15.0 * _np.sqrt(5.0 * _np.pi)
    for mj in ms:
        jj = _np.where(mval == mj)[0]
        for mi in ms:
            ii = _np.where(mval == mi)[0]
            M = mj - mi
            # coeff(lj,mj,L,M,li) = (-1)^mj gaunt3(lj,-mj; L,M; li,mi)   for L = 0..Lcap
            A = th[-mj][lval[jj]] * wl                      # (nj, nq)
            Bt = th[M][:Lcap + 1]                           # (Lcap+1, nq)
            C = th[mi][lval[ii]]                            # (ni, nq)
            full = _np.einsum("jq,Lq,iq->jiL", A, Bt, C) * inv * (-1.0 if (mj & 1) else 1.0)
            # selection rules: exact zeros
            for a, lj in enumerate(lval[jj]):
                for b, li in enumerate(lval[ii]):
                    for L in range(Lcap + 1):
                        if (lj + li + L) % 2 or L < abs(lj - li) or L > lj + li or L < abs(M):
                            full[a, b, L] = 0.0
            full = full.astype(float)
            g2[_np.ix_(jj, ii)] = full[:, :, :NL]
            if diatomic:
                # cos^2 Y_L^M = sum_Lp t(Lp,L) Y_Lp^M,  t = c0 d(Lp,L) coeff(L,M,0,0,L) + c2 coeff(Lp,M,2,0,L)
                # coeff(Lp,M,2,0,L) = (-1)^M gaunt3(Lp,-M; 2,0; L,M)
                tm = _np.einsum("pq,q,Lq->pL", th[-M][:Lcap + 1] * wl, th[0][2], th[M][:Lcap + 1]) * inv * (-1.0 if (M & 1) else 1.0)
                t0 = _np.einsum("pq,q,Lq->pL", th[-M][:Lcap + 1] * wl, th[0][0], th[M][:Lcap + 1]) * inv * (-1.0 if (M & 1) else 1.0)
                tm = tm.astype(float); t0 = t0.astype(float)
                mod = _np.zeros((len(jj), len(ii), NL))
                for L in range(abs(M), NL):
                    acc = c0 * t0[L, L] * full[:, :, L]
                    for Lp in range(max(max(L - 2, 0), abs(M)), min(L + 2, Lcap) + 1):
                        if (Lp + L) % 2 == 0:
                            acc = acc + c2 * tm[Lp, L] * full[:, :, Lp]
                    mod[:, :, L] = acc
                g0[_np.ix_(jj, ii)] = mod
    return g0, g2
