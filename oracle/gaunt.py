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
