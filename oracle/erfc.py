"""Radial Green's function of the erfc-attenuated Coulomb kernel (oracle, test infrastructure only).

    erfc(mu r12)/r12 = sum_L (4 pi mu/(2L+1)) Phi_L(mu r, mu r') sum_M Y_LM Y*_LM

Restates libhelfem/src/erfc_expn.cpp:27-235 (J. G. Angyan, I. Gerber, M. Marsman, J. Phys. A 39, 8613
(2006)) vectorised over numpy arrays: `Phi` (:225-235) switches between the closed form `Phi_general`
(:94-116: F_n, H_n of :66-92) and the Taylor series in the smaller argument `Phi_short` (:148-186, coefficients
D_nk of :118-146).  Pinned against the reference's own erfc_expn.cpp compiled into oracle/_ref/liberfc_ref.so
(tests/test_oracle.py).
"""
import math

import numpy as np

_SQRTPI = math.sqrt(math.pi)


def _dfact(n):
    """n!! (erfc_expn.cpp:27-33)."""
    v = 1.0
    while n >= 2:
        v *= n
        n -= 2
    return v


def _choose(n, m):
    """Generalised binomial coefficient, negative upper index allowed (erfc_expn.cpp:42-64)."""
    if n == -1:
        return (-1.0) ** m
    if n == 0:
        return 1.0 if m == 0 else 0.0
    if m == 0:
        return 1.0
    if m == 1:
        return float(n)
    if n > 0 and m > n:
        return 0.0
    if n < 0:
        return _choose(n + m - 1, m) * (-1.0) ** m
    v = 1.0
    for i in range(min(m, n - m)):
        v *= (n - i) / (i + 1)
    return v


def _Fn(n, Xi, xi):
    """erfc_expn.cpp:66-77."""
    explus = np.exp(-(Xi + xi) ** 2)
    exminus = np.exp(-(Xi - xi) ** 2)
    prefac = -1.0 / (4.0 * Xi * xi)
    F = np.zeros_like(Xi)
    for p in range(n + 1):
        c = math.factorial(n + p) / (math.factorial(p) * math.factorial(n - p))
        F = F + prefac ** (p + 1) * c * ((-1.0) ** (n - p) * explus - exminus)
    return 2.0 / _SQRTPI * F


def _Hn(n, Xi, xi):
    """erfc_expn.cpp:79-92."""
    from scipy.special import erfc
    X = Xi ** (2 * n + 1)
    x = xi ** (2 * n + 1)
    H = (X + x) * erfc(Xi + xi) - (X - x) * erfc(Xi - xi)
    return H / (2.0 * (xi * Xi) ** (n + 1))


def _phi_general(n, Xi, xi):
    """erfc_expn.cpp:94-116 (Xi >= xi on entry)."""
    Farr = [_Fn(i, Xi, xi) for i in range(n)]
    s = np.zeros_like(Xi)
    for m in range(1, n + 1):
        Xm, xm = Xi ** m, xi ** m
        s = s + Farr[n - m] * ((Xm * Xm + xm * xm) / (Xm * xm))
    return _Fn(n, Xi, xi) + s + _Hn(n, Xi, xi)


def _Dnk(n, k, Xi):
    """erfc_expn.cpp:118-146."""
    from scipy.special import erfc
    prefac = np.exp(-Xi ** 2) / _SQRTPI * 2.0 ** (n + 1) * Xi ** (2 * n + 1)
    if k == 0:
        s = np.zeros_like(Xi)
        for m in range(1, n + 1):
            s = s + 1.0 / (_dfact(2 * (n - m) + 1) * (2.0 * Xi * Xi) ** m)
        return erfc(Xi) + prefac * s
    s = np.zeros_like(Xi)
    for m in range(1, k + 1):
        s = s + _choose(m - k - 1, m - 1) * (2.0 * Xi * Xi) ** (k - m) / _dfact(2 * (n + k - m) + 1)
    return prefac * (2.0 * n + 1.0) / (math.factorial(k) * (2.0 * (n + k) + 1.0)) * s


def _phi_short(n, Xi, xi):
    """erfc_expn.cpp:148-186 (Xi >= xi on entry): pairs of Taylor terms until |dPhi| < eps max(|Phi|, 1)."""
    out = np.zeros_like(Xi)
    zero_small = xi == 0.0
    if n == 0:
        both = zero_small & (Xi == 0.0)
        out[both] = 1.0
        todo = ~both
    else:
        todo = ~zero_small
    idx = np.nonzero(todo)[0]
    if idx.size == 0:
        return out
    X, x = Xi[idx], xi[idx]
    Phi = np.zeros_like(X)
    active = np.ones(X.shape, dtype=bool)
    eps = np.finfo(float).eps
    for k in range(0, 202, 2):
        a = np.nonzero(active)[0]
        if a.size == 0:
            break
        Xa, xa = X[a], x[a]
        d = _Dnk(n, k, Xa) * xa ** (n + 2 * k) + _Dnk(n, k + 1, Xa) * xa ** (n + 2 * (k + 1))
        Phi[a] += d
        active[a[np.abs(d) < eps * np.maximum(np.abs(Phi[a]), 1.0)]] = False
    if active.any():
        raise RuntimeError("Phi_short Taylor series failed to converge")
    out[idx] = Phi / X ** (n + 1)
    return out


def Phi(n, Xi, xi):
    """erfc_expn.cpp:225-235, elementwise over broadcastable arrays."""
    Xi, xi = np.broadcast_arrays(np.asarray(Xi, dtype=float), np.asarray(xi, dtype=float))
    shape = Xi.shape
    big = np.maximum(Xi, xi).ravel()
    small = np.minimum(Xi, xi).ravel()
    out = np.empty_like(big)
    short = (small < 0.4) | ((big < 0.5) & (small < 2.0 * big))
    if short.any():
        out[short] = _phi_short(n, big[short], small[short])
    if (~short).any():
        out[~short] = _phi_general(n, big[~short], small[~short])
    return out.reshape(shape)
