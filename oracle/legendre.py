"""Associated Legendre functions P_l^m(x), Q_l^m(x) for x >= 1 (Hobson form).

Oracle restatement, vectorised over x, of src/legendre/Legendre.h:92-378
(seed + forward l recurrence for P; Christoffel decomposition below
cosh(ln(1e3)/(2 lmax)), Miller's downward recurrence with a Lentz-Thompson
continued fraction above it, upward m recurrence for Q).  Validated against
oracle/_ref (the reference header compiled as-is).  Test infrastructure only.
"""
import numpy as np


def plm(lmax, m, x):
    """P_l^m(x), l=0..lmax, for one m.  Returns (lmax+1, len(x)).
    Legendre.h:92-132 with s=-1 (x>1): P_m^m=(2m-1) w P_(m-1)^(m-1)."""
    x = np.atleast_1d(np.asarray(x, dtype=float))
    P = np.zeros((lmax + 1, len(x)))
    if m > lmax:
        return P
    s = np.where(np.abs(x) <= 1.0, 1.0, -1.0)
    w = np.sqrt(np.maximum(s * (1.0 - x * x), 0.0))
    pmm = np.ones(len(x))
    for mm in range(1, m + 1):
        pmm = -s * (2 * mm - 1) * w * pmm
    P[m] = pmm
    if m + 1 <= lmax:
        P[m + 1] = (2 * m + 1) * x * pmm
    for l in range(m + 1, lmax):
        P[l + 1] = ((2 * l + 1) * x * P[l] - (l + m) * P[l - 1]) / (l + 1 - m)
    return P


def _q00(x):
    return 0.5 * (np.log(np.abs(x + 1.0)) - np.log(np.abs(x - 1.0)))


def _cf_ratio(x, lmax, m):
    """Lentz-Thompson continued fraction for Q_lmax^m / Q_(lmax-1)^m;
    Legendre.h:156-190."""
    tiny = np.finfo(float).tiny * 1e4
    tol = 8 * np.finfo(float).eps
    f = np.full(len(x), tiny)
    C = f.copy()
    D = np.zeros(len(x))
    a = 1.0
    n = lmax
    active = np.ones(len(x), dtype=bool)
    for _ in range(1000000):
        b = (2 * n + 1) * x / (n + m)
        Dn = b + a * D
        Dn[Dn == 0] = tiny
        Cn = b + a / C
        Cn[Cn == 0] = tiny
        Dn = 1.0 / Dn
        delta = Cn * Dn
        D = np.where(active, Dn, D)
        C = np.where(active, Cn, C)
        f = np.where(active, f * delta, f)
        active = active & ~(np.abs(delta - 1.0) < tol)
        if not active.any():
            return f
        a = -(n - m + 1) / (n + m)
        n += 1
    raise RuntimeError("continued fraction did not converge")


def _m_up(Q0, Q1, lmax, M, x):
    """Upward m recurrence Q_l^(m+1) = -2m x/w Q_l^m - s (l+m)(l-m+1) Q_l^(m-1);
    Legendre.h:218-238.  Returns the column m=M."""
    if M == 0:
        return Q0
    if M == 1:
        return Q1
    s = -1.0
    w = np.sqrt(np.maximum(s * (1.0 - x * x), 0.0))
    l = np.arange(lmax + 1)[:, None]
    prev, cur = Q0, Q1
    for m in range(1, M):
        nxt = -(2 * m) * x / w * cur - s * (l + m) * (l - m + 1) * prev
        prev, cur = cur, nxt
    return cur


def qlm(lmax, M, x):
    """Q_l^M(x), l=0..lmax, x>1.  Returns (lmax+1, len(x)).  Legendre.h:335-378."""
    x = np.atleast_1d(np.asarray(x, dtype=float))
    if np.any(x <= 1.0):
        raise ValueError("oracle qlm restated for x>1 only")
    n = len(x)
    Q0 = np.zeros((lmax + 1, n))
    Q1 = np.zeros((lmax + 1, n))
    xmax = np.cosh(np.log(1e3) / (2 * lmax)) if lmax > 0 else np.inf
    chris = np.abs(x) < xmax
    q00 = _q00(x)
    w = np.sqrt(np.maximum(x * x - 1.0, 0.0))
    # seeds (Legendre.h:134-154); both paths overwrite what they own
    Q0[0] = q00
    if lmax >= 1:
        Q0[1] = x * q00 - 1.0
    Q1[0] = -1.0 / w
    if lmax >= 1:
        Q1[1] = w * (q00 + x / (1.0 - x * x))
    if chris.any():
        # Christoffel path, Legendre.h:252-300
        xc = x[chris]
        nc = len(xc)
        P = np.zeros((lmax + 1, nc)); W = np.zeros((lmax + 1, nc))
        Pp = np.zeros((lmax + 1, nc)); Wp = np.zeros((lmax + 1, nc))
        P[0] = 1.0
        if lmax >= 1:
            P[1] = xc; W[1] = 1.0; Pp[1] = 1.0
        for k in range(1, lmax):
            inv = 1.0 / (k + 1)
            a = 2 * k + 1
            P[k + 1] = (a * xc * P[k] - k * P[k - 1]) * inv
            W[k + 1] = (a * xc * W[k] - k * W[k - 1]) * inv
            Pp[k + 1] = (a * (P[k] + xc * Pp[k]) - k * Pp[k - 1]) * inv
            Wp[k + 1] = (a * (W[k] + xc * Wp[k]) - k * Wp[k - 1]) * inv
        qc = q00[chris]; wc = w[chris]
        Q0[:, chris] = P * qc - W
        q1 = wc * Pp * qc - P / wc - wc * Wp
        q1[0] = -1.0 / wc
        Q1[:, chris] = q1
    mil = ~chris
    if mil.any() and lmax >= 1:
        # Miller path, Legendre.h:304-330
        xm = x[mil]
        low = np.finfo(float).tiny * 1e4
        for m, Qc, norm in ((0, Q0, q00[mil]), (1, Q1, -1.0 / w[mil])):
            ratio = _cf_ratio(xm, lmax, m)
            col = np.zeros((lmax + 1, len(xm)))
            col[lmax - 1] = low
            col[lmax] = low * ratio
            for l in range(lmax - 1, 0, -1):
                col[l - 1] = ((2 * l + 1) * xm * col[l] - (l + 1 - m) * col[l + 1]) / (l + m)
            col *= norm / col[0]
            Qc[:, mil] = col
    return _m_up(Q0, Q1, lmax, M, x)


def filter_normal(v):
    """Zero out subnormal / non-finite entries; src/diatomic/quadrature.h:63-66."""
    v = np.array(v, dtype=float)
    bad = (v != 0.0) & ~(np.isfinite(v) & (np.abs(v) >= np.finfo(float).tiny))
    v[bad] = 0.0
    return v
