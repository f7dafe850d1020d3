"""Finite-element primitives: quadrature rules, element grids, LIP basis.

Oracle restatement (numpy) of the reference's FEM core.  Test infrastructure
only -- see oracle/__init__.py.
"""
import numpy as np


def lobatto(n):
    """n-point Gauss-Lobatto rule on [-1,1], ascending nodes.

    Follows libhelfem/include/lobatto.h:36-109 (Newton on P'_{n-1} from the
    Chebyshev-Gauss-Lobatto guess, w_i = 2/(n(n-1)P_{n-1}(x_i)^2)).
    """
    if n < 2:
        raise ValueError("Lobatto needs n>=2")
    x = np.cos(np.pi * np.arange(n) / (n - 1))
    tol = 100 * np.finfo(float).eps
    while True:
        xold = x.copy()
        p = np.zeros((n, n))
        p[:, 0] = 1.0
        p[:, 1] = x
        for j in range(2, n):
            p[:, j] = ((2 * j - 1) * x * p[:, j - 1] - (j - 1) * p[:, j - 2]) / j
        x = xold - (x * p[:, n - 1] - p[:, n - 2]) / (n * p[:, n - 1])
        if np.max(np.abs(x - xold)) <= tol:
            break
    x = x[::-1].copy()
    pn = np.ones(n)
    pnm1 = np.ones(n)
    pn = x.copy()
    for j in range(2, n):
        pnm2 = pnm1
        pnm1 = pn
        pn = ((2 * j - 1) * x * pnm1 - (j - 1) * pnm2) / j
    w = 2.0 / ((n - 1) * n * pn * pn)
    return x, w


def chebyshev(n):
    """Modified Gauss-Chebyshev rule of the second kind on [-1,1].

    Follows libhelfem/include/chebyshev.h:32-57.
    """
    i = np.arange(1, n + 1)
    ang = i * np.pi / (n + 1)
    s = np.sin(ang)
    c = np.cos(ang)
    w = 16.0 / 3.0 / (n + 1) * s ** 4
    x = 1.0 - 2.0 * i / (n + 1) + 2.0 / np.pi * (1.0 + 2.0 / 3.0 * s * s) * c * s
    return x[::-1].copy(), w[::-1].copy()


def get_grid(rmax, num_el, igrid, zexp):
    """Element boundaries; libhelfem/include/grid.h:38-103."""
    if igrid == 1:
        b = np.linspace(0.0, rmax, num_el + 1)
    elif igrid == 2:
        i = np.arange(num_el + 1)
        b = i * i * rmax / (num_el * num_el)
    elif igrid == 3:
        b = rmax * (np.arange(num_el + 1) / num_el) ** zexp
    elif igrid == 4:
        upper = np.log(rmax + 1.0) ** (1.0 / zexp)
        t = np.linspace(0.0, upper, num_el + 1)
        b = np.exp(t ** zexp) - 1.0
    else:
        raise ValueError("grid type not restated in the oracle")
    b[0] = 0.0
    b[-1] = rmax
    return b


def lip_eval(x, x0, n):
    """d^n L_i/dx^n of the Lagrange polynomials on nodes x0 at points x.

    Returns (len(x), len(x0)).  libhelfem/include/LIPBasis_eval.h:29-88.
    """
    x = np.asarray(x, dtype=float)
    N = len(x0)
    out = np.zeros((len(x), N))
    for fi in range(N):
        others = [ip for ip in range(N) if ip != fi]
        if n == 0:
            v = np.ones(len(x))
            for ip in others:
                v = v * (x - x0[ip]) / (x0[fi] - x0[ip])
            out[:, fi] = v
        elif n == 1:
            el = np.zeros(len(x))
            for d1 in others:
                v = np.ones(len(x))
                for ip in others:
                    if ip == d1:
                        continue
                    v = v * (x - x0[ip]) / (x0[fi] - x0[ip])
                el += v / (x0[fi] - x0[d1])
            out[:, fi] = el
        elif n == 2:
            el = np.zeros(len(x))
            for d1 in others:
                for d2 in others:
                    if d2 >= d1:
                        continue
                    v = np.ones(len(x))
                    for ip in others:
                        if ip == d1 or ip == d2:
                            continue
                        v = v * (x - x0[ip]) / (x0[fi] - x0[ip])
                    el += v / ((x0[fi] - x0[d1]) * (x0[fi] - x0[d2]))
            out[:, fi] = 2.0 * el
        else:
            raise ValueError("derivative order not restated")
    return out


class FEBasis:
    """1D finite-element basis of n-node LIPs (primbas=4), noverlap = 1.

    libhelfem/src/FiniteElementBasis.cpp:52-64 (function numbering),
    :297-313 (boundary drops), :253-260 (coordinates), :341-346 (derivative
    scaling), libhelfem/include/LIPBasis.h:32-116.
    """

    def __init__(self, nnodes, bval, zero_func_left, zero_func_right):
        self.x0, _ = lobatto(nnodes)
        self.nnodes = nnodes
        self.bval = np.asarray(bval, dtype=float)
        self.nel = len(bval) - 1
        self.zl = zero_func_left
        self.zr = zero_func_right
        self.first = np.zeros(self.nel, dtype=int)
        self.last = np.zeros(self.nel, dtype=int)
        for iel in range(self.nel):
            self.first[iel] = 0 if iel == 0 else self.last[iel - 1]
            self.last[iel] = self.first[iel] + len(self.enabled(iel)) - 1
        self.nbf = int(self.last[-1] + 1)

    def enabled(self, iel):
        en = list(range(self.nnodes))
        if iel == 0 and self.zl:
            en = en[1:]
        if iel == self.nel - 1 and self.zr:
            en = en[:-1]
        return en

    def nprim(self, iel):
        return len(self.enabled(iel))

    def idx(self, iel):
        return int(self.first[iel]), int(self.last[iel])

    def begin(self, iel):
        return self.bval[iel]

    def end(self, iel):
        return self.bval[iel + 1]

    def mid(self, iel):
        return 0.5 * (self.bval[iel + 1] + self.bval[iel])

    def scale(self, iel):
        return 0.5 * (self.bval[iel + 1] - self.bval[iel])

    def coord(self, x, iel):
        return self.mid(iel) + self.scale(iel) * np.asarray(x)

    def eval_dnf(self, x, n, iel):
        prim = lip_eval(x, self.x0, n)
        return prim[:, self.enabled(iel)] / self.scale(iel) ** n

    def eval_over_r(self, x, n, iel):
        """B(r)/r and derivatives on the first element by analytic deflation
        of the (x+1) factor; libhelfem/include/LIPBasis.h:87-112."""
        if abs(self.begin(iel)) > 1e-14:
            raise ValueError("eval_over_r only valid on the element at r=0")
        en = self.enabled(iel)
        if en[0] == 0:
            raise ValueError("eval_over_r needs the first function dropped")
        red = lip_eval(x, self.x0[1:], n)
        sc = 1.0 / self.scale(iel) ** (n + 1)
        out = np.zeros((len(x), len(en)))
        for k, ifull in enumerate(en):
            out[:, k] = sc * red[:, ifull - 1] / (self.x0[ifull] + 1.0)
        return out

    def matrix_element(self, iel, lh, rh, xq, wq, f=None, x_left=-1.0, x_right=1.0):
        """sum_q w_q f(r_q) lh(q,i) rh(q,j) on a sub-panel of element iel.

        libhelfem/src/FiniteElementBasis.cpp:441-479.
        """
        a = 0.5 * (x_right - x_left)
        b = 0.5 * (x_right + x_left)
        xs = a * xq + b
        r = self.coord(xs, iel)
        wp = wq * self.scale(iel) * a
        if f is not None:
            with np.errstate(all="ignore"):
                fv = f(r)
            wp = np.where(np.isfinite(fv), wp * fv, 0.0)
        L = lh(xs, iel)
        R = rh(xs, iel)
        return (L * wp[:, None]).T @ R

    def matrix_element_auto(self, iel, lh, rh, f=None, x_left=-1.0, x_right=1.0, poly_degree_f=-1):
        """Order-doubling Gauss-Lobatto refinement to 8 eps;
        libhelfem/src/FiniteElementBasis.cpp:500-604."""
        basisdeg = max(0, self.nnodes - 1)
        deg = 2 * basisdeg + max(0, poly_degree_f)
        nstart = max(5, (deg + 4) // 2 + 2)
        return converge(lambda n: self.matrix_element(iel, lh, rh, *lobatto(n), f, x_left, x_right),
                        nstart, 512)


def converge(probe, nstart, nmax, floor_rel=None, seed_fallback=False, want_n=False):
    """Order-doubling convergence loop shared by the reference's integrators.

    floor_rel=None: libhelfem/src/RadialBasis.cpp:114-162 and
    FiniteElementBasis.cpp:500-555 (three exits).  floor_rel=256 eps:
    src/diatomic/basis.cpp:125-200 (two exits + seed fallback).
    """
    eps = np.finfo(float).eps
    tol = 8 * eps
    sqrteps = np.sqrt(eps)
    prev = None
    seed = None
    prevdiff = -1.0
    prevprevdiff = -1.0
    n = max(nstart, 2)
    while True:
        cur = probe(n)
        if prev is None:
            seed = cur
        else:
            diff = np.max(np.abs(cur - prev))
            scale = np.max(np.abs(cur))
            if diff <= tol * (scale + tol):
                return (cur, n) if want_n else cur
            if floor_rel is None:
                if prevdiff >= 0 and diff <= sqrteps * (scale + tol) and diff > 0.5 * prevdiff:
                    return (cur, n) if want_n else cur
                if prevprevdiff >= 0 and diff <= sqrteps * (scale + tol) and diff > 0.125 * prevprevdiff:
                    return (cur, n) if want_n else cur
                prevprevdiff = prevdiff
            else:
                if diff <= sqrteps * (scale + tol) and (
                        diff <= floor_rel * (scale + tol) or (prevdiff >= 0 and diff > 0.5 * prevdiff)):
                    return (cur, n) if want_n else cur
            prevdiff = diff
        prev = cur
        if n >= nmax:
            if seed_fallback:
                return (seed, max(nstart, 2)) if want_n else seed
            return (cur, n) if want_n else cur
        n = min(2 * n, nmax)
