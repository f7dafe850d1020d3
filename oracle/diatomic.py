"""Diatomic prolate-spheroidal (mu,nu,phi) basis: setup, J and K.

Oracle restatement (numpy); test infrastructure only.  Follows
src/diatomic/basis.cpp (RadialBasis :258-408, lm_to_l_m :505-520, ctor
:525-647, pure_indices :718-737, 1e matrices :1032-1166, compute_tei
:1334-1547, LMfac :1549-1560, coulomb :1627-1816, exchange :1818-2089,
expand/remove_boundaries :2091-2118) and src/diatomic/quadrature.{h,cpp}
(:47-84, :133-257).
"""
import numpy as np

from . import fem, legendre
from .gaunt import Gaunt

TWOE_NMAX = 512


def lm_to_l_m(lmax_per_m):
    """basis.cpp:505-520."""
    lv, mv = [], []
    for mabs, lm in enumerate(lmax_per_m):
        for l in range(mabs, lm + 1):
            lv.append(l); mv.append(mabs)
            if mabs > 0:
                lv.append(l); mv.append(-mabs)
    return np.array(lv), np.array(mv)


class LegCache:
    """quadrature.h:47-84: one |M|, all L up to Lmax, values filtered to
    normal numbers; Q reports 0 at cosh(mu)==1."""

    def __init__(self, Mabs, Lmax):
        self.M, self.Lmax = Mabs, Lmax

    def P(self, chmu):
        return legendre.filter_normal(legendre.plm(self.Lmax, self.M, chmu))

    def Q(self, chmu):
        chmu = np.asarray(chmu, dtype=float)
        out = np.zeros((self.Lmax + 1, len(chmu)))
        ok = chmu != 1.0
        if ok.any():
            out[:, ok] = legendre.filter_normal(legendre.qlm(self.Lmax, self.M, chmu[ok]))
        return out


class TwoeElement:
    """quadrature.cpp:133-185: everything in the in-element integrals that
    depends only on the element and the rule."""

    def __init__(self, febasis, iel, n):
        x, wx = fem.chebyshev(n)
        en = febasis.enabled(iel)
        mumin, mumax = febasis.begin(iel), febasis.end(iel)
        mumid, self.mulen = 0.5 * (mumax + mumin), 0.5 * (mumax - mumin)
        self.x, self.wx = x, wx
        self.mu = mumid + self.mulen * x
        self.chmu, self.shmu = np.cosh(self.mu), np.sinh(self.mu)
        self.bf = fem.lip_eval(x, febasis.x0, 0)[:, en]
        nbf = self.bf.shape[1]
        self.bfprod = (self.bf[:, :, None] * self.bf[:, None, :]).reshape(n, nbf * nbf)
        self.sub = []
        for ip in range(n):
            submin = mumin if ip == 0 else self.mu[ip - 1]
            submax = self.mu[ip]
            submid, sublen = 0.5 * (submax + submin), 0.5 * (submax - submin)
            submu = submid + sublen * x
            xpoly = (submu - mumid) / self.mulen
            self.sub.append((sublen, np.cosh(submu), np.sinh(submu),
                             fem.lip_eval(xpoly, febasis.x0, 0)[:, en]))
        # all subinterval points at once, for the Legendre table
        self.sub_chmu = np.concatenate([s[1] for s in self.sub])


def twoe_integral(el, k, l, L, Pall, Qouter):
    """quadrature.cpp:188-257.  Pall: P_L at el.sub_chmu (flattened, nq*nq),
    Qouter: Q_L at el.chmu."""
    nq = len(el.x)
    nbf = el.bf.shape[1]

    def inner_integral(lval):
        inner = np.zeros((nq, nbf * nbf))
        acc = np.zeros(nbf * nbf)
        for ip in range(nq):
            sublen, ch, sh, bf = el.sub[ip]
            wp = sublen * el.wx * sh
            if lval != 0:
                wp = wp * ch ** lval
            wp = wp * Pall[ip * nq:(ip + 1) * nq]
            acc = acc + ((bf * wp[:, None]).T @ bf).reshape(-1, order="F")
            inner[ip] = acc
        return inner

    def outer(kval, lval):
        inner = inner_integral(lval)
        wp = el.mulen * el.wx * el.shmu
        if kval != 0:
            wp = wp * el.chmu ** kval
        wp = wp * Qouter
        return (el.bfprod * wp[:, None]).T @ inner

    return outer(k, l) + outer(l, k).T


def sign_cholesky(W, thresh):
    """Sign-aware diagonal-pivoted Cholesky W = B diag(sigma) B^T;
    basis.cpp:1498-1537."""
    N = W.shape[0]
    d = W.diagonal().copy()
    dmax0 = np.max(np.abs(d))
    cols, sgn = [], []
    for _ in range(N):
        piv = int(np.argmax(np.abs(d)))
        dpiv = abs(d[piv])
        if dmax0 <= 0.0 or dpiv <= thresh * dmax0:
            break
        dp = d[piv]
        s = 1.0 if dp >= 0.0 else -1.0
        col = W[:, piv].copy()
        for q in range(len(cols)):
            col -= sgn[q] * cols[q][piv] * cols[q]
        col /= np.sqrt(abs(dp))
        d -= s * col * col
        d[piv] = 0.0
        cols.append(col); sgn.append(s)
    return np.array(cols).T.reshape(N, len(cols)), np.array(sgn)


class RadialBasis:
    def __init__(self, febasis, nquad):
        self.fem = febasis
        self.xq, self.wq = fem.chebyshev(nquad)

    def Nel(self): return self.fem.nel
    def Nbf(self): return self.fem.nbf
    def get_idx(self, iel): return self.fem.idx(iel)

    def _assemble(self, der, x, w, f):
        M = np.zeros((self.Nbf(), self.Nbf()))
        ev = lambda xx, iel: self.fem.eval_dnf(xx, der, iel)
        for iel in range(self.Nel()):
            a, b = self.get_idx(iel)
            M[a:b + 1, a:b + 1] += self.fem.matrix_element(iel, ev, ev, x, w, f)
        return M

    def _conv(self, probe, **kw):
        return fem.converge(probe, max(len(self.xq), 5), TWOE_NMAX, floor_rel=256 * np.finfo(float).eps, **kw)

    def radial_integral(self, m, n):
        """int B B sinh^m cosh^n dmu; basis.cpp:258-281."""
        f = lambda mu: np.sinh(mu) ** m * np.cosh(mu) ** n
        return self._conv(lambda nq: self._assemble(0, *fem.lobatto(nq), f))

    def kinetic(self):
        """basis.cpp:378-391."""
        return self._conv(lambda nq: self._assemble(1, *fem.lobatto(nq), np.sinh))

    def PQ_integral(self, which, k, iel, L, leg):
        """Plm_integral / Qlm_integral, basis.cpp:352-376 (Gauss-Chebyshev,
        seed fallback)."""
        ev = lambda xx, ie: self.fem.eval_dnf(xx, 0, ie)

        def f(mu):
            ch = np.cosh(mu)
            tab = leg.P(ch) if which == "P" else leg.Q(ch)
            return np.sinh(mu) * ch ** k * tab[L]

        return self._conv(lambda n: self.fem.matrix_element(iel, ev, ev, *fem.chebyshev(n), f), seed_fallback=True)


class TwoDBasis:
    def __init__(self, Z1, Z2, Rhalf, nnodes, nquad, bval, lval, mval):
        self.Z1, self.Z2, self.Rhalf = Z1, Z2, Rhalf
        self.radial = RadialBasis(fem.FEBasis(nnodes, bval, False, True), nquad)
        self.lval, self.mval = np.asarray(lval), np.asarray(mval)
        self.gaunt = Gaunt()
        lm, LM = set(), set()
        na = len(self.lval)
        for i in range(na):
            for j in range(na):
                M = self.mval[j] - self.mval[i]
                for L in range(max(abs(self.lval[j] - self.lval[i]) - 2, abs(M)), self.lval[j] + self.lval[i] + 3):
                    lm.add((L, abs(M))); LM.add((L, M))
        self.lm_map = sorted(lm, key=lambda p: (p[1], p[0]))      # |M|-major (lm_less)
        self.LM_map = sorted(LM)                                   # L-major
        self.lm_index = {p: i for i, p in enumerate(self.lm_map)}
        self.LM_index = {p: i for i, p in enumerate(self.LM_map)}
        self.pure_idx = self.pure_indices()
        self.absm_symmetric = False
        self.cd_B = None

    def Nrad(self): return self.radial.Nbf()
    def Nang(self): return len(self.lval)
    def Ndummy(self): return self.Nang() * self.Nrad()
    def Nbf(self): return len(self.pure_idx)

    def pure_indices(self):
        N = self.Nrad()
        idx = []
        for i, m in enumerate(self.mval):
            idx += list(range(i * N + (0 if m == 0 else 1), (i + 1) * N))
        return np.array(idx)

    def remove_boundaries(self, F):
        return F[np.ix_(self.pure_idx, self.pure_idx)]

    def expand_boundaries(self, P):
        out = np.zeros((self.Ndummy(), self.Ndummy()))
        out[np.ix_(self.pure_idx, self.pure_idx)] = P
        return out

    # ---- one-electron matrices (only to pin the oracle through SCF energies)
    def _fill(self, fn):
        N = self.Nrad(); na = self.Nang()
        M = np.zeros((self.Ndummy(), self.Ndummy()))
        for i in range(na):
            for j in range(na):
                blk = fn(i, j)
                if blk is not None:
                    M[i * N:(i + 1) * N, j * N:(j + 1) * N] = blk
        return M

    def overlap(self):
        I10, I12 = self.radial.radial_integral(1, 0), self.radial.radial_integral(1, 2)
        lv, mv, g = self.lval, self.mval, self.gaunt

        def fn(i, j):
            if mv[i] != mv[j]:
                return None
            blk = I12.copy() if lv[i] == lv[j] else np.zeros_like(I12)
            return blk - g.cosine2_coupling(lv[j], mv[j], lv[i], mv[i]) * I10

        return self.remove_boundaries(self._fill(fn) * self.Rhalf ** 3)

    def kinetic(self):
        Trad, Ip1, Im1 = self.radial.kinetic(), self.radial.radial_integral(1, 0), self.radial.radial_integral(-1, 0)
        lv, mv = self.lval, self.mval

        def fn(i, j):
            if i != j:
                return None
            return Trad + lv[i] * (lv[i] + 1) * Ip1 + mv[i] * mv[i] * Im1

        return self.remove_boundaries(self._fill(fn) * self.Rhalf / 2.0)

    def nuclear(self):
        I10, I11 = self.radial.radial_integral(1, 0), self.radial.radial_integral(1, 1)
        lv, mv, g = self.lval, self.mval, self.gaunt

        def fn(i, j):
            if mv[i] != mv[j]:
                return None
            blk = (self.Z1 + self.Z2) * I11 if lv[i] == lv[j] else np.zeros_like(I11)
            if self.Z1 != self.Z2:
                blk = blk + (self.Z2 - self.Z1) * g.cosine_coupling(lv[j], mv[j], lv[i], mv[i]) * I10
            return blk

        return self.remove_boundaries(self._fill(fn) * (-self.Rhalf ** 2))

    # ---- two-electron setup
    def _kernel_W(self, el, L, M, Ptab, Qtab):
        T00 = twoe_integral(el, 0, 0, L, Ptab[L], Qtab[L])
        T02 = twoe_integral(el, 0, 2, L, Ptab[L], Qtab[L])
        T22 = twoe_integral(el, 2, 2, L, Ptab[L], Qtab[L])
        W = np.block([[T00, -T02], [-T02.T, T22]])
        return 0.5 * (W + W.T)

    def converged_twoe_order(self, iel):
        """basis.cpp:1334-1380: probe the hardest multipole."""
        nstart = min(max(len(self.radial.xq), 5), TWOE_NMAX)
        L, M = max(self.lm_map, key=lambda p: (p[0], p[1]))

        def probe(n):
            el = TwoeElement(self.radial.fem, iel, n)
            leg = LegCache(M, L)
            return self._kernel_W(el, L, M, leg.P(el.sub_chmu), leg.Q(el.chmu))

        _, n = fem.converge(probe, nstart, TWOE_NMAX, floor_rel=256 * np.finfo(float).eps, want_n=True)
        return n

    def compute_tei(self):
        """basis.cpp:1382-1547."""
        Nel = self.radial.Nel()
        nlm = len(self.lm_map)
        self.disjoint_P0 = [None] * (Nel * nlm); self.disjoint_P2 = [None] * (Nel * nlm)
        self.disjoint_Q0 = [None] * (Nel * nlm); self.disjoint_Q2 = [None] * (Nel * nlm)
        runs = {}
        for ilm, (L, M) in enumerate(self.lm_map):
            runs.setdefault(M, []).append(ilm)
        for M, ilms in runs.items():
            Lhi = max(self.lm_map[i][0] for i in ilms)
            leg = LegCache(M, Lhi)
            for iel in range(Nel):
                for ilm in ilms:
                    L = self.lm_map[ilm][0]
                    self.disjoint_P0[ilm * Nel + iel] = self.radial.PQ_integral("P", 0, iel, L, leg)
                    self.disjoint_P2[ilm * Nel + iel] = self.radial.PQ_integral("P", 2, iel, L, leg)
                    self.disjoint_Q0[ilm * Nel + iel] = self.radial.PQ_integral("Q", 0, iel, L, leg)
                    self.disjoint_Q2[ilm * Nel + iel] = self.radial.PQ_integral("Q", 2, iel, L, leg)
        self.cd_thresh = 1e-12
        self.cd_B = [None] * (Nel * nlm)
        self.cd_sigma = [None] * (Nel * nlm)
        for iel in range(Nel):
            nconv = self.converged_twoe_order(iel)
            el = TwoeElement(self.radial.fem, iel, nconv)
            for M, ilms in runs.items():
                Lhi = max(self.lm_map[i][0] for i in ilms)
                leg = LegCache(M, Lhi)
                Ptab, Qtab = leg.P(el.sub_chmu), leg.Q(el.chmu)
                for ilm in ilms:
                    L = self.lm_map[ilm][0]
                    W = self._kernel_W(el, L, M, Ptab, Qtab)
                    B, s = sign_cholesky(W, self.cd_thresh)
                    self.cd_B[ilm * Nel + iel] = B
                    self.cd_sigma[ilm * Nel + iel] = s

    def LMfac_abs(self):
        """basis.cpp:1549-1560."""
        out = []
        for L, Ma in self.lm_map:
            fr = 1.0
            for p in range(L + Ma, L - Ma, -1):
                fr *= p
            out.append(4.0 * np.pi * self.Rhalf ** 5 / fr)
        return out

    def coulomb(self, P_in):
        """basis.cpp:1627-1816."""
        if self.cd_B is None:
            raise RuntimeError("Primitive teis have not been computed!")
        P = self.expand_boundaries(P_in)
        Nel, N, na = self.radial.Nel(), self.Nrad(), self.Nang()
        lv, mv, g = self.lval, self.mval, self.gaunt
        nLM = len(self.LM_map)
        Paux0 = [np.zeros((N, N)) for _ in range(nLM)]
        Paux2 = [np.zeros((N, N)) for _ in range(nLM)]
        for k in range(na):
            for l in range(na):
                M = mv[k] - mv[l]
                Prad = P[k * N:(k + 1) * N, l * N:(l + 1) * N]
                for L in range(max(abs(lv[k] - lv[l]) - 2, abs(M)), lv[k] + lv[l] + 3):
                    iLM = self.LM_index[(L, M)]
                    c0 = g.mod_coeff(lv[k], mv[k], L, M, lv[l], mv[l])
                    c2 = g.coeff(lv[k], mv[k], L, M, lv[l])
                    if c0 != 0.0:
                        Paux0[iLM] += c0 * Prad
                    if c2 != 0.0:
                        Paux2[iLM] += c2 * Prad
        Jaux0 = [np.zeros((N, N)) for _ in range(nLM)]
        Jaux2 = [np.zeros((N, N)) for _ in range(nLM)]
        fac = self.LMfac_abs()
        for iLM, (L, M) in enumerate(self.LM_map):
            ilm = self.lm_index[(L, abs(M))]
            LMfac = (-1.0 if (M & 1) else 1.0) * fac[ilm]
            for jel in range(Nel):
                a, b = self.radial.get_idx(jel)
                Ps0 = Paux0[iLM][a:b + 1, a:b + 1]; Ps2 = Paux2[iLM][a:b + 1, a:b + 1]
                js0 = LMfac * np.trace(self.disjoint_P0[ilm * Nel + jel] @ Ps0)
                jb0 = LMfac * np.trace(self.disjoint_Q0[ilm * Nel + jel] @ Ps0)
                js2 = LMfac * np.trace(self.disjoint_P2[ilm * Nel + jel] @ Ps2)
                jb2 = LMfac * np.trace(self.disjoint_Q2[ilm * Nel + jel] @ Ps2)
                for iel in range(jel):
                    c, d = self.radial.get_idx(iel)
                    Jaux0[iLM][c:d + 1, c:d + 1] += (jb0 - jb2) * self.disjoint_P0[ilm * Nel + iel]
                    Jaux2[iLM][c:d + 1, c:d + 1] += (-jb0 + jb2) * self.disjoint_P2[ilm * Nel + iel]
                for iel in range(jel + 1, Nel):
                    c, d = self.radial.get_idx(iel)
                    Jaux0[iLM][c:d + 1, c:d + 1] += (js0 - js2) * self.disjoint_Q0[ilm * Nel + iel]
                    Jaux2[iLM][c:d + 1, c:d + 1] += (-js0 + js2) * self.disjoint_Q2[ilm * Nel + iel]
                B, sig = self.cd_B[ilm * Nel + jel], self.cd_sigma[ilm * Nel + jel]
                Ni = b - a + 1; nn = Ni * Ni
                p2 = np.concatenate([Ps0.reshape(-1, order="F"), Ps2.reshape(-1, order="F")])
                jv = LMfac * (B @ (sig * (B.T @ p2)))
                Jaux0[iLM][a:b + 1, a:b + 1] += jv[:nn].reshape(Ni, Ni, order="F")
                Jaux2[iLM][a:b + 1, a:b + 1] += jv[nn:].reshape(Ni, Ni, order="F")
        J = np.zeros_like(P)
        for i in range(na):
            for j in range(na):
                M = mv[j] - mv[i]
                for L in range(max(abs(lv[j] - lv[i]) - 2, abs(M)), lv[j] + lv[i] + 3):
                    iLM = self.LM_index[(L, M)]
                    c0 = g.mod_coeff(lv[j], mv[j], L, M, lv[i], mv[i])
                    if c0 != 0.0:
                        J[i * N:(i + 1) * N, j * N:(j + 1) * N] += c0 * Jaux0[iLM]
                    c2 = g.coeff(lv[j], mv[j], L, M, lv[i])
                    if c2 != 0.0:
                        J[i * N:(i + 1) * N, j * N:(j + 1) * N] += c2 * Jaux2[iLM]
        return self.remove_boundaries(J)

    def exchange(self, P_in):
        """basis.cpp:1818-2089; returns -K like the reference."""
        if self.cd_B is None:
            raise RuntimeError("Primitive teis have not been computed!")
        P = self.expand_boundaries(P_in)
        Nel, N, na = self.radial.Nel(), self.Nrad(), self.Nang()
        lv, mv, g = self.lval, self.mval, self.gaunt
        fac = self.LMfac_abs()
        K = np.zeros_like(P)
        mabs = int(np.max(np.abs(mv)))
        pairs = {}
        for i in range(na):
            for l in range(na):
                if np.linalg.norm(P[i * N:(i + 1) * N, l * N:(l + 1) * N]) < 10 * np.finfo(float).eps:
                    continue
                pairs.setdefault(mv[i] - mv[l], []).append((i, l))
        for j in range(na):
            for k in range(na):
                if self.absm_symmetric and (mv[j] < 0 or mv[k] < 0):
                    continue
                cand = pairs.get(mv[j] - mv[k], [])
                R = {}
                for (i, l) in cand:
                    M = mv[j] - mv[i]
                    Lmin = max(max(abs(lv[i] - lv[j]), abs(lv[k] - lv[l])) - 2, abs(M))
                    Lmax = min(lv[i] + lv[j], lv[k] + lv[l]) + 2
                    Psub = P[i * N:(i + 1) * N, l * N:(l + 1) * N]
                    for L in range(Lmin, Lmax + 1):
                        mj0 = g.mod_coeff(lv[j], mv[j], L, M, lv[i], mv[i]); mk0 = g.mod_coeff(lv[k], mv[k], L, M, lv[l], mv[l])
                        cj2 = g.coeff(lv[j], mv[j], L, M, lv[i]); ck2 = g.coeff(lv[k], mv[k], L, M, lv[l])
                        c00, c02, c20, c22 = mj0 * mk0, -mj0 * ck2, -cj2 * mk0, cj2 * ck2
                        if c00 == 0.0 and c02 == 0.0 and c20 == 0.0 and c22 == 0.0:
                            continue
                        ilm = self.lm_index[(L, abs(M))]
                        LMfac = (-1.0 if (M & 1) else 1.0) * fac[ilm]
                        if ilm not in R:
                            R[ilm] = [np.zeros((N, N)) for _ in range(4)]
                        R[ilm][0] += (LMfac * c00) * Psub; R[ilm][1] += (LMfac * c02) * Psub
                        R[ilm][2] += (LMfac * c20) * Psub; R[ilm][3] += (LMfac * c22) * Psub
                if not R:
                    continue
                for iel in range(Nel):
                    a, b = self.radial.get_idx(iel); Ni = b - a + 1
                    for jel in range(Nel):
                        c, d = self.radial.get_idx(jel)
                        blk = (slice(j * N + a, j * N + b + 1), slice(k * N + c, k * N + d + 1))
                        if iel == jel:
                            Ks = np.zeros((Ni, Ni)); nn = Ni * Ni
                            for ilm in sorted(R):
                                B, sig = self.cd_B[ilm * Nel + iel], self.cd_sigma[ilm * Nel + iel]
                                R00, R02, R20, R22 = (r[a:b + 1, c:d + 1] for r in R[ilm])
                                for p in range(B.shape[1]):
                                    M0t = B[:nn, p].reshape(Ni, Ni, order="F").T
                                    M2t = B[nn:, p].reshape(Ni, Ni, order="F").T
                                    Ks += sig[p] * (M0t @ R00 @ M0t - M0t @ R02 @ M2t - M2t @ R20 @ M0t + M2t @ R22 @ M2t)
                            K[blk] -= Ks
                        else:
                            Ks = np.zeros((Ni, d - c + 1))
                            for ilm in sorted(R):
                                if iel > jel:
                                    i0, i2 = self.disjoint_Q0[ilm * Nel + iel], self.disjoint_Q2[ilm * Nel + iel]
                                    j0, j2 = self.disjoint_P0[ilm * Nel + jel], self.disjoint_P2[ilm * Nel + jel]
                                else:
                                    i0, i2 = self.disjoint_P0[ilm * Nel + iel], self.disjoint_P2[ilm * Nel + iel]
                                    j0, j2 = self.disjoint_Q0[ilm * Nel + jel], self.disjoint_Q2[ilm * Nel + jel]
                                R00, R02, R20, R22 = (r[a:b + 1, c:d + 1] for r in R[ilm])
                                Ks -= i0 @ (R00 @ j0.T + R02 @ j2.T)
                                Ks -= i2 @ (R20 @ j0.T + R22 @ j2.T)
                            K[blk] += Ks
        if self.absm_symmetric:
            mirror = {}
            for a1 in range(na):
                for b1 in range(na):
                    if lv[b1] == lv[a1] and mv[b1] == -mv[a1]:
                        mirror[a1] = b1
                        break
            for j in range(na):
                if mv[j] >= 0 or j not in mirror:
                    continue
                for k in range(na):
                    if mv[k] >= 0 or k not in mirror:
                        continue
                    jm, km = mirror[j], mirror[k]
                    K[j * N:(j + 1) * N, k * N:(k + 1) * N] = K[jm * N:(jm + 1) * N, km * N:(km + 1) * N]
        return self.remove_boundaries(K)
