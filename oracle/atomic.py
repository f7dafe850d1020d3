"""Atomic (l,m) x radial-FE basis: setup, J and K.  Oracle restatement (numpy).

Test infrastructure only.  Follows src/atomic/TwoDBasis.cpp (ctor :66-94,
1e matrices :320-375, compute_tei :708-735, coulomb :773-877, exchange
:879-999), libhelfemqc/include/CoulombExchangeFE.h:180-217,432-530,
libhelfem/src/RadialBasis.cpp:245-257,454-497,643-716,868-878 and
libhelfem/src/quadrature.cpp:37-161.
"""
import numpy as np

from . import fem
from .gaunt import Gaunt


def angular_basis(lmax, mmax):
    """src/atomic/basis.cpp:179-203 (m-major: |m|, then l, +m before -m)."""
    lval, mval = [], []
    for mabs in range(mmax + 1):
        for l in range(mabs, lmax + 1):
            lval.append(l); mval.append(mabs)
            if mabs > 0:
                lval.append(l); mval.append(-mabs)
    return np.array(lval), np.array(mval)


def bessel_il(x, L):
    """Modified spherical Bessel function i_L; power series below 0.5 as libhelfem/include/math.h:68-91."""
    from scipy.special import iv
    x = np.atleast_1d(np.asarray(x, dtype=float))
    out = np.zeros_like(x)
    small = np.abs(x) < 0.5
    if small.any():
        a = np.abs(x[small])
        dfac = 1.0
        for j in range(3, 2 * L + 2, 2):
            dfac *= j
        term = a ** L / dfac
        val = term.copy()
        for k in range(1, 256):
            term = term * (0.5 * a * a) / (k * (2 * L + 2 * k + 1))
            val += term
        out[small] = val
    if (~small).any():
        a = np.abs(x[~small])
        out[~small] = iv(L + 0.5, a) * np.sqrt(np.pi / (2 * a))
    return out if out.size > 1 else float(out[0])


def bessel_kl(x, L):
    """k_L(x) = sqrt(2/(pi x)) K_(L+1/2)(x); libhelfem/include/math.h:101-105."""
    from scipy.special import kv
    x = np.asarray(x, dtype=float)
    with np.errstate(all="ignore"):
        return kv(L + 0.5, x) * np.sqrt(2 / (np.pi * x))


def pivoted_cholesky(A, tol):
    """Diagonal-pivoted Cholesky, absolute tolerance on the residual diagonal;
    libhelfem/src/RadialBasis.cpp:670-709."""
    n = A.shape[0]
    D = A.diagonal().copy()
    done = np.zeros(n, dtype=bool)
    cols = []
    for _ in range(n):
        cand = np.where(~done & (D > tol))[0]
        if len(cand) == 0:
            break
        pivot = cand[np.argmax(D[cand])]
        pv = D[pivot]
        done[pivot] = True
        sd = np.sqrt(pv)
        s = A[:, pivot].copy()
        for c in cols:
            s -= c * c[pivot]
        col = s / sd
        col[done] = 0.0
        col[pivot] = sd
        cols.append(col)
        D[~done] -= col[~done] ** 2
    return np.array(cols).T.reshape(n, len(cols))


class RadialBasis:
    """libhelfem FEMRadialBasisT<double> in u = rR form."""

    def __init__(self, febasis, nquad):
        self.fem = febasis
        self.xq, self.wq = fem.chebyshev(nquad)

    def Nel(self): return self.fem.nel
    def Nbf(self): return self.fem.nbf
    def Nprim(self, iel): return self.fem.nprim(iel)
    def get_idx(self, iel): return self.fem.idx(iel)

    def get_bf(self, x, iel):
        """B/r; RadialBasis.cpp:868-878."""
        if iel == 0:
            return self.fem.eval_over_r(x, 0, iel)
        return self.fem.eval_dnf(x, 0, iel) / self.fem.coord(x, iel)[:, None]

    def _B(self, n):
        return lambda x, iel: self.fem.eval_dnf(x, n, iel)

    def radial_integral(self, Rexp, iel):
        """int (B/r)(B/r) r^(Rexp+2) dr; RadialBasis.cpp:245-257."""
        return self.fem.matrix_element_auto(iel, self.get_bf, self.get_bf, lambda r: r ** float(Rexp + 2))

    def assemble(self, blockfn):
        M = np.zeros((self.Nbf(), self.Nbf()))
        for iel in range(self.Nel()):
            a, b = self.get_idx(iel)
            M[a:b + 1, a:b + 1] += blockfn(iel)
        return M

    def overlap(self):
        return self.assemble(lambda iel: self.fem.matrix_element_auto(iel, self._B(0), self._B(0), None, poly_degree_f=0))

    def kinetic(self):
        return self.assemble(lambda iel: 0.5 * self.fem.matrix_element_auto(iel, self._B(1), self._B(1), None, poly_degree_f=0))

    def kinetic_l(self):
        return self.assemble(lambda iel: 0.5 * self.fem.matrix_element_auto(iel, self.get_bf, self.get_bf, None))

    def nuclear(self):
        return self.assemble(lambda iel: -self.fem.matrix_element_auto(iel, self.get_bf, self.get_bf, lambda r: r))

    # ---- two-electron in-element integrals: libhelfem/src/quadrature.cpp:37-161
    def _twoe_fixed(self, iel, L, x, wx, fsb=None, fbig=None):
        rmin, rmax = self.fem.begin(iel), self.fem.end(iel)
        en = self.fem.enabled(iel)
        x0 = self.fem.x0
        rmid, rlen = 0.5 * (rmax + rmin), 0.5 * (rmax - rmin)
        r = rmid + rlen * x
        nq = len(x)
        nbf = len(en)
        inner = np.zeros((nq, nbf * nbf))

        def wrk(a, b):
            # quadrature.cpp:37-75 with fsmallbig(r,R)=(r/R)^L/R
            smid, slen = 0.5 * (b + a), 0.5 * (b - a)
            rs = smid + slen * x
            fv = (rs / b) ** L / b if fsb is None else fsb(rs, b)
            wp = wx * fv * slen
            xpoly = (rs - rmid) / rlen
            bf = fem.lip_eval(xpoly, x0, 0)[:, en]
            return ((bf * wp[:, None]).T @ bf).reshape(-1, order="F")

        empty_first = (r[0] == rmin)
        if not empty_first:
            inner[0] = wrk(rmin, r[0])
        for ip in range(1, nq):
            inner[ip] = wrk(r[ip - 1], r[ip])
        for ip in range(1, nq):
            if ip == 1 and empty_first:
                continue
            if fbig is None:
                inner[ip] += inner[ip - 1] * (r[ip] ** float(-L - 1) / r[ip - 1] ** float(-L - 1))
            else:
                inner[ip] += inner[ip - 1] * (fbig(r[ip]) / fbig(r[ip - 1]))
        bf = fem.lip_eval(x, x0, 0)[:, en]
        bfprod = (bf[:, :, None] * bf[:, None, :]).reshape(nq, nbf * nbf)
        bfprod = bfprod * (wx * rlen)[:, None]
        ints = bfprod.T @ inner
        return ints + ints.T

    def twoe_integral(self, L, iel):
        """RadialBasis.cpp:643-663 (order-doubling Gauss-Lobatto)."""
        nstart = min(max(len(self.xq), 5), 512)
        return fem.converge(lambda n: self._twoe_fixed(iel, L, *fem.lobatto(n)), nstart, 512)

    # ---- Yukawa kernel: libhelfem/src/quadrature.cpp:164-201, libhelfem/include/math.h:68-110
    def yukawa_integral(self, L, lam, iel):
        il = lambda x: bessel_il(x, L)
        kl = lambda x: bessel_kl(x, L)
        nstart = min(max(len(self.xq), 5), 512)
        return fem.converge(lambda n: self._twoe_fixed(iel, L, *fem.lobatto(n),
                                                       fsb=lambda r, R: il(r * lam) * kl(R * lam), fbig=lambda r: kl(r * lam)),
                            nstart, 512)

    def bessel_integral(self, which, L, lam, iel):
        """int B B i_L(lam r) dr / int B B k_L(lam r) dr; RadialBasis.cpp:320-330."""
        if which == "i":
            f = lambda r: bessel_il(r * lam, L)
        else:
            f = lambda r: bessel_kl(r * lam, L)
        return self.fem.matrix_element_auto(iel, self._B(0), self._B(0), f)

    def erfc_integral(self, L, mu, iel, kel):
        """RadialBasis.cpp:742-810 + quadrature.cpp:201-249: dense (pair in iel) x (pair in kel) tensor of the
        erfc Green's function Phi_L(mu r, mu r'); fixed rule, Nq sub-intervals in r' across the cusp when
        iel == kel, symmetrised there.  Pair index fi*N+fj (make_bfprod, quadrature.cpp:126-133)."""
        from . import erfc as erfc_mod
        x0 = self.fem.x0
        Nq = len(self.xq)
        xi, wi = fem.chebyshev(Nq)
        Nint = Nq if iel == kel else 1
        xk = np.empty(Nq * Nint); wk = np.empty(Nq * Nint)
        for ii in range(Nint):
            istart = ii * 2.0 / Nint - 1.0
            iend = (ii + 1) * 2.0 / Nint - 1.0
            imid, ilen = 0.5 * (iend + istart), 0.5 * (iend - istart)
            xk[ii * Nq:(ii + 1) * Nq] = imid + xi * ilen
            wk[ii * Nq:(ii + 1) * Nq] = wi * ilen
        ibf = fem.lip_eval(xi, x0, 0)[:, self.fem.enabled(iel)]
        kbf = fem.lip_eval(xk, x0, 0)[:, self.fem.enabled(kel)]
        rmini, rmaxi = self.fem.begin(iel), self.fem.end(iel)
        rmink, rmaxk = self.fem.begin(kel), self.fem.end(kel)
        rmidi, rleni = 0.5 * (rmaxi + rmini), 0.5 * (rmaxi - rmini)
        rmidk, rlenk = 0.5 * (rmaxk + rmink), 0.5 * (rmaxk - rmink)
        ri = rmidi + rleni * xi
        rk = rmidk + rlenk * xk
        Fn = erfc_mod.Phi(L, mu * ri[:, None], mu * rk[None, :])
        ni, nk = ibf.shape[1], kbf.shape[1]
        bpi = (ibf[:, :, None] * ibf[:, None, :]).reshape(len(xi), ni * ni) * (wi * rleni)[:, None]
        bpk = (kbf[:, :, None] * kbf[:, None, :]).reshape(len(xk), nk * nk) * (wk * rlenk)[:, None]
        tei = bpi.T @ Fn @ bpk
        if iel == kel:
            tei = 0.5 * (tei + tei.T)
        return tei

    def twoe_integral_cholesky(self, L, iel, tol=1e-12):
        return pivoted_cholesky(self.twoe_integral(L, iel), tol)


class TwoDBasis:
    """src/atomic/TwoDBasis.{h,cpp}, T=double, point nucleus, LIP primbas=4."""

    def __init__(self, Z, nnodes, nquad, bval, lval, mval):
        self.Z = Z
        fe = fem.FEBasis(nnodes, bval, True, True)
        self.radial = RadialBasis(fe, nquad)
        self.lval = np.asarray(lval)
        self.mval = np.asarray(mval)
        self.prim_chol = None

    def Nrad(self): return self.radial.Nbf()
    def Nang(self): return len(self.lval)
    def Nbf(self): return self.Nang() * self.Nrad()

    def _diag(self, blockfn):
        N = self.Nrad()
        M = np.zeros((self.Nbf(), self.Nbf()))
        for ia in range(self.Nang()):
            M[ia * N:(ia + 1) * N, ia * N:(ia + 1) * N] = blockfn(ia)
        return M

    def overlap(self):
        O = self.radial.assemble(lambda iel: self.radial.radial_integral(0, iel))
        return self._diag(lambda ia: O)

    def kinetic(self):
        T = self.radial.kinetic()
        Tl = self.radial.kinetic_l()
        return self._diag(lambda ia: T + self.lval[ia] * (self.lval[ia] + 1) * Tl)

    def nuclear(self):
        V = self.radial.assemble(lambda iel: self.radial.radial_integral(-1, iel))
        return self._diag(lambda ia: -self.Z * V)

    def compute_tei(self):
        """TwoDBasis.cpp:708-735 + CoulombExchangeFE.h:180-217."""
        N_L = 2 * int(self.lval.max()) + 1
        Nel = self.radial.Nel()
        self.N_L = N_L
        self.disjoint_L = [None] * (N_L * Nel)
        self.disjoint_m1L = [None] * (N_L * Nel)
        self.prim_chol = [None] * (N_L * Nel)
        for L in range(N_L):
            for iel in range(Nel):
                self.disjoint_L[L * Nel + iel] = self.radial.radial_integral(L, iel)
                if iel > 0:
                    self.disjoint_m1L[L * Nel + iel] = self.radial.radial_integral(-L - 1, iel)
                self.prim_chol[L * Nel + iel] = self.radial.twoe_integral_cholesky(L, iel)

    def compute_yukawa(self, lam):
        """TwoDBasis.cpp:737-758."""
        N_L = 2 * int(self.lval.max()) + 1
        Nel = self.radial.Nel()
        self.lam = lam
        self.yukawa = True
        self.disjoint_iL = [None] * (N_L * Nel); self.disjoint_kL = [None] * (N_L * Nel); self.rs_chol = [None] * (N_L * Nel)
        for L in range(N_L):
            for iel in range(Nel):
                self.disjoint_iL[L * Nel + iel] = self.radial.bessel_integral("i", L, lam, iel)
                if iel > 0:
                    self.disjoint_kL[L * Nel + iel] = self.radial.bessel_integral("k", L, lam, iel)
                self.rs_chol[L * Nel + iel] = pivoted_cholesky(self.radial.yukawa_integral(L, lam, iel), 1e-12)

    def compute_erfc(self, mu):
        """TwoDBasis.cpp:762-771 + CoulombExchangeFE.h:275-297: dense exchange-ordered pair tensors
        rs_ktei[L*Nel*Nel + iel*Nel + jel] = exchange_tei(erfc_integral(L, mu, iel, jel)) (utils.cpp:55-80:
        ktei[kk*Ni + jj, ll*Ni + ii] = tei[jj*Ni + ii, ll*Nj + kk])."""
        N_L = 2 * int(self.lval.max()) + 1
        Nel = self.radial.Nel()
        self.lam = mu
        self.yukawa = False
        self.rs_ktei = [None] * (N_L * Nel * Nel)
        for L in range(N_L):
            for iel in range(Nel):
                Ni = self.radial.Nprim(iel)
                for jel in range(Nel):
                    Nj = self.radial.Nprim(jel)
                    tei = self.radial.erfc_integral(L, mu, iel, jel).reshape(Ni, Ni, Nj, Nj)   # [jj, ii, ll, kk]
                    # ktei[(kk, jj), (ll, ii)], row index kk*Ni + jj, column index ll*Ni + ii
                    self.rs_ktei[(L * Nel + iel) * Nel + jel] = tei.transpose(3, 0, 2, 1).reshape(Nj * Ni, Nj * Ni)

    def _assemble_K_pairwise(self, L, P):
        """CoulombExchangeFE.h:396-422."""
        Nel = self.radial.Nel()
        K = np.zeros_like(P)
        for iel in range(Nel):
            a, b = self.radial.get_idx(iel)
            Ni = b - a + 1
            for jel in range(Nel):
                c, d = self.radial.get_idx(jel)
                Nj = d - c + 1
                Psub = P[a:b + 1, c:d + 1]
                Kv = self.rs_ktei[(L * Nel + iel) * Nel + jel] @ Psub.reshape(-1, order="F")
                K[a:b + 1, c:d + 1] += Kv.reshape(Ni, Nj, order="F")
        return K

    def rs_exchange(self, P):
        """TwoDBasis.cpp:1001-1131.  Yukawa branch: same loops as exchange() with the screened caches and
        prefactor 4 pi lambda; erfc branch: prefactor 4 pi mu / (2L+1) and the pairwise assembler."""
        if getattr(self, "yukawa", True) is False:
            save_K = self._assemble_K
            self._assemble_K = self._assemble_K_pairwise
            try:
                return self.exchange(P, Lfac=lambda L: 4.0 * np.pi * self.lam / (2 * L + 1))
            finally:
                self._assemble_K = save_K
        save = (self.disjoint_L, self.disjoint_m1L, self.prim_chol)
        self.disjoint_L, self.disjoint_m1L, self.prim_chol = self.disjoint_iL, self.disjoint_kL, self.rs_chol
        try:
            return self.exchange(P, Lfac=lambda L: 4.0 * np.pi * self.lam)
        finally:
            self.disjoint_L, self.disjoint_m1L, self.prim_chol = save

    # ---- FE assemblers, CoulombExchangeFE.h:432-530
    def _assemble_J(self, L, P):
        Nel = self.radial.Nel()
        J = np.zeros_like(P)
        for jel in range(Nel):
            a, b = self.radial.get_idx(jel)
            Psub = P[a:b + 1, a:b + 1]
            jsmall = np.trace(self.disjoint_L[L * Nel + jel] @ Psub)
            jbig = np.trace(self.disjoint_m1L[L * Nel + jel] @ Psub) if jel > 0 else 0.0
            for iel in range(jel):
                c, d = self.radial.get_idx(iel)
                J[c:d + 1, c:d + 1] += jbig * self.disjoint_L[L * Nel + iel]
            for iel in range(jel + 1, Nel):
                c, d = self.radial.get_idx(iel)
                J[c:d + 1, c:d + 1] += jsmall * self.disjoint_m1L[L * Nel + iel]
            Lf = self.prim_chol[L * Nel + jel]
            sc = Lf.T @ Psub.reshape(-1, order="F")
            Nj = b - a + 1
            J[a:b + 1, a:b + 1] += (Lf @ sc).reshape(Nj, Nj, order="F")
        return J

    def _assemble_K(self, L, P):
        Nel = self.radial.Nel()
        K = np.zeros_like(P)
        for iel in range(Nel):
            a, b = self.radial.get_idx(iel)
            Ni = b - a + 1
            for jel in range(Nel):
                c, d = self.radial.get_idx(jel)
                Psub = P[a:b + 1, c:d + 1]
                if iel == jel:
                    Lf = self.prim_chol[L * Nel + iel]
                    Ksub = np.zeros((Ni, Ni))
                    for p in range(Lf.shape[1]):
                        Mp = Lf[:, p].reshape(Ni, Ni, order="F")
                        Ksub += Mp @ (Psub @ Mp.T)
                    K[a:b + 1, c:d + 1] += Ksub
                else:
                    if iel > jel:
                        iint, jint = self.disjoint_m1L[L * Nel + iel], self.disjoint_L[L * Nel + jel]
                    else:
                        iint, jint = self.disjoint_L[L * Nel + iel], self.disjoint_m1L[L * Nel + jel]
                    K[a:b + 1, c:d + 1] += iint @ (Psub @ jint.T)
        return K

    def coulomb(self, P):
        """TwoDBasis.cpp:773-877."""
        if self.prim_chol is None:
            raise RuntimeError("Primitive teis have not been computed!")
        g = Gaunt()
        N = self.Nrad(); na = self.Nang()
        lv, mv = self.lval, self.mval
        Paux = {}
        for k in range(na):
            for l in range(na):
                M = mv[k] - mv[l]
                for L in range(max(abs(lv[k] - lv[l]), abs(M)), lv[k] + lv[l] + 1):
                    cpl = g.coeff(lv[k], mv[k], L, M, lv[l])
                    Paux[(L, M)] = Paux.get((L, M), 0.0) + cpl * P[k * N:(k + 1) * N, l * N:(l + 1) * N]
        Jaux = {key: 4.0 * np.pi / (2 * key[0] + 1) * self._assemble_J(key[0], val) for key, val in Paux.items()}
        J = np.zeros_like(P)
        for i in range(na):
            for j in range(na):
                M = mv[j] - mv[i]
                for L in range(max(abs(lv[j] - lv[i]), abs(M)), lv[j] + lv[i] + 1):
                    cpl = g.coeff(lv[j], mv[j], L, M, lv[i])
                    if cpl != 0.0 and (L, M) in Jaux:
                        J[i * N:(i + 1) * N, j * N:(j + 1) * N] += cpl * Jaux[(L, M)]
        return J

    def exchange(self, P, Lfac=None):
        """TwoDBasis.cpp:879-999; returns -K like the reference."""
        if self.prim_chol is None:
            raise RuntimeError("Primitive teis have not been computed!")
        g = Gaunt()
        N = self.Nrad(); na = self.Nang()
        lv, mv = self.lval, self.mval
        thr = 10 * np.finfo(float).eps
        K = np.zeros_like(P)
        norms = np.array([[np.linalg.norm(P[i * N:(i + 1) * N, l * N:(l + 1) * N]) for l in range(na)] for i in range(na)])
        for j in range(na):
            for k in range(na):
                R = {}
                for i in range(na):
                    for l in range(na):
                        M = mv[j] - mv[i]
                        if M != mv[k] - mv[l] or norms[i, l] < thr:
                            continue
                        Lmin = max(abs(lv[i] - lv[j]), abs(lv[k] - lv[l]), abs(M))
                        Lmax = min(lv[i] + lv[j], lv[k] + lv[l])
                        for L in range(Lmin, Lmax + 1):
                            cpl = g.coeff(lv[j], mv[j], L, M, lv[i]) * g.coeff(lv[k], mv[k], L, M, lv[l])
                            if cpl == 0.0:
                                continue
                            lf = 4.0 * np.pi / (2 * L + 1) if Lfac is None else Lfac(L)
                            R[L] = R.get(L, 0.0) + (lf * cpl) * P[i * N:(i + 1) * N, l * N:(l + 1) * N]
                Kb = np.zeros((N, N))
                for L, RL in sorted(R.items()):
                    Kb += self._assemble_K(L, RL)
                K[j * N:(j + 1) * N, k * N:(k + 1) * N] -= Kb
        return K
