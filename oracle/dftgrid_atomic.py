"""Atomic 3D DFT quadrature grid (r x theta x phi): density, gradient, tau, Laplacian and the
assembly of the XC matrix.  Oracle restatement (numpy, materialised basis tables exactly as the
reference builds them); test infrastructure only.

Follows src/atomic/dftgrid.cpp (compute_bf :470-576, update_density :51-128 / :130-242,
eval_Fxc :304-359 / :361-465, driver :591-712), src/atomic/TwoDBasis.cpp:1133-1255
(eval_bf/df/lf, bf_list), src/general/angular.cpp:21-69, src/general/spherical_harmonics.cpp:20-35,
src/general/dftgrid_common.{h,cpp} (accumulators :128-269, eval_Exc, compute_Nel).
The functional itself is NOT part of this file: callers provide exc/vrho/vsigma/vtau/vlapl in
libxc's layout (point-major, interleaved spin components).
"""
import numpy as np
from scipy.special import gammaln, lpmv

from . import fem


def sph_harm(l, m, cth, phi):
    """Y_lm with Condon-Shortley phase, as src/general/spherical_harmonics.cpp:20-35."""
    if m < 0:
        return np.conj((-1.0) ** m * sph_harm(l, -m, cth, phi))
    norm = np.sqrt((2 * l + 1) / (4 * np.pi) * np.exp(gammaln(l - m + 1) - gammaln(l + m + 1)))
    return norm * lpmv(m, l, cth) * np.exp(1j * m * phi)


def angular_chebyshev(lang, mang):
    """src/general/angular.cpp:21-69."""
    x, w = fem.chebyshev(lang)
    dphi = 2.0 * np.pi / mang
    cth = np.repeat(x, mang)
    phi = np.tile(np.arange(mang) * dphi, lang)
    wang = np.repeat(w, mang) * dphi
    return cth, phi, wang


class AtomicDFTGrid:
    def __init__(self, basis, lang, mang):
        self.b = basis
        self.cth, self.phi, self.wang = angular_chebyshev(lang, mang)

    # --- radial tables at the Chebyshev nodes, libhelfem/src/RadialBasis.cpp:868-926
    def _radial(self, iel):
        rb = self.b.radial
        xq = rb.xq
        r = rb.fem.coord(xq, iel)
        wrad = rb.wq * rb.fem.scale(iel)
        if iel == 0:
            f = rb.fem.eval_over_r(xq, 0, iel); d = rb.fem.eval_over_r(xq, 1, iel); l2 = rb.fem.eval_over_r(xq, 2, iel)
        else:
            B0 = rb.fem.eval_dnf(xq, 0, iel); B1 = rb.fem.eval_dnf(xq, 1, iel); B2 = rb.fem.eval_dnf(xq, 2, iel)
            ir = 1.0 / r[:, None]
            f = B0 * ir
            d = (-B0 * ir + B1) * ir
            l2 = ((2 * B0 * ir - 2 * B1) * ir + B2) * ir
        return r, wrad, f, d, l2

    def compute_bf(self, iel):
        """Materialised tables (nbf_el x npts), point index ia*nrad + ir."""
        b = self.b
        lv, mv = b.lval, b.mval
        r, wrad, f, d, l2 = self._radial(iel)
        nrad, nang, Nr = len(r), len(self.wang), f.shape[1]
        a0, a1 = b.radial.get_idx(iel)
        N = b.Nrad()
        self.bf_ind = np.array([N * ia + a0 + j for ia in range(len(lv)) for j in range(Nr)])
        sth = np.sqrt(1.0 - self.cth ** 2)
        self.scale = np.stack([np.ones(nang * nrad), np.tile(r, nang), (sth[:, None] * r[None, :]).ravel()])
        self.wtot = (self.wang[:, None] * (wrad * r * r)[None, :]).ravel()
        nbf = len(self.bf_ind)
        bf = np.zeros((nbf, nang * nrad), complex); dr = np.zeros_like(bf); dth = np.zeros_like(bf)
        dph = np.zeros_like(bf); lf = np.zeros_like(bf)
        for ia, (c, p) in enumerate(zip(self.cth, self.phi)):
            sinth = np.sqrt(max((1.0 - c) * (1.0 + c), 0.0))
            cot = c / sinth if sinth > 0 else 0.0
            for i, (l, m) in enumerate(zip(lv, mv)):
                y = sph_harm(int(l), int(m), c, p)
                ang = m * cot * y
                if m < l:
                    ang = ang + np.sqrt((l - m) * (l + m + 1)) * np.exp(-1j * p) * sph_harm(int(l), int(m) + 1, c, p)
                rows = slice(i * Nr, (i + 1) * Nr); cols = slice(ia * nrad, (ia + 1) * nrad)
                # the reference stores the ADJOINT of the (point x function) blocks
                bf[rows, cols] = np.conj(y * f).T
                dr[rows, cols] = np.conj(y * d).T
                dph[rows, cols] = np.conj(1j * m * y * f).T
                dth[rows, cols] = np.conj(ang * f).T
                lf[rows, cols] = np.conj((l2 + 2 * d / r[:, None] - l * (l + 1) * f / (r * r)[:, None]) * y).T
        self.t = {"f": bf, "r": dr, "t": dth, "p": dph, "l": lf}

    def density(self, P0, grad, tau, lapl):
        """One spin channel; returns dict with rho, grho(3,N), kin (sum_c D_cc/s_c^2), lap."""
        P = P0[np.ix_(self.bf_ind, self.bf_ind)]
        t = self.t
        A, B = t["f"].real, t["f"].imag
        PvA, PvB = P @ A, P @ B
        out = {"rho": np.sum(PvA * A + PvB * B, axis=0)}
        if grad:
            out["grho"] = np.stack([2.0 * np.sum(PvA * t[k].real + PvB * t[k].imag, axis=0) / self.scale[c]
                                    for c, k in enumerate("rtp")])
        if tau or lapl:
            kin = 0.0
            for c, k in enumerate("rtp"):
                kin = kin + np.sum((P @ t[k].real) * t[k].real + (P @ t[k].imag) * t[k].imag, axis=0) / self.scale[c] ** 2
            out["tau"] = 0.5 * kin
            if lapl:
                out["lapl"] = 2.0 * (kin + np.sum(PvA * t["l"].real + PvB * t["l"].imag, axis=0))
        return out

    @staticmethod
    def _lda(H, v, f):
        H += (f.real * v) @ f.real.T + (f.imag * v) @ f.imag.T

    def _gga(self, H, gn):
        t = self.t
        gre = sum(gn[c] * t[k].real for c, k in enumerate("rtp"))
        gim = sum(gn[c] * t[k].imag for c, k in enumerate("rtp"))
        X = gre @ t["f"].real.T + gim @ t["f"].imag.T
        H += X + X.T

    def _lapl(self, H, v):
        t = self.t
        X = (t["f"].real * v) @ t["l"].real.T + (t["f"].imag * v) @ t["l"].imag.T
        H += X + X.T

    def fxc_block(self, vrho, w_gn=None, vtl=None, vl=None):
        nbf = len(self.bf_ind)
        H = np.zeros((nbf, nbf))
        self._lda(H, vrho * self.wtot, self.t["f"])
        if w_gn is not None:
            self._gga(H, w_gn)
        if vtl is not None:
            for c, k in enumerate("rtp"):
                self._lda(H, vtl * self.wtot / self.scale[c] ** 2, self.t[k])
        if vl is not None:
            self._lapl(H, vl * self.wtot)
        return H

    # ------------------------------------------------------------------ drivers
    def npoints(self):
        return self.b.radial.Nel() * len(self.wang) * len(self.b.radial.xq)

    def eval_density(self, Pa, Pb=None, grad=False, tau=False, lapl=False):
        """All elements; libxc layout (N x ncomp row-major == ncomp x N column-major)."""
        pol = Pb is not None
        res = {"rho": [], "sigma": [], "tau": [], "lapl": [], "w": []}
        self._store = []
        nel = ekin = 0.0
        for iel in range(self.b.radial.Nel()):
            self.compute_bf(iel)
            da = self.density(Pa, grad, tau, lapl)
            db = self.density(Pb, grad, tau, lapl) if pol else None
            self._store.append((da, db))
            w = self.wtot
            res["w"].append(w)
            if pol:
                res["rho"].append(np.stack([da["rho"], db["rho"]], axis=1))
                nel += np.sum(w * (da["rho"] + db["rho"]))
                if grad:
                    ga, gb = da["grho"], db["grho"]
                    res["sigma"].append(np.stack([np.sum(ga * ga, 0), np.sum(ga * gb, 0), np.sum(gb * gb, 0)], axis=1))
                if tau or lapl:
                    res["tau"].append(np.stack([da["tau"], db["tau"]], axis=1))
                    ekin += np.sum(w * (da["tau"] + db["tau"]))
                if lapl:
                    res["lapl"].append(np.stack([da["lapl"], db["lapl"]], axis=1))
            else:
                res["rho"].append(da["rho"][:, None])
                nel += np.sum(w * da["rho"])
                if grad:
                    res["sigma"].append(np.sum(da["grho"] ** 2, 0)[:, None])
                if tau or lapl:
                    res["tau"].append(da["tau"][:, None])
                    ekin += np.sum(w * da["tau"])
                if lapl:
                    res["lapl"].append(da["lapl"][:, None])
        out = {k: (np.concatenate(v) if v else None) for k, v in res.items()}
        out["Nel"], out["Ekin"] = nel, ekin
        return out

    def eval_fxc(self, n, exc, vrho, vsigma=None, vtau=None, vlapl=None, polarized=False, beta=True):
        """Assemble H (Ha, Hb) from functional output in libxc layout (uses the densities kept
        from the last eval_density); returns (Ha, Hb, Exc).  src/atomic/dftgrid.cpp:304-465."""
        Ha = np.zeros((n, n)); Hb = np.zeros((n, n)) if polarized else None
        Exc = 0.0
        off = 0
        for iel in range(self.b.radial.Nel()):
            self.compute_bf(iel)
            npt = len(self.wtot)
            sl = slice(off, off + npt); off += npt
            da, db = self._store[iel]
            w = self.wtot
            idx = np.ix_(self.bf_ind, self.bf_ind)
            if not polarized:
                Exc += np.sum(w * exc[sl] * da["rho"])
                gn = None
                if vsigma is not None:
                    gn = np.stack([2.0 * w * vsigma[sl, 0] * da["grho"][c] / self.scale[c] for c in range(3)])
                vtl = None
                if vtau is not None or vlapl is not None:
                    vtl = (0.5 * vtau[sl, 0] if vtau is not None else 0.0) + (2.0 * vlapl[sl, 0] if vlapl is not None else 0.0)
                Ha[idx] += self.fxc_block(vrho[sl, 0], gn, vtl, vlapl[sl, 0] if vlapl is not None else None)
            else:
                Exc += np.sum(w * exc[sl] * (da["rho"] + db["rho"]))
                for spin, (H, dd, do, s_same) in enumerate(((Ha, da, db, 0), (Hb, db, da, 2))):
                    if spin == 1 and not beta:
                        continue
                    gn = None
                    if vsigma is not None:
                        gn = np.stack([w * (2.0 * vsigma[sl, s_same] * dd["grho"][c] + vsigma[sl, 1] * do["grho"][c]) / self.scale[c]
                                       for c in range(3)])
                    vtl = None
                    if vtau is not None:   # reference: unrestricted branch is gated on do_mgga_t only (:426)
                        vtl = 0.5 * vtau[sl, spin] + (2.0 * vlapl[sl, spin] if vlapl is not None else 0.0)
                    H[idx] += self.fxc_block(vrho[sl, spin], gn, vtl, vlapl[sl, spin] if vlapl is not None else None)
        return Ha, Hb, Exc


class Diatomic3DGrid(AtomicDFTGrid):
    """General 3D diatomic grid (symmetry=0 path), src/diatomic/dftgrid.cpp:414-518 with
    eval_bf/eval_df of src/diatomic/basis.cpp:2120-2243: FE functions B(mu) (not B/r), scale
    factors h_mu = h_nu = Rh sqrt(sinh^2 mu + sin^2 nu), h_phi = Rh sinh mu sin nu, weight
    w_ang w_mu Rh^3 sinh mu (sinh^2 mu + sin^2 nu).  No Laplacian in the reference."""

    def compute_bf(self, iel):
        b = self.b
        lv, mv = b.lval, b.mval
        rb = b.radial
        xq = rb.xq
        mu = rb.fem.coord(xq, iel)
        wrad = rb.wq * rb.fem.scale(iel)
        f, d = rb.fem.eval_dnf(xq, 0, iel), rb.fem.eval_dnf(xq, 1, iel)
        nrad, nang, Nr = len(mu), len(self.wang), f.shape[1]
        a0, _ = rb.get_idx(iel)
        N = b.Nrad()
        self.bf_ind = np.array([N * ia + a0 + j for ia in range(len(lv)) for j in range(Nr)])
        sth = np.sqrt(1.0 - self.cth ** 2)
        sh = np.sinh(mu)
        h = (b.Rhalf * np.sqrt(sh[None, :] ** 2 + sth[:, None] ** 2)).ravel()
        self.scale = np.stack([h, h, (b.Rhalf * sh[None, :] * sth[:, None]).ravel()])
        self.wtot = (self.wang[:, None] * (wrad * b.Rhalf ** 3 * sh)[None, :] * (sh[None, :] ** 2 + sth[:, None] ** 2)).ravel()
        nbf = len(self.bf_ind)
        bf = np.zeros((nbf, nang * nrad), complex); dr = np.zeros_like(bf); dth = np.zeros_like(bf); dph = np.zeros_like(bf)
        for ia, (c, p) in enumerate(zip(self.cth, self.phi)):
            sinth = np.sqrt(max((1.0 - c) * (1.0 + c), 0.0))
            cot = c / sinth if sinth > 0 else 0.0
            for i, (l, m) in enumerate(zip(lv, mv)):
                y = sph_harm(int(l), int(m), c, p)
                ang = m * cot * y
                if m < l:
                    ang = ang + np.sqrt((l - m) * (l + m + 1)) * np.exp(-1j * p) * sph_harm(int(l), int(m) + 1, c, p)
                rows = slice(i * Nr, (i + 1) * Nr); cols = slice(ia * nrad, (ia + 1) * nrad)
                bf[rows, cols] = np.conj(y * f).T
                dr[rows, cols] = np.conj(y * d).T
                dph[rows, cols] = np.conj(1j * m * y * f).T
                dth[rows, cols] = np.conj(ang * f).T
        self.t = {"f": bf, "r": dr, "t": dth, "p": dph, "l": np.zeros_like(bf)}

    def eval_density(self, Pa, Pb=None, grad=False, tau=False, lapl=False):
        if lapl:
            raise ValueError("Laplacian not implemented.")   # src/diatomic/dftgrid.cpp:503-505
        ex = self.b.expand_boundaries
        return super().eval_density(ex(Pa), None if Pb is None else ex(Pb), grad, tau, False)

    def eval_fxc(self, exc, vrho, vsigma=None, vtau=None, polarized=False, beta=True):
        Ha, Hb, E = super().eval_fxc(self.b.Ndummy(), exc, vrho, vsigma, vtau, None, polarized, beta)
        rb = self.b.remove_boundaries
        return rb(Ha), (rb(Hb) if Hb is not None else None), E
