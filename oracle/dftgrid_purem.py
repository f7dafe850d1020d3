"""Diatomic pure-m DFT grid (mu x nu, phi integrated analytically): density, gradient, tau,
Laplacian and XC-matrix assembly.  Oracle restatement (numpy, materialised per-m tables as the
reference builds them); test infrastructure only.

Follows src/diatomic/dftgrid_purem.cpp: ctor :30-86, compute_bf :87-200, update_density
:212-442, eval_Fxc :474-657, driver :675-744.  The reference walks one radial point at a time;
here the same arithmetic is vectorised over the radial points of an element and the points are
ordered (element, nu point, mu point) -- a pure relabelling of the quadrature sum.
"""
import numpy as np

from . import fem
from .dftgrid_atomic import sph_harm


class PureMDFTGrid:
    def __init__(self, basis, lang):
        self.b = basis
        self.cth, self.wang = fem.chebyshev(lang)
        lv, mv = basis.lval, basis.mval
        self.mlist = sorted(set(int(m) for m in mv))
        self.shells = {m: [i for i in range(len(lv)) if mv[i] == m] for m in self.mlist}
        sth = np.sqrt(np.maximum((1 - self.cth) * (1 + self.cth), 0.0))
        cot = np.where(sth > 0, self.cth / np.where(sth > 0, sth, 1.0), 0.0)
        self.Y, self.dY = {}, {}
        for i, (l, m) in enumerate(zip(lv, mv)):
            l, m = int(l), int(m)
            y = np.real(sph_harm(l, m, self.cth, 0.0))
            dy = m * cot * y
            if m < l:
                dy = dy + np.sqrt((l - m) * (l + m + 1.0)) * np.real(sph_harm(l, m + 1, self.cth, 0.0))
            self.Y[i], self.dY[i] = y, dy
        self.sth = sth

    def npoints(self):
        return self.b.radial.Nel() * len(self.cth) * len(self.b.radial.xq)

    def compute_bf(self, iel):
        b = self.b
        rb = b.radial
        xq = rb.xq
        mu = rb.fem.coord(xq, iel)
        wr = rb.wq * rb.fem.scale(iel)
        B0, B1, B2 = (rb.fem.eval_dnf(xq, k, iel) for k in range(3))
        Nr = B0.shape[1]
        a0, _ = rb.get_idx(iel)
        N = b.Nrad()
        Rh = b.Rhalf
        sh, ch = np.sinh(mu), np.cosh(mu)
        nang, nrad = len(self.cth), len(mu)
        h = Rh * np.sqrt(sh[None, :] ** 2 + self.sth[:, None] ** 2)          # (nang, nrad)
        hphi = Rh * sh[None, :] * self.sth[:, None]
        self.h = h.ravel()
        self.inv_h2 = np.where(self.h > 0, 1.0 / self.h ** 2, 0.0)
        self.inv_hphi2 = np.where(hphi.ravel() > 0, 1.0 / hphi.ravel() ** 2, 0.0)
        self.wtot = (2.0 * np.pi * self.wang[:, None] * wr[None, :] * Rh ** 3 * sh[None, :]
                     * (sh[None, :] ** 2 + self.sth[:, None] ** 2)).ravel()
        coth = ch / sh
        self.blocks = {}
        for m in self.mlist:
            idx, bf, dr, dth, lf = [], [], [], [], []
            for i in self.shells[m]:
                l = int(b.lval[i])
                radop = B2 + coth[:, None] * B1 - (l * (l + 1) + m * m / sh[:, None] ** 2) * B0      # (nrad, Nr)
                for j in range(Nr):
                    idx.append(N * i + a0 + j)
                    bf.append(np.outer(self.Y[i], B0[:, j]).ravel())
                    dr.append(np.outer(self.Y[i], B1[:, j]).ravel())
                    dth.append(np.outer(self.dY[i], B0[:, j]).ravel())
                    lf.append(np.outer(self.Y[i], radop[:, j]).ravel() * self.inv_h2)
            self.blocks[m] = (np.array(idx), np.array(bf), np.array(dr), np.array(dth), np.array(lf))

    def density(self, P, grad, tau, lapl):
        out = {"rho": 0.0, "grho": np.zeros((2, len(self.wtot)))}
        kin = 0.0
        lap = 0.0
        for m, (idx, bf, dr, dth, lf) in self.blocks.items():
            Pb = P[np.ix_(idx, idx)]
            Pv = Pb @ bf
            rho_m = np.sum(Pv * bf, axis=0)
            out["rho"] = out["rho"] + rho_m
            if grad:
                out["grho"][0] += 2.0 * np.sum(Pv * dr, axis=0)
                out["grho"][1] += 2.0 * np.sum(Pv * dth, axis=0)
            if tau or lapl:
                kr = np.sum((Pb @ dr) * dr, axis=0) * self.inv_h2
                kt = np.sum((Pb @ dth) * dth, axis=0) * self.inv_h2
                kin = kin + kr + kt + m * m * rho_m * self.inv_hphi2
                if lapl:
                    lap = lap + 2.0 * np.sum(Pv * lf, axis=0)
        if grad:
            out["grho"] = out["grho"] / self.h
        if tau or lapl:
            out["tau"] = 0.5 * kin
            if lapl:
                out["lapl"] = lap + 2.0 * kin
        return out

    def fxc_into(self, H, vrho, gn=None, vtl=None, vl=None):
        for m, (idx, bf, dr, dth, lf) in self.blocks.items():
            Hm = (bf * (vrho * self.wtot)) @ bf.T
            if gn is not None:
                X = (gn[0] * dr + gn[1] * dth) @ bf.T
                Hm += X + X.T
            if vtl is not None:
                wr = vtl * self.wtot * self.inv_h2
                Hm += (dr * wr) @ dr.T + (dth * wr) @ dth.T
                if m != 0:
                    Hm += (bf * (vtl * self.wtot * self.inv_hphi2 * m * m)) @ bf.T
            if vl is not None:
                X = (bf * (vl * self.wtot)) @ lf.T
                Hm += X + X.T
            H[np.ix_(idx, idx)] += Hm

    def eval_density(self, Pa, Pb=None, grad=False, tau=False, lapl=False):
        b = self.b
        Pa = b.expand_boundaries(Pa)
        pol = Pb is not None
        if pol:
            Pb = b.expand_boundaries(Pb)
        res = {"rho": [], "sigma": [], "tau": [], "lapl": [], "w": []}
        self._store = []
        nel = ekin = 0.0
        for iel in range(b.radial.Nel()):
            self.compute_bf(iel)
            da = self.density(Pa, grad, tau, lapl)
            db = self.density(Pb, grad, tau, lapl) if pol else None
            self._store.append((da, db))
            w = self.wtot
            res["w"].append(w)
            comps = [da] + ([db] if pol else [])
            res["rho"].append(np.stack([d["rho"] for d in comps], axis=1))
            nel += sum(np.sum(w * d["rho"]) for d in comps)
            if grad:
                if pol:
                    ga, gb = da["grho"], db["grho"]
                    res["sigma"].append(np.stack([np.sum(ga * ga, 0), np.sum(ga * gb, 0), np.sum(gb * gb, 0)], axis=1))
                else:
                    res["sigma"].append(np.sum(da["grho"] ** 2, 0)[:, None])
            if tau or lapl:
                res["tau"].append(np.stack([d["tau"] for d in comps], axis=1))
                ekin += sum(np.sum(w * d["tau"]) for d in comps)
            if lapl:
                res["lapl"].append(np.stack([d["lapl"] for d in comps], axis=1))
        out = {k: (np.concatenate(v) if v else None) for k, v in res.items()}
        out["Nel"], out["Ekin"] = nel, ekin
        return out

    def eval_fxc(self, exc, vrho, vsigma=None, vtau=None, vlapl=None, polarized=False, beta=True):
        b = self.b
        nd = b.Ndummy()
        Ha = np.zeros((nd, nd)); Hb = np.zeros((nd, nd)) if polarized else None
        Exc = 0.0
        off = 0
        for iel in range(b.radial.Nel()):
            self.compute_bf(iel)
            npt = len(self.wtot)
            sl = slice(off, off + npt); off += npt
            da, db = self._store[iel]
            w = self.wtot
            if not polarized:
                Exc += np.sum(w * exc[sl] * da["rho"])
                gn = None
                if vsigma is not None:
                    gn = 2.0 * w * vsigma[sl, 0] / self.h * da["grho"]
                vtl = None
                if vtau is not None or vlapl is not None:
                    vtl = (0.5 * vtau[sl, 0] if vtau is not None else 0.0) + (2.0 * vlapl[sl, 0] if vlapl is not None else 0.0)
                self.fxc_into(Ha, vrho[sl, 0], gn, vtl, vlapl[sl, 0] if vlapl is not None else None)
            else:
                Exc += np.sum(w * exc[sl] * (da["rho"] + db["rho"]))
                for spin, (H, dd, do, s_same) in enumerate(((Ha, da, db, 0), (Hb, db, da, 2))):
                    if spin == 1 and not beta:
                        continue
                    gn = None
                    if vsigma is not None:
                        gn = w * (2.0 * vsigma[sl, s_same] * dd["grho"] + vsigma[sl, 1] * do["grho"]) / self.h
                    vtl = None
                    if vtau is not None or vlapl is not None:
                        vtl = (0.5 * vtau[sl, spin] if vtau is not None else 0.0) + (2.0 * vlapl[sl, spin] if vlapl is not None else 0.0)
                    self.fxc_into(H, vrho[sl, spin], gn, vtl, vlapl[sl, spin] if vlapl is not None else None)
        Ha = b.remove_boundaries(Ha)
        if polarized:
            Hb = b.remove_boundaries(Hb)
        return Ha, Hb, Exc
