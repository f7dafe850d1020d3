"""Spherically averaged atom: Coulomb and exchange.  Oracle restatement of
src/sadatom/basis.cpp:186-207 (coulomb, L = 0 only) and :209-312 (exchange with the m-averaged
squared Gaunt coupling).  Radial caches are those of the atomic oracle.  Test infrastructure only."""
import numpy as np

from .gaunt import Gaunt


class SadatomBasis:
    def __init__(self, atomic_basis, lmax):
        """atomic_basis: oracle.atomic.TwoDBasis with compute_tei() done for N_L >= 2*lmax+1."""
        self.b = atomic_basis
        self.lmax = lmax

    def coulomb(self, P_in):
        return 4.0 * np.pi * self.b._assemble_J(0, P_in)

    def rs_exchange(self, P):
        """src/sadatom/basis.cpp:314-420: same loops with the range-separated caches of the atomic oracle
        (compute_yukawa / compute_erfc) and prefactor 4 pi lambda (Yukawa) or 4 pi mu/(2L+1) (erfc), :385."""
        b = self.b
        if getattr(b, "yukawa", True) is False:
            return self.exchange(P, Lfac=lambda L: 4.0 * np.pi * b.lam / (2 * L + 1), assemble=b._assemble_K_pairwise)
        save = (b.disjoint_L, b.disjoint_m1L, b.prim_chol)
        b.disjoint_L, b.disjoint_m1L, b.prim_chol = b.disjoint_iL, b.disjoint_kL, b.rs_chol
        try:
            return self.exchange(P, Lfac=lambda L: 4.0 * np.pi * b.lam)
        finally:
            b.disjoint_L, b.disjoint_m1L, b.prim_chol = save

    def exchange(self, P, Lfac=None, assemble=None):
        g = Gaunt()
        assemble = assemble or self.b._assemble_K
        Lfac = Lfac or (lambda L: 4.0 * np.pi / (2 * L + 1))
        gmax = self.lmax
        N = self.b.Nrad()
        K = [np.zeros((N, N)) for _ in range(gmax + 1)]
        for lout in range(gmax + 1):
            Prad = {}
            for lin in range(gmax + 1):
                if np.linalg.norm(P[lin]) == 0.0:
                    continue
                Lmin, Lmax = abs(lin - lout), lin + lout
                tot = np.zeros(Lmax + 1)
                for mout in range(-lout, lout + 1):
                    for min_ in range(-lin, lin + 1):
                        M = mout - min_
                        for L in range(Lmin, Lmax + 1):
                            tot[L] += g.coeff(lout, mout, L, M, lin) ** 2
                tot /= 2 * lout + 1
                for L in range(Lmin, Lmax + 1):
                    if tot[L] == 0.0:
                        continue
                    Prad[L] = Prad.get(L, 0.0) + (Lfac(L) * tot[L]) * P[lin]
            for L, PL in sorted(Prad.items()):
                K[lout] -= assemble(L, PL)
        return K


class SadatomDFTGrid:
    """Radial-only DFT quadrature of the spherically averaged atom: oracle restatement of
    src/sadatom/dftgrid.cpp (compute_bf :464-486, update_density :45-125 restricted / :127-241 unrestricted,
    eval_Fxc :256-328 / :330-460, DFTGrid::eval_Fxc :505-653).  Densities and Fock matrices are per-l cubes
    (lists of Nrad x Nrad matrices); point index = iel*nquad + iquad; libxc layout (point-major, spins
    interleaved) for the per-point arrays."""

    def __init__(self, atomic_basis, lmax):
        from .dftgrid_atomic import AtomicDFTGrid
        self.b = atomic_basis
        self._nl = lmax + 1
        self._rad = AtomicDFTGrid(atomic_basis, 1, 1)._radial      # RadialBasis.cpp:868-926 tables

    def _el(self, iel):
        r, wrad, f, d, l2 = self._rad(iel)
        a, b = self.b.radial.get_idx(iel)
        return r, wrad, 4.0 * np.pi * wrad * r * r, f.T, d.T, l2.T, a, b     # tables nbf_el x npts

    def npoints(self):
        return self.b.radial.Nel() * len(self.b.radial.xq)

    def eval_density(self, Pa, Pb=None, grad=False, tau=False, lapl=False):
        pol = Pb is not None
        cubes = [Pa, Pb] if pol else [Pa]
        ns = len(cubes)
        N = self.npoints()
        out = {"rho": np.zeros((N, ns)), "w": np.zeros(N)}
        g = np.zeros((N, ns))
        if tau:
            out["tau"] = np.zeros((N, ns))
        if lapl:
            out["lapl"] = np.zeros((N, ns))
            out["lapl_scale"] = np.zeros((N, ns))     # sum of |summands|: conditioning of the Laplacian sum
        npt = len(self.b.radial.xq)
        for iel in range(self.b.radial.Nel()):
            r, wrad, wtot, bf, bfr, bfr2, a, b = self._el(iel)
            sl = slice(iel * npt, (iel + 1) * npt)
            out["w"][sl] = wtot
            for s, cube in enumerate(cubes):
                Pc = [np.asarray(Pl)[a:b + 1, a:b + 1] for Pl in cube]
                P = sum(Pc)
                Plm = sum(l * (l + 1) * Pc[l] for l in range(len(Pc)))
                Pv = P @ bf
                out["rho"][sl, s] = np.sum(Pv * bf, axis=0)
                if grad:
                    g[sl, s] = 2.0 * np.sum(Pv * bfr, axis=0)
                if tau or lapl:
                    Pvp = P @ bfr
                    t1 = np.sum(Pvp * bfr, axis=0)
                    if tau:
                        t2 = np.sum((Plm @ bf) * bf, axis=0) / (r * r)
                        out["tau"][sl, s] = 0.5 * (t1 + np.maximum(t2, 0.0))
                    if lapl:
                        t2l, t3l = 2.0 * np.sum(Pv * bfr2, axis=0), 4.0 * np.sum(Pv * bfr, axis=0) / r
                        out["lapl"][sl, s] = 2.0 * t1 + t2l + t3l
                        aP = np.abs(P)
                        out["lapl_scale"][sl, s] = (2.0 * np.sum((aP @ np.abs(bfr)) * np.abs(bfr), axis=0)
                                                    + 2.0 * np.sum((aP @ np.abs(bf)) * np.abs(bfr2), axis=0)
                                                    + 4.0 * np.sum((aP @ np.abs(bf)) * np.abs(bfr), axis=0) / r)
        if grad:
            out["grho"] = g
            out["sigma"] = (np.stack([g[:, 0] ** 2, g[:, 0] * g[:, 1], g[:, 1] ** 2], axis=1) if pol else g ** 2)
        out["Nel"] = float(np.sum(out["w"] * out["rho"].sum(axis=1)))
        self._g, self._pol, self._rho, self._w = g, pol, out["rho"], out["w"]
        return out

    def eval_fxc(self, exc, vrho, vsigma=None, vtau=None, vlapl=None, beta=True):
        """Returns (Ha cube, Hb cube or None, Exc); uses the density and gradient of the last eval_density
        call.  exc: energy density per particle, Exc = sum_p w exc rho_total (DFTGridWorkerBase::eval_Exc)."""
        pol = self._pol
        N = self.b.Nrad()
        L = self._nl
        Ha = [np.zeros((N, N)) for _ in range(L)]
        Hb = [np.zeros((N, N)) for _ in range(L)] if (pol and beta) else None
        npt = len(self.b.radial.xq)
        vrho = np.asarray(vrho).reshape(self.npoints(), -1)
        rs = lambda v: None if v is None else np.asarray(v).reshape(self.npoints(), -1)
        vsigma, vtau, vlapl = rs(vsigma), rs(vtau), rs(vlapl)
        for iel in range(self.b.radial.Nel()):
            r, wrad, wtot, bf, bfr, bfr2, a, b = self._el(iel)
            sl = slice(iel * npt, (iel + 1) * npt)
            for s, Hc in enumerate([Ha, Hb] if pol else [Ha]):
                if Hc is None:
                    continue
                H = (bf * (vrho[sl, s] * wtot)) @ bf.T
                Hl = np.zeros_like(H)
                if vsigma is not None:
                    if pol:
                        gs, go = self._g[sl, s], self._g[sl, 1 - s]
                        gr = wtot * (2.0 * vsigma[sl, 2 * s] * gs + vsigma[sl, 1] * go)
                    else:
                        gr = 2.0 * wtot * vsigma[sl, 0] * self._g[sl, 0]
                    if vlapl is not None:
                        gr = gr + 2.0 * vlapl[sl, s] * r * (wrad * 4.0 * np.pi)
                    X = (bfr * gr) @ bf.T            # increment_gga: H += X + X^T
                    H += X + X.T
                if vtau is not None or vlapl is not None:
                    vtl = np.zeros(npt)
                    if vtau is not None:
                        vtl = vtl + 0.5 * vtau[sl, s]
                    if vlapl is not None:
                        vtl = vtl + 2.0 * vlapl[sl, s]
                    H += (bfr * (vtl * wtot)) @ bfr.T
                    if vtau is not None:
                        Hl += (bf * (vtau[sl, s] * 0.5 * wrad * 4.0 * np.pi)) @ bf.T
                    if vlapl is not None:
                        Y = (bf * (vlapl[sl, s] * wtot)) @ bfr2.T      # increment_mgga_lapl: H += Y + Y^T
                        H += Y + Y.T
                for l in range(L):
                    Hc[l][a:b + 1, a:b + 1] += H + l * (l + 1) * Hl
        Exc = float(np.sum(self._w * np.asarray(exc).reshape(-1) * self._rho.sum(axis=1))) if exc is not None else 0.0
        return Ha, Hb, Exc
