"""Spherically averaged atom: Coulomb and exchange.  Oracle restatement of
src/sadatom/basis.cpp:186-207 (coulomb, L = 0 only) and :209-312 (exchange with the m-averaged
squared Gaunt coupling).  Radial caches are those of the atomic oracle.  Test infrastructure only."""
import numpy as np

from . import fem as fem_mod
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


class SapTable:
    """Radial effective-potential ("SAP") table of the spherically averaged atom: oracle restatement of
    src/sadatom/main.cpp:55-107 (effective_potential_table) and the pieces it calls in src/sadatom/basis.cpp:
    radii :564-578, quadrature_weights :486-499, electron_density :683-718 (nucleus value via
    libhelfem/src/RadialBasis.cpp:962-977), electron_density_gradient :912-936, electron_density_laplacian
    :938-964, kinetic_energy_density :966-1017, coulomb_screening :501-562 (in-element potential
    libhelfem/src/quadrature.cpp:251-292), xc_screening :1019-1179 for LDA exchange (the SAP functional,
    src/general/sap.h:40-43).  Point 0 is the nucleus; points 1.. are (element, quadrature node).
    Columns: r, rho, grad rho, lapl rho, tau, v_coul, v_xc, quadrature weight, Z_eff = Z - (v_coul + v_xc).
    Test infrastructure only."""

    LDA_X_DENS_THRESHOLD = 1e-24    # libxc's default density threshold of XC_LDA_X

    def __init__(self, atomic_basis, Z):
        from .dftgrid_atomic import AtomicDFTGrid
        self.b = atomic_basis
        self.Z = Z
        self._rad = AtomicDFTGrid(atomic_basis, 1, 1)._radial

    def _blocks(self, Prad):
        rb = self.b.radial
        for iel in range(rb.Nel()):
            a, b = rb.get_idx(iel)
            r, wrad, f, d, l2 = self._rad(iel)
            yield iel, np.asarray(Prad)[a:b + 1, a:b + 1], r, wrad, f, d, l2

    def _assemble(self, per_el, first=0.0):
        return np.concatenate([[first]] + list(per_el))

    def radii(self):
        return self._assemble(r for _, _, r, *_ in self._blocks(np.zeros((self.b.Nrad(),) * 2)))

    def quadrature_weights(self):
        return self._assemble(w for _, _, _, w, *_ in self._blocks(np.zeros((self.b.Nrad(),) * 2)))

    def nuclear_density(self, Prad):
        """P_uv B_u'(0) B_v'(0) / (4 pi), RadialBasis.cpp:962-977, basis.cpp:479-481."""
        rb = self.b.radial
        der = rb.fem.eval_dnf(np.array([-1.0]), 1, 0)
        a, b = rb.get_idx(0)
        return (der @ np.asarray(Prad)[a:b + 1, a:b + 1] @ der.T).item() / (4.0 * np.pi)

    def electron_density(self, Prad):
        return self._assemble((np.sum((f @ P) * f, axis=1) for _, P, r, w, f, d, l2 in self._blocks(Prad)),
                              4.0 * np.pi * self.nuclear_density(Prad))

    def electron_density_gradient(self, Prad):
        return self._assemble(2.0 * np.sum((f @ P) * d, axis=1) for _, P, r, w, f, d, l2 in self._blocks(Prad))

    def electron_density_laplacian(self, Prad):
        return self._assemble(2.0 * (np.sum((d @ P) * d, axis=1) + np.sum((f @ P) * l2, axis=1))
                              + 4.0 * np.sum((f @ P) * d, axis=1) / r for _, P, r, w, f, d, l2 in self._blocks(Prad))

    def kinetic_energy_density(self, cube):
        P = sum(np.asarray(c) for c in cube)
        Pl = sum(l * (l + 1) * np.asarray(c) for l, c in enumerate(cube))
        rb = self.b.radial
        out = []
        for iel, Psub, r, w, f, d, l2 in self._blocks(P):
            a, b = rb.get_idx(iel)
            t1 = np.sum((d @ Psub) * d, axis=1)
            t2 = np.sum((f @ Pl[a:b + 1, a:b + 1]) * f, axis=1) / (r * r)
            out.append(0.5 * (t1 + np.maximum(t2, 0.0)))
        return self._assemble(out)

    def spherical_potential(self, iel):
        """V[ip, (i,j)] = (1/r_ip) int_rmin^r_ip B_i B_j dr + int_r_ip^rmax B_i B_j / r dr, each sub-interval
        integrated with the element's own rule mapped onto it (quadrature.cpp:251-292, :37-74)."""
        rb = self.b.radial
        x, wx = rb.xq, rb.wq
        rmin, rmax = rb.fem.begin(iel), rb.fem.end(iel)
        rmid0, rlen0 = 0.5 * (rmax + rmin), 0.5 * (rmax - rmin)
        r = rmid0 + rlen0 * x
        en = rb.fem.enabled(iel)
        x0 = rb.fem.x0
        nq = len(x)

        def seg(lo, hi, wfun):
            mid, ln = 0.5 * (hi + lo), 0.5 * (hi - lo)
            rs = mid + ln * x
            bf = fem_mod.lip_eval((rs - rmid0) / rlen0, x0, 0)[:, en]
            wp = wx * wfun(rs, hi) * ln
            return ((bf * wp[:, None]).T @ bf).reshape(-1, order="F")

        zero = np.array([seg(r[ip - 1] if ip else rmin, r[ip], lambda rr, R: np.full_like(rr, 1.0 / R)) for ip in range(nq)])
        minusone = np.array([seg(r[ip], r[ip + 1] if ip < nq - 1 else rmax, lambda rr, R: 1.0 / rr) for ip in range(nq)])
        V = np.zeros_like(zero)
        for ip in range(nq):
            V[ip] = (zero[:ip + 1] * r[:ip + 1, None]).sum(axis=0) / r[ip] + minusone[ip:].sum(axis=0)
        return V

    def coulomb_screening(self, Prad):
        """r * V_H(r) at the table points (0 at the nucleus), basis.cpp:501-562."""
        rb = self.b.radial
        Nel = rb.Nel()
        zero, minusone = np.zeros(Nel), np.zeros(Nel)
        for iel, Psub, *_ in self._blocks(Prad):
            zero[iel] = np.sum(Psub * rb.radial_integral(0, iel).T)
            minusone[iel] = np.sum(Psub * rb.radial_integral(-1, iel).T)
        zero = np.cumsum(zero)
        minusone = np.cumsum(minusone[::-1])[::-1]
        out = []
        for iel, Psub, r, *_ in self._blocks(Prad):
            V = self.spherical_potential(iel) @ Psub.reshape(-1, order="F")
            if iel > 0:
                V = V + zero[iel - 1] / r
            if iel != Nel - 1:
                V = V + minusone[iel + 1]
            out.append(V * r)
        return self._assemble(out)

    def xc_screening(self, Pa, Pb):
        """r * v_x^sigma(r) for LDA exchange, spin-polarised formula v_sigma = -(6 rho_sigma / pi)^(1/3)
        (what libxc's XC_LDA_X returns for XC_POLARIZED), densities divided by 4 pi (basis.cpp:1025-1027)."""
        r = self.radii()
        cols = []
        for P in (Pa, Pb):
            rho = self.electron_density(P) / (4.0 * np.pi)
            v = np.where(rho > self.LDA_X_DENS_THRESHOLD, -np.cbrt(6.0 * np.maximum(rho, 0.0) / np.pi), 0.0)
            cols.append(v * r)
        return np.stack(cols, axis=1)

    def table(self, Pl_a, Pl_b=None):
        """effective_potential_table, main.cpp:55-107; restricted when Pl_b is None (Pl_a = total per-l density)."""
        restricted = Pl_b is None
        Pl = [np.asarray(c) for c in Pl_a] if restricted else [np.asarray(a) + np.asarray(b) for a, b in zip(Pl_a, Pl_b)]
        P = sum(Pl)
        if restricted:
            v = self.xc_screening(P / 2, P / 2)
            vxc = 0.5 * (v[:, 0] + v[:, 1])
        else:
            vxc = self.xc_screening(sum(np.asarray(c) for c in Pl_a), sum(np.asarray(c) for c in Pl_b)).mean(axis=1)
        vcoul = self.coulomb_screening(P)
        cols = [self.radii(), self.electron_density(P), self.electron_density_gradient(P), self.electron_density_laplacian(P),
                self.kinetic_energy_density(Pl), vcoul, vxc, self.quadrature_weights(), self.Z - (vcoul + vxc)]
        return np.stack(cols, axis=1)
