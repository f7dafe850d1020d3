"""Batched spherically averaged atomic SCFs on the GPU -- the gen_sap_table workload (BASELINE.json configs[4]:
"Batched SAP potentials for all elements Z = 1-86, distributed by element").

Mirrors helfem::sadatom::scf::run_atomic_scf (src/sadatom/scf.cpp:50-350) for the spin-restricted, LDA-exchange
case the SAP potential is defined by (src/general/sap.h:40-43; src/sadatom/main.cpp --method=lda_x --pot=lda_x), with
the per-l occupations frozen like the reference's own sub-SCFs (src/diatomic/twodquadrature.cpp:506-560).  The Fock
build follows scf.cpp:145-283 line by line; what the reference evaluates one density at a time (or a few at once in
its batched builder, :313-350) runs here for ALL atoms of the batch in a handful of launches:

  * XC   hfq_grid_density / hfq_grid_fxc on a batch context: the atom index is laid out along the radial grid's
         "angular point" axis, so one pass of the separable grid engine yields every atom's density and Fock blocks
         (the LDA exchange itself is evaluated point-wise between the two calls, in the caller's role of libxc);
  * J    hfq_coulomb_radial_batch: one launch of the radial Coulomb kernel, one CTA per atom;
  * the solver around them (S^-1/2 transform, per-l eigenproblems, DIIS) is batched torch on the same device --
    plumbing: the reference delegates it to OpenOrbitalOptimizer.

All matrices stay on the device; the two-electron caches do not depend on Z and are shared by all atoms.
Multi-GPU: the elements are dealt round-robin to the ranks (replicas, no collective) -- `elements_of_rank`.
"""
import ctypes
import json
import os

import numpy as np

from . import Tables, TablesBasis, _check, lib

_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "ground_states.json")


def ground_state_occupations():
    """{Z: [n_s, n_p, n_d, n_f]} and {Z: symbol}: the reference's table of frozen per-l occupations
    (helfem_b200/data/ground_states.json, extracted by tests/golden/make_ground_states.py)."""
    d = json.load(open(_DATA))
    return {int(k): v for k, v in d["occ"].items()}, {int(k): v for k, v in d["symbol"].items()}


def elements_of_rank(zs, rank, world):
    """Distribution by element: round-robin over the heaviest-first list, so that every rank gets light and heavy atoms."""
    order = sorted(zs, reverse=True)
    return sorted(order[rank::world])


def shell_occupations(n_l, l, nmax):
    """Electrons of angular momentum l filled shell by shell (capacity 2 (2l+1)); the last shell may be partial."""
    occ = np.zeros(nmax)
    cap = 2 * (2 * l + 1)
    left, i = float(n_l), 0
    while left > 0:
        occ[i] = min(cap, left)
        left -= occ[i]
        i += 1
    return occ


class SadatomBatchSCF:
    def __init__(self, zs, occs=None, nelem=5, nnodes=15, Rmax=40.0, igrid=4, zexp=2.0, device=0):
        import torch
        self.torch = torch
        self.dev = torch.device("cuda", device)
        self.zs = list(zs)
        table, self.symbols = ground_state_occupations()
        self.occ_l = [list(occs[z]) if occs else table[z] for z in self.zs]
        self.nl = 4
        nb = self.nb = len(self.zs)
        # one-electron radial matrices from an lmax = 1 basis with Z = 1: blocks l = 0, 1 give T, T + 2 Tl
        t1 = Tables.atomic(1, 1, 0, nelem, nnodes, Rmax, igrid, zexp)
        N = self.N = t1.Nrad
        S, T, V = t1.one_electron()
        f64 = dict(dtype=torch.float64, device=self.dev)
        self.S = torch.tensor(S[:N, :N], **f64)
        self.T = torch.tensor(T[:N, :N], **f64)
        self.Tl = torch.tensor(0.5 * (T[N:2 * N, N:2 * N] - T[:N, :N]), **f64)
        self.V1 = torch.tensor(V[:N, :N], **f64)     # -1 * <1/r>
        # S^-1/2 (libhelfem/src/utils.cpp:121-158)
        d = 1.0 / torch.sqrt(torch.diagonal(self.S))
        w, U = torch.linalg.eigh(self.S * d[:, None] * d[None, :])
        self.X = (U * w.rsqrt()[None, :]) @ U.T * d[:, None]
        # the batch context: radial Coulomb + radial grid for all atoms at once
        self.tables = Tables.sadatom_batch(self.nl - 1, nb, nelem, nnodes, Rmax, igrid, zexp)
        self.basis = TablesBasis(self.tables, device=device)
        self.ctx = self.basis._context()
        _check(lib().hfq_grid_attach(self.ctx, 1, 1))
        self.npts = int(lib().hfq_grid_npoints(self.ctx))
        self.nel_fe = self.tables.Nel
        self.nquad = self.npts // (self.nel_fe * nb)
        self.w = torch.zeros(self.npts, **f64)
        # occupations (nb, nl, N)
        occ = np.zeros((nb, self.nl, N))
        for a, ol in enumerate(self.occ_l):
            for l in range(self.nl):
                occ[a, l] = shell_occupations(ol[l], l, N)
        self.occ = torch.tensor(occ, **f64)
        self.Z = torch.tensor([float(z) for z in self.zs], **f64)
        self.ll1 = torch.tensor([l * (l + 1.0) for l in range(self.nl)], **f64)
        # block-compact matrices of the grid calls: (nb * nl, N + 1, N), block = [:, :N, :] (column-major N x N)
        self.Pc = torch.zeros((nb * self.nl, N + 1, N), **f64)
        self.Hc = torch.zeros((nb * self.nl, N + 1, N), **f64)
        self.rho = torch.zeros(self.npts, **f64)
        self.launches = 0
        self._tab_tables = None
        self.timing = {"eigh": 0.0, "xc": 0.0, "coulomb": 0.0, "other": 0.0}

    # ---- the two native batched operators -----------------------------------------------------------------------
    def xc(self, Pl):
        """Pl: (nb, nl, N, N) per-l densities.  Returns XC (nb, nl, N, N), Exc (nb), Nel (nb): the reference's
        grid.eval_Fxc(x_func = 1) on the cube / 4 pi, result / 4 pi (src/sadatom/scf.cpp:167-171)."""
        torch = self.torch
        nb, nl, N = self.nb, self.nl, self.N
        angfac = 4.0 * np.pi
        self.Pc[:, :N, :] = (Pl / angfac).reshape(nb * nl, N, N).transpose(1, 2)
        nel, ekin = ctypes.c_double(), ctypes.c_double()
        _check(lib().hfq_grid_density(self.ctx, self.Pc.data_ptr(), N, None, 0, 0, self.rho.data_ptr(), None, None, None,
                                      self.w.data_ptr(), ctypes.byref(nel), ctypes.byref(ekin)))
        rho = self.rho
        cx = -0.75 * (3.0 / np.pi) ** (1.0 / 3.0)
        ok = rho >= 1e-12
        r13 = torch.where(ok, rho.clamp_min(0.0) ** (1.0 / 3.0), torch.zeros_like(rho))
        exc = cx * r13
        vrho = (4.0 / 3.0) * cx * r13
        e = ctypes.c_double()
        _check(lib().hfq_grid_fxc(self.ctx, 0, 1, None, vrho.data_ptr(), None, None, None, self.Hc.data_ptr(), N, None, 0,
                                  ctypes.byref(e)))
        self.launches += 12
        # per-atom integrals: point p = (element, atom, radial node)
        shape = (self.nel_fe, nb, self.nquad)
        wr = (self.w * rho).view(shape)
        Nel = wr.sum(dim=(0, 2))
        Exc = (wr * exc.view(shape)).sum(dim=(0, 2))
        XC = self.Hc[:, :N, :].transpose(1, 2).reshape(nb, nl, N, N) / angfac
        return XC, Exc, Nel

    def coulomb(self, Prad):
        """J_a = coulomb(Prad_a / 4 pi) (src/sadatom/scf.cpp:199), all atoms in one launch."""
        P = (Prad / (4.0 * np.pi)).contiguous()
        J = self.torch.empty_like(P)
        _check(lib().hfq_coulomb_radial_batch(self.ctx, P.data_ptr(), J.data_ptr(), self.nb, 1.0, None))
        self.launches += 1
        return J

    def local_potential(self, v):
        """Matrix of a local radial potential per atom: v (nb, nel_fe * nquad) at the grid's radial points ->
        (nb, N, N) with <B_i| v |B_j>, assembled by the grid engine (used for the initial guess)."""
        torch = self.torch
        nb, N = self.nb, self.N
        vp = v.view(nb, self.nel_fe, self.nquad).permute(1, 0, 2).contiguous().view(-1)   # point = (element, atom, node)
        self.Pc.zero_()
        nel, ekin = ctypes.c_double(), ctypes.c_double()
        _check(lib().hfq_grid_density(self.ctx, self.Pc.data_ptr(), N, None, 0, 0, None, None, None, None, None,
                                      ctypes.byref(nel), ctypes.byref(ekin)))
        e = ctypes.c_double()
        _check(lib().hfq_grid_fxc(self.ctx, 0, 1, None, vp.data_ptr(), None, None, None, self.Hc.data_ptr(), N, None, 0,
                                  ctypes.byref(e)))
        H = self.Hc[:, :N, :].transpose(1, 2).reshape(nb, self.nl, N, N)
        return H[:, 0] / (4.0 * np.pi)

    def radii(self):
        """Radial quadrature points (nel_fe * nquad), element by element."""
        if self._tab_tables is None:
            self._tab_tables = Tables.sadatom(1, 0, self.tables.Nel)
        z = np.zeros((self.N, self.N))
        need = int(lib().hfq_sap_table(self._tab_tables._h, z.ctypes.data, None, 1, 0, None, 0))
        out = np.zeros(need)
        rows = int(lib().hfq_sap_table(self._tab_tables._h, z.ctypes.data, None, 1, 0, out.ctypes.data, need))
        return out.reshape(9, rows)[0, 1:]

    def guess_potential(self):
        """Thomas-Fermi screened nuclear attraction (Latter's fit of the TF function), never weaker than -1/r:
        the starting point of the SCF (the reference starts from its tabulated SAP potential, --iguess=2)."""
        torch = self.torch
        r = torch.tensor(self.radii(), dtype=torch.float64, device=self.dev)[None, :]
        Z = self.Z[:, None]
        x = r / (0.88534138 * Z ** (-1.0 / 3.0))
        sx = x.sqrt()
        phi = 1.0 / (1.0 + 0.02747 * sx + 1.243 * x - 0.1486 * x * sx + 0.2302 * x * x + 0.007298 * x * x * sx + 0.006944 * x ** 3)
        return -torch.maximum(Z * phi, torch.ones_like(phi)) / r

    # ---- SCF -------------------------------------------------------------------------------------------------
    def _eigh(self, Fo):
        """Eigenvectors of the (nb, nl, N, N) symmetric matrices, ascending eigenvalues: one CTA per matrix
        (hfq_syev_batch, cyclic Jacobi in shared memory)."""
        torch = self.torch
        V = Fo.contiguous().clone()
        W = torch.empty(V.shape[:-1], dtype=torch.float64, device=self.dev)
        _check(lib().hfq_syev_batch(V.data_ptr(), W.data_ptr(), self.N, V.shape[0] * V.shape[1],
                                    torch.cuda.current_stream().cuda_stream))
        self.launches += 1
        order = torch.argsort(W, dim=-1)
        return torch.gather(V, -1, order[..., None, :].expand_as(V))

    def _densities(self, F):
        """F: (nb, nl, N, N) -> per-l densities P_l = C occ C^T, C = X c (scf.cpp:119-130)."""
        t0 = self._tick()
        Fo = self.X.T @ F @ self.X
        C = self.X @ self._eigh(Fo)
        P = (C * self.occ[:, :, None, :]) @ C.transpose(-1, -2)
        self._tock("eigh", t0)
        return P

    def _tick(self):
        import time
        self.torch.cuda.synchronize()
        return time.perf_counter()

    def _tock(self, key, t0):
        import time
        self.torch.cuda.synchronize()
        self.timing[key] += time.perf_counter() - t0

    def run(self, maxit=150, conv=1e-10, errtol=1e-7, verbose=False, damp_above=0.3):
        """damp_above: an atom whose largest commutator element exceeds it takes a damped Roothaan step (30 % of the
        new density) instead of the DIIS extrapolation."""
        torch = self.torch
        nb, nl, N = self.nb, self.nl, self.N
        H0 = (self.T[None, None] + self.ll1[None, :, None, None] * self.Tl[None, None]
              + self.Z[:, None, None, None] * self.V1[None, None])
        Vg = self.local_potential(self.guess_potential())
        Pl = self._densities(self.T[None, None] + self.ll1[None, :, None, None] * self.Tl[None, None] + Vg[:, None])
        hist_F, hist_e = [], []
        Eold = torch.zeros(nb, dtype=torch.float64, device=self.dev)
        done = torch.zeros(nb, dtype=torch.bool, device=self.dev)
        diis_on = torch.zeros(nb, dtype=torch.bool, device=self.dev)   # latches once an atom is close; dropped if it strays far
        best = torch.full((nb,), float("inf"), dtype=torch.float64, device=self.dev)
        restart = torch.zeros(nb, dtype=torch.long, device=self.dev)
        hist_it = []
        for it in range(maxit):
            Prad = Pl.sum(dim=1)
            t0 = self._tick()
            XC, Exc, Nel = self.xc(Pl)
            self._tock("xc", t0)
            t0 = self._tick()
            J = self.coulomb(Prad)
            self._tock("coulomb", t0)
            Ekin = (Pl * self.T).sum(dim=(1, 2, 3)) + (self.ll1[None, :] * (Pl * self.Tl).sum(dim=(2, 3))).sum(dim=1)
            Enuc = self.Z * (Prad * self.V1).sum(dim=(1, 2))
            Ecoul = 0.5 * (Prad * J).sum(dim=(1, 2))
            E = Ekin + Enuc + Ecoul + Exc
            F = H0 + J[:, None] + XC
            err = F @ Pl @ self.S - self.S @ Pl @ F
            emax = err.abs().amax(dim=(1, 2, 3))
            done = ((E - Eold).abs() < conv * E.abs().clamp_min(1.0)) & (emax < errtol)
            if verbose:
                print("it %3d  not converged %3d  max err %.2e" % (it, int((~done).sum()), float(emax.max())))
            self.energies = {"E": E, "Ekin": Ekin, "Enuc": Enuc, "Coulomb": Ecoul, "XC": Exc, "Nel": Nel}
            self.Pl, self.iterations, self.last_emax = Pl, it + 1, emax
            if bool(done.all()):
                break
            Eold = E
            # DIIS per atom, batched
            hist_F.append(F)
            hist_e.append(err.reshape(nb, -1))
            hist_F, hist_e, hist_it = hist_F[-10:], hist_e[-10:], (hist_it + [it])[-10:]
            m = len(hist_F)
            ev = torch.stack(hist_e, dim=1)                      # (nb, m, n)
            B = torch.zeros((nb, m + 1, m + 1), dtype=torch.float64, device=self.dev)
            B[:, :m, :m] = ev @ ev.transpose(1, 2)
            # per-atom DIIS restart: entries recorded before the atom's restart iteration are priced out
            best = torch.minimum(best, emax)
            restart = torch.where(diis_on & (emax > 4.0 * best) & (emax > 1e-5), torch.full_like(restart, it), restart)
            best = torch.where(restart == it, emax, best)
            old = torch.tensor(hist_it, device=self.dev)[None, :] < restart[:, None]        # (nb, m)
            B[:, :m, :m] += torch.diag_embed(old.to(torch.float64) * 1e30)
            B[:, m, :m] = -1.0
            B[:, :m, m] = -1.0
            rhs = torch.zeros((nb, m + 1, 1), dtype=torch.float64, device=self.dev)
            rhs[:, m] = -1.0
            dg = B[:, :m, :m].diagonal(dim1=1, dim2=2)
            scale = torch.where(dg < 1e29, dg, torch.zeros_like(dg)).amax(dim=1).clamp_min(1e-300)
            B[:, :m, :m] /= scale[:, None, None]
            try:
                c = torch.linalg.solve(B, rhs)[:, :m, 0]
                ok = torch.isfinite(c).all(dim=1)
            except Exception:
                c, ok = None, torch.zeros(nb, dtype=torch.bool, device=self.dev)
            diis_on = (diis_on | (emax < damp_above)) & (emax < 10.0 * damp_above)
            Fd = F
            far = ~diis_on
            if c is not None:
                mix = (torch.stack(hist_F, dim=1) * c[:, :, None, None, None]).sum(dim=1)
                Fd = torch.where((ok & ~far)[:, None, None, None], mix, F)
            Pn = self._densities(Fd)
            Pl = torch.where(far[:, None, None, None], 0.7 * Pl + 0.3 * Pn, Pn)
        self.converged = done
        return {k: v.cpu().numpy() for k, v in self.energies.items()}

    # ---- products --------------------------------------------------------------------------------------------
    def sap_table(self, a):
        """effective_potential_table of atom a (src/sadatom/main.cpp:55-107): (Nel * nquad + 1) x 9."""
        z = self.zs[a]
        lmax = max(l for l in range(self.nl) if self.occ_l[a][l] > 0)
        if self._tab_tables is None:
            # the table needs the radial basis only (the screening integrals are evaluated by quadrature): one
            # set of tables serves every atom, the nuclear charge enters through the last column alone
            self._tab_tables = Tables.sadatom(1, 0, self.tables.Nel)
        t = self._tab_tables
        Pl = np.ascontiguousarray(self.Pl[a, :lmax + 1].transpose(-1, -2).cpu().numpy())   # column-major blocks
        need = int(lib().hfq_sap_table(t._h, Pl.ctypes.data, None, lmax + 1, 1, None, 0))
        out = np.zeros(need)
        rows = int(lib().hfq_sap_table(t._h, Pl.ctypes.data, None, lmax + 1, 1, out.ctypes.data, need))
        tab = out.reshape(9, rows).T
        tab[:, 8] = z - (tab[:, 5] + tab[:, 6])     # Z_eff = Z - r (V_H + V_xc), src/sadatom/main.cpp:98
        return tab

    def write_results(self, directory):
        """result_<El>.dat per atom, the reference's raw-ascii layout (io::write_raw_ascii, src/general/eigen_io.h:92-101:
        scientific, 16 digits after the point, field width 24, one table row per line)."""
        os.makedirs(directory, exist_ok=True)
        paths = []
        for a, z in enumerate(self.zs):
            tab = self.sap_table(a)
            path = os.path.join(directory, "result_%s.dat" % self.symbols[z])
            with open(path, "w") as f:
                for row in tab:
                    f.write("".join(" %24.16e" % v for v in row) + "\n")
            paths.append(path)
        return paths
