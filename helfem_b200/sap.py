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
        self.profile = False
        self.warm_start = os.environ.get("HFQ_SAP_WARM_START", "1") != "0"
        self._vprev = {}

    # ---- the two native batched operators -----------------------------------------------------------------------
    def xc(self, Pl):
        """Pl: (nb, nl, N, N) per-l densities.  Returns XC (nb, nl, N, N), Exc (nb), Nel (nb): the reference's
        grid.eval_Fxc(x_func = 1) on the cube / 4 pi, result / 4 pi (src/sadatom/scf.cpp:167-171)."""
        torch = self.torch
        nb, nl, N = self.nb, self.nl, self.N
        angfac = 4.0 * np.pi
        self.Pc[:, :N, :] = (Pl / angfac).reshape(nb * nl, N, N).transpose(1, 2)
        # the grid engine works on its own stream: what torch has queued (the matrix just written) must be complete;
        # the library synchronises its stream before it returns, which orders the other direction
        torch.cuda.synchronize()
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
        torch.cuda.synchronize()      # vrho was produced on torch's stream
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
        # torch's default stream has the handle 0, which the C ABI reads as "the context's own stream": the call is
        # asynchronous on that stream, so it is fenced on both sides here
        self.torch.cuda.synchronize()
        _check(lib().hfq_coulomb_radial_batch(self.ctx, P.data_ptr(), J.data_ptr(), self.nb, 1.0, None))
        self.torch.cuda.synchronize()
        self.launches += 1
        return J

    def local_potential(self, v):
        """Matrix of a local radial potential per atom: v (nb, nel_fe * nquad) at the grid's radial points ->
        (nb, N, N) with <B_i| v |B_j>, assembled by the grid engine (used for the initial guess)."""
        torch = self.torch
        nb, N = self.nb, self.N
        vp = v.view(nb, self.nel_fe, self.nquad).permute(1, 0, 2).contiguous().view(-1)   # point = (element, atom, node)
        self.Pc.zero_()
        torch.cuda.synchronize()      # Pc and vp come from torch's stream, the grid engine has its own
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
        """Eigenvalues (ascending) and eigenvectors of the (..., N, N) symmetric matrices: one CTA per matrix
        (hfq_syev_batch, cyclic Jacobi in shared memory)."""
        torch = self.torch
        V = Fo.contiguous().clone()
        W = torch.empty(V.shape[:-1], dtype=torch.float64, device=self.dev)
        _check(lib().hfq_syev_batch(V.data_ptr(), W.data_ptr(), self.N, V.numel() // (self.N * self.N),
                                    torch.cuda.current_stream().cuda_stream))
        self.launches += 1
        W, order = torch.sort(W, dim=-1)
        return W, torch.gather(V, -1, order[..., None, :].expand_as(V))

    def _fermi(self, eps, ntot, kT, by_l):
        """Occupations g_l / (1 + exp((eps - mu) / kT)) of the (na, nl, N) orbital energies, g_l = 2 (2l + 1): one
        chemical potential per l-block holding ntot[:, l] electrons (by_l) or one per atom holding ntot.sum(1)."""
        torch = self.torch
        cap = (2.0 * (2.0 * torch.arange(self.nl, device=self.dev, dtype=torch.float64) + 1.0))[None, :, None]
        red = (lambda x: x.sum(dim=2, keepdim=True)) if by_l else (lambda x: x.sum(dim=(1, 2), keepdim=True))
        n = ntot[:, :, None] if by_l else ntot.sum(dim=1)[:, None, None]
        lo = eps.amin(dim=2, keepdim=True) - 1.0 if by_l else eps.amin(dim=(1, 2), keepdim=True) - 1.0
        hi = eps.amax(dim=2, keepdim=True) + 1.0 if by_l else eps.amax(dim=(1, 2), keepdim=True) + 1.0
        for _ in range(60):
            mu = 0.5 * (lo + hi)
            low = red(cap * torch.sigmoid((mu - eps) / kT)) < n
            lo, hi = torch.where(low, mu, lo), torch.where(low, hi, mu)
        return cap * torch.sigmoid((0.5 * (lo + hi) - eps) / kT)

    def _densities(self, F, idx=None, kT=0.0, auto=False):
        """F: (na, nl, N, N) Fock matrices of the atoms idx (all if None) -> per-l densities P_l = C occ C^T, C = X c
        (scf.cpp:119-130).  kT = 0: the frozen per-l occupations self.occ, lowest orbitals first; kT > 0:
        Fermi-Dirac occupations at that electronic temperature, per l-block with the frozen electron counts, or
        (auto) across the blocks with one chemical potential per atom."""
        t0 = self._tick()
        Fo = self.X.T @ F @ self.X
        # warm start: in the eigenvector basis of the previous Fock matrices of the same atoms the new ones are nearly
        # diagonal, and the Jacobi sweeps (the latency of an SCF iteration) drop from ~10 to 2-3
        key = None if idx is None else tuple(idx.shape)
        Vp = self._vprev.get(key) if self.warm_start else None
        if Vp is not None and Vp.shape == Fo.shape:
            W, V = self._eigh(Vp.transpose(-1, -2) @ Fo @ Vp)
            V = Vp @ V
        else:
            W, V = self._eigh(Fo)
        self._vprev = {key: V}
        C = self.X @ V
        occ = self.occ if idx is None else self.occ[idx]
        if kT > 0.0:
            occ = self._fermi(W, occ.sum(dim=2), kT, not auto)
        P = (C * occ[:, :, None, :]) @ C.transpose(-1, -2)
        self._tock("eigh", t0)
        return P, occ

    def _tick(self):
        """Phase timers (self.timing) synchronise the device around every phase: only with self.profile set."""
        import time
        if not self.profile:
            return 0.0
        self.torch.cuda.synchronize()
        return time.perf_counter()

    def _tock(self, key, t0):
        import time
        if not self.profile:
            return
        self.torch.cuda.synchronize()
        self.timing[key] += time.perf_counter() - t0

    def _solve(self, Pl, active, kT=0.0, auto=False, maxit=150, conv=1e-10, errtol=1e-7, beta=0.3, mh=8, verbose=False,
               patience=None):
        """Self-consistency by Pulay mixing of the per-l densities (residual = output density of the Fock matrix
        minus input density; with thermal occupations it carries the occupation mismatch too, which the orbital
        commutator [F, P] does not see).  Fock builds run for the whole batch; the eigenproblems, the mixing and the
        convergence test only for the atoms in `active` (the others keep their density).  An atom whose residual
        doubles drops its history and halves its mixing factor.  patience: give up on the stragglers when the number
        of converged atoms has not grown for that many iterations.  Returns (Pl, converged mask of the batch)."""
        torch = self.torch
        nb, nl, N = self.nb, self.nl, self.N
        idx = torch.nonzero(active).flatten()
        na = int(idx.numel())
        f64 = dict(dtype=torch.float64, device=self.dev)
        H0 = (self.T[None, None] + self.ll1[None, :, None, None] * self.Tl[None, None]
              + self.Z[:, None, None, None] * self.V1[None, None])
        hP, hR, hit = [], [], []
        start = torch.zeros(na, dtype=torch.long, device=self.dev)
        prev = torch.full((na,), float("inf"), **f64)
        bet = torch.full((na,), beta, **f64)
        Eold = torch.zeros(na, **f64)
        done = torch.zeros(na, dtype=torch.bool, device=self.dev)
        eye = torch.eye(mh, **f64)
        ndone, since = 0, 0
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
            self.energies = {"E": E, "Ekin": Ekin, "Enuc": Enuc, "Coulomb": Ecoul, "XC": Exc, "Nel": Nel}
            F = (H0 + J[:, None] + XC)[idx]
            Pa = Pl[idx]
            Pout, occ = self._densities(F, idx, kT, auto)
            R = Pout - Pa
            rmax = R.abs().amax(dim=(1, 2, 3))
            err = F @ Pa @ self.S - self.S @ Pa @ F
            emax = torch.maximum(err.abs().amax(dim=(1, 2, 3)), rmax)
            Ea = E[idx]
            done = ((Ea - Eold).abs() < conv * Ea.abs().clamp_min(1.0)) & (emax < errtol)
            self.total_iterations += 1
            self.last_emax[idx] = emax
            self._last_occ = occ
            if verbose:
                print("it %3d  kT %.3g  not converged %3d  max err %.2e" % (it, kT, int((~done).sum()), float(emax.max())))
            if bool(done.all()):
                break
            nd = int(done.sum())
            ndone, since = (nd, 0) if nd > ndone else (ndone, since + 1)
            if patience and ndone > 0 and since >= patience:
                break
            Eold = Ea
            worse = rmax > 2.0 * prev
            start = torch.where(worse, torch.full_like(start, it), start)
            bet = torch.where(worse, (bet * 0.5).clamp_min(0.02), (bet * 1.1).clamp_max(beta))
            prev = rmax
            hP.append(Pa)
            hR.append(R.reshape(na, -1))
            hP, hR, hit = hP[-mh:], hR[-mh:], (hit + [it])[-mh:]
            m = len(hP)
            rv = torch.stack(hR, dim=1)                          # (na, m, n)
            B = torch.zeros((na, m + 1, m + 1), **f64)
            B[:, :m, :m] = rv @ rv.transpose(1, 2)
            old = torch.tensor(hit, device=self.dev)[None, :] < start[:, None]       # entries before the atom's restart
            B[:, :m, :m] += torch.diag_embed(old.to(torch.float64) * 1e30)
            dg = B[:, :m, :m].diagonal(dim1=1, dim2=2)
            scale = torch.where(dg < 1e29, dg, torch.zeros_like(dg)).amax(dim=1).clamp_min(1e-300)
            B[:, :m, :m] /= scale[:, None, None]
            B[:, :m, :m] += 1e-12 * eye[None, :m, :m]
            B[:, m, :m] = -1.0
            B[:, :m, m] = -1.0
            rhs = torch.zeros((na, m + 1, 1), **f64)
            rhs[:, m] = -1.0
            c = torch.linalg.solve(B, rhs)[:, :m, 0]
            ok = torch.isfinite(c).all(dim=1) & (c.abs().amax(dim=1) < 20.0)
            last = torch.zeros(m, **f64)
            last[-1] = 1.0
            c = torch.where(ok[:, None], c, last[None])
            Pm = (torch.stack(hP, dim=1) * c[:, :, None, None, None]).sum(dim=1)
            Rm = (rv * c[:, :, None]).sum(dim=1).reshape(Pa.shape)
            Pl = Pl.clone()
            Pl[idx] = Pm + bet[:, None, None, None] * Rm
        conv_mask = torch.zeros(nb, dtype=torch.bool, device=self.dev)
        conv_mask[idx] = done
        return Pl, conv_mask

    def run(self, maxit=150, conv=1e-10, errtol=1e-7, verbose=False, refill=True, refill_kT=(0.01, 0.005)):
        """Stage 1: SCF of every atom with its frozen per-l electron counts (self.occ_l: the reference's tabulated
        ground-state configurations unless given).  Those are PBE configurations; with the exchange-only SAP
        functional a few of them have no bound Aufbau solution (Gd - Tm: the 4f level of the frozen 4f^n 6s^1
        configurations rises above zero) and cannot converge.  Stage 2 (refill): such atoms get their per-l counts
        from a finite-temperature SCF with ONE chemical potential across the l-blocks (what the reference's
        `--occs auto` leaves to OpenOrbitalOptimizer's occupation optimisation), annealed over refill_kT (0.01 -> 0.005 Eh
        reaches the same counts as 0.02 -> 0.005 in two thirds of the iterations), and are then
        converged again at zero temperature with those (fractional) counts frozen -- the reference's
        fixed_per_l path (src/sadatom/scf.cpp:390-405).  self.refilled maps Z to the new counts."""
        torch = self.torch
        nb, nl, N = self.nb, self.nl, self.N
        self.total_iterations = 0
        self.refilled = {}
        self.last_emax = torch.full((nb,), float("inf"), dtype=torch.float64, device=self.dev)
        every = torch.ones(nb, dtype=torch.bool, device=self.dev)
        Vg = self.local_potential(self.guess_potential())
        Pl, _ = self._densities(self.T[None, None] + self.ll1[None, :, None, None] * self.Tl[None, None] + Vg[:, None])
        Pl, done = self._solve(Pl, every, maxit=maxit, conv=conv, errtol=errtol, verbose=verbose, patience=40 if refill else None)
        if refill and maxit >= 60 and not bool(done.all()):
            bad = ~done
            for kT in refill_kT:
                Pl, _ = self._solve(Pl, bad, kT=kT, auto=True, maxit=400, conv=1e-9, errtol=1e-6, verbose=verbose)
            counts = self._last_occ.sum(dim=2)                                # (bad atoms, nl)
            counts = torch.where((counts - counts.round()).abs() < 1e-6, counts.round(), counts)
            counts = counts * (self.Z[bad] / counts.sum(dim=1))[:, None]
            occ = self.occ.clone()
            cn = counts.cpu().numpy()
            for k, a in enumerate(torch.nonzero(bad).flatten().tolist()):
                for l in range(nl):
                    occ[a, l] = torch.as_tensor(shell_occupations(cn[k, l], l, N), dtype=torch.float64, device=self.dev)
                self.refilled[self.zs[a]] = [float(x) for x in cn[k]]
            self.occ = occ
            Pl, again = self._solve(Pl, bad, maxit=maxit, conv=conv, errtol=errtol, verbose=verbose)
            done = done | again
        self.Pl, self.converged, self.iterations = Pl, done, self.total_iterations
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
        from concurrent.futures import ThreadPoolExecutor
        os.makedirs(directory, exist_ok=True)
        self.sap_table(0)      # creates the shared (read-only) table set before the threads start

        def one(a):
            tab = self.sap_table(a)
            path = os.path.join(directory, "result_%s.dat" % self.symbols[self.zs[a]])
            with open(path, "w") as f:
                f.write("".join("".join(" %24.16e" % v for v in row) + "\n" for row in tab))
            return path

        # the screening integrals are host-side post-processing (like the reference's): one atom per host thread
        with ThreadPoolExecutor(max_workers=min(16, os.cpu_count() or 1)) as ex:
            return list(ex.map(one, range(self.nb)))

    def write_atomdb_dump(self, path):
        """The "Z r Zeff" rows that the reference's tools/gen_sap_table.py reads from atomdb_dump
        (src/general/atomdb_dump.cpp) to regenerate the tabulated SAP charge of src/general/sap.cpp: for every atom of the
        batch, in order of Z, the effective charge Z_eff(r) = Z - r (V_H + V_xc) on the common radial grid (origin +
        quadrature points, identical for all atoms: the tool requires it).  `gen_sap_table.py sap.cpp < path` then
        splices the table (%.14e, three per line) into the reference's source unchanged."""
        order = sorted(range(self.nb), key=lambda a: self.zs[a])
        with open(path, "w") as f:
            for a in order:
                tab = self.sap_table(a)
                f.write("".join("%d %.16e %.16e\n" % (self.zs[a], r, z) for r, z in zip(tab[:, 0], tab[:, 8])))
        return path
