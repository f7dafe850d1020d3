"""Device-resident restricted Hartree-Fock around the GPU Fock build -- the CALLER side of the hot path
(SURVEY.md section 8f-2): density build, Fock transform and the per-symmetry-block eigenproblems stay on the GPU, so
that no dense matrix crosses PCIe during the SCF.

Follows the reference's driver glue: per-block S^-1/2 (src/general/scf_driver_common.h:176-187,
libhelfem/src/utils.cpp:121-158), P = sum_blocks (S^-1/2 C) occ (S^-1/2 C)^T scattered by the block's index list
(:406-425), F_block = S^-1/2^T (H0 + J + K) S^-1/2 (:489-520), energies as src/diatomic/main.cpp:426-459
(Exx = Tr P_sigma K_sigma summed over spins).  The solver itself (the reference delegates it to the third-party
OpenOrbitalOptimizer) is a plain Roothaan iteration with Pulay DIIS and Aufbau occupations over the m-blocks; dense
linear algebra of the solver goes through torch (cuSOLVER / cuBLAS) -- it is not part of the hot path.
"""
import time

import numpy as np


class DeviceRHF:
    def __init__(self, basis, nocc, Enucr=0.0, lang=None, occ_by_m=None):
        """basis: helfem_b200 basis with compute_tei() done (atomic or diatomic); nocc: doubly occupied orbitals.
        lang: pure-m / atomic grid order for the eval_Fxc(x_func = -1) call of the HF build (integrates Nel).
        occ_by_m: {m: doubly occupied orbitals of that block} freezes the occupations (default: Aufbau over all blocks)."""
        import torch
        from . import DFTGrid
        self.torch, self.basis, self.nocc, self.Enucr = torch, basis, int(nocc), float(Enucr)
        self.occ_by_m = occ_by_m
        t = basis.tables
        self.dev = torch.device("cuda", basis._device)
        n = self.n = t.Nbf
        S, T, V = t.one_electron()
        # symmetry blocks: functions of equal m
        mv = np.asarray(t.mval)
        drop = t.kind == 1
        sizes = [t.Nrad - (1 if (drop and m != 0) else 0) for m in mv]
        owner = np.repeat(mv, sizes)
        self.ms = sorted(set(int(m) for m in mv))
        f64 = dict(dtype=torch.float64, device=self.dev)
        self.idx, self.X, self.H0b, self.Sb = [], [], [], []
        for m in self.ms:
            ix = np.nonzero(owner == m)[0]
            self.idx.append(torch.from_numpy(ix).to(self.dev))
            Sb = torch.tensor(S[np.ix_(ix, ix)], **f64)
            d = torch.diagonal(Sb).rsqrt()
            w, U = torch.linalg.eigh(Sb * d[:, None] * d[None, :])
            self.X.append((U * w.rsqrt()[None, :]) @ U.T * d[:, None])
            self.H0b.append(torch.tensor(T[np.ix_(ix, ix)] + V[np.ix_(ix, ix)], **f64))   # blocks first: T + V is 1.8 GB
            self.Sb.append(Sb)
        del S, T, V
        # dense device matrices of the Fock build (column-major == transposed row-major; all symmetric here)
        self.P = torch.zeros((n, n), **f64)
        self.J = torch.empty_like(self.P)
        self.K = torch.empty_like(self.P)
        if lang is None:
            lang = 4 * int(np.max(t.lval)) + 12
        self.grid = DFTGrid(basis, lang) if t.kind == 1 else DFTGrid(basis, lang, 4 * int(np.max(np.abs(mv))) + 12)

    def _blocks_of(self, M):
        return [M.index_select(0, ix).index_select(1, ix) for ix in self.idx]

    def _density(self, Fb):
        """Aufbau over all blocks (degenerate +-m levels enter together), P scattered into the dense matrix."""
        torch = self.torch
        eps, C = [], []
        for F, X in zip(Fb, self.X):
            w, c = torch.linalg.eigh(X.T @ F @ X)
            eps.append(w)
            C.append(X @ c)
        allw = torch.cat(eps)
        thr = torch.sort(allw).values[self.nocc - 1] + 1e-9
        self.occ_per_block = []
        Pb = []
        for m, w, c in zip(self.ms, eps, C):
            k = int((w <= thr).sum()) if self.occ_by_m is None else int(self.occ_by_m.get(m, 0))
            self.occ_per_block.append(k)
            Co = c[:, :k]
            Pb.append(2.0 * Co @ Co.T)
        return Pb

    def _scatter(self, Pb):
        self.P.zero_()
        for ix, Pm in zip(self.idx, Pb):
            self.P[ix[:, None], ix[None, :]] = Pm

    def run(self, maxit=120, conv=1e-10, errtol=1e-7, verbose=False, damp_above=0.3):
        """damp_above: while the largest commutator element exceeds it the density is mixed (30 % new) instead of
        DIIS-extrapolated (the core-Hamiltonian guess of N2 is far from the solution)."""
        torch = self.torch
        stream = torch.cuda.current_stream().cuda_stream
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        Pb = self._density(self.H0b)
        self._scatter(Pb)
        hist_F, hist_e = [], []
        Eold, t_build = 0.0, 0.0
        for it in range(maxit):
            torch.cuda.synchronize()
            tb = time.perf_counter()
            exc, nel = self.basis.fock_build_device(self.P.data_ptr(), self.J.data_ptr(), self.K.data_ptr(), 0.5, -1, 0, None,
                                                    1e-12, stream)
            t_build += time.perf_counter() - tb
            Jb, Kb = self._blocks_of(self.J), self._blocks_of(self.K)
            E1 = sum(float((P * H).sum()) for P, H in zip(Pb, self.H0b))
            Ecoul = 0.5 * sum(float((P * J).sum()) for P, J in zip(Pb, Jb))
            Exx = 2.0 * 0.5 * sum(float((0.5 * P * K).sum()) for P, K in zip(Pb, Kb))
            E = E1 + Ecoul + Exx
            Fb = [H + J + K for H, J, K in zip(self.H0b, Jb, Kb)]
            err = [F @ P @ S - S @ P @ F for F, P, S in zip(Fb, Pb, self.Sb)]
            emax = max(float(e.abs().max()) for e in err)
            if verbose:
                print("it %2d  E %.10f  err %.2e  Nel %.10f  occ %s" % (it, E + self.Enucr, emax, nel, self.occ_per_block))
            self.result = {"E": E + self.Enucr, "E_electronic": E, "E1": E1, "Coulomb": Ecoul, "Exx": Exx, "Nel": nel,
                           "iterations": it + 1}
            if abs(E - Eold) < conv and emax < errtol:
                break
            Eold = E
            if damp_above is not None and emax > damp_above:
                hist_F, hist_e = [], []
                Pb = [0.7 * Po + 0.3 * Pn for Po, Pn in zip(Pb, self._density(Fb))]
                self._scatter(Pb)
                continue
            hist_F.append(Fb)
            hist_e.append(torch.cat([e.reshape(-1) for e in err]))
            hist_F, hist_e = hist_F[-8:], hist_e[-8:]
            m = len(hist_F)
            ev = torch.stack(hist_e)
            B = torch.zeros((m + 1, m + 1), dtype=torch.float64, device=self.dev)
            B[:m, :m] = ev @ ev.T
            B[m, :m] = -1.0
            B[:m, m] = -1.0
            rhs = torch.zeros(m + 1, dtype=torch.float64, device=self.dev)
            rhs[m] = -1.0
            try:
                c = torch.linalg.solve(B, rhs)[:m]
                Fd = [sum(c[i] * hist_F[i][b] for i in range(m)) for b in range(len(Fb))]
            except Exception:
                Fd = Fb
            Pb = self._density(Fd)
            self._scatter(Pb)
        torch.cuda.synchronize()
        self.result["seconds"] = time.perf_counter() - t0
        self.result["fock_build_seconds"] = t_build
        return self.result
