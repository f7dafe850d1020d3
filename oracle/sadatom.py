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

    def exchange(self, P):
        g = Gaunt()
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
                    Prad[L] = Prad.get(L, 0.0) + (4.0 * np.pi / (2 * L + 1) * tot[L]) * P[lin]
            for L, PL in sorted(Prad.items()):
                K[lout] -= self.b._assemble_K(L, PL)
        return K
