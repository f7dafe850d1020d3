"""Minimal Roothaan + DIIS SCF used only to pin the oracle against the
reference's recorded converged energy components (tests/refs/ci.json).

The reference drives its SCF with the third-party OpenOrbitalOptimizer (not in
the tree); any solver converging to the same Aufbau stationary point gives the
same energies.  Energy bookkeeping follows src/atomic/main.cpp:369-470 and
src/general/scf_driver_common.h:459-474 (Exx = 1/2 Tr P_sigma K_sigma summed
over spins).  Test infrastructure only.
"""
import numpy as np


def sinvh(S):
    """S^(-1/2), libhelfem/src/utils.cpp:121-158."""
    d = 1.0 / np.sqrt(S.diagonal())
    Sn = S * d[:, None] * d[None, :]
    w, V = np.linalg.eigh(Sn)
    X = V @ np.diag(w ** -0.5) @ V.T
    return X * d[:, None]


def rhf(S, H0, coulomb, exchange, nocc_by_block, blocks, maxit=60, conv=1e-10, verbose=False):
    """Restricted closed-shell HF with fixed occupations per symmetry block.

    blocks: list of index arrays; nocc_by_block: doubly-occupied count each.
    Returns dict of energy components (Ekin etc. need T,V passed via H0 split
    by the caller) and the density P (total, Pa+Pb).
    """
    n = S.shape[0]
    X = [sinvh(S[np.ix_(b, b)]) for b in blocks]

    def density(F):
        P = np.zeros((n, n))
        for b, Xb, no in zip(blocks, X, nocc_by_block):
            if no == 0:
                continue
            Fb = Xb.T @ F[np.ix_(b, b)] @ Xb
            w, C = np.linalg.eigh(Fb)
            Co = Xb @ C[:, :no]
            P[np.ix_(b, b)] += 2.0 * Co @ Co.T
        return P

    P = density(H0)
    Fs, Es = [], []
    Eold = 0.0
    for it in range(maxit):
        J = coulomb(P)
        K = exchange(P / 2.0)          # reference passes P_sigma; returns -K
        F = H0 + J + K
        Ecoul = 0.5 * np.sum(P * J)
        Exx = 2.0 * 0.5 * np.sum((P / 2.0) * K)
        E1 = np.sum(P * H0)
        E = E1 + Ecoul + Exx
        err = F @ P @ S - S @ P @ F
        Fs.append(F); Es.append(err)
        Fs, Es = Fs[-8:], Es[-8:]
        if verbose:
            print(it, E, np.max(np.abs(err)))
        if abs(E - Eold) < conv and np.max(np.abs(err)) < 1e-7:
            break
        Eold = E
        m = len(Fs)
        B = -np.ones((m + 1, m + 1)); B[m, m] = 0.0
        for a in range(m):
            for b in range(m):
                B[a, b] = np.sum(Es[a] * Es[b])
        rhs = np.zeros(m + 1); rhs[m] = -1.0
        try:
            c = np.linalg.solve(B, rhs)[:m]
            Fd = sum(ci * Fi for ci, Fi in zip(c, Fs))
        except np.linalg.LinAlgError:
            Fd = F
        P = density(Fd)
    return {"E": E, "E1": E1, "Coulomb": Ecoul, "Exx": Exx, "P": P, "J": J, "K": K, "iterations": it + 1}
