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


def rhf(S, H0, coulomb, exchange, nocc_by_block, blocks, maxit=60, conv=1e-10, verbose=False, damp_above=None):
    """Restricted closed-shell HF with fixed occupations per symmetry block.

    blocks: list of index arrays; nocc_by_block: doubly-occupied count each.
    damp_above: while the largest commutator element exceeds it, the density is mixed (30 % new) instead of
    extrapolated (a poor core-Hamiltonian guess, e.g. N2, sends plain DIIS astray).
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
        emax = np.max(np.abs(err))
        if verbose:
            print(it, E, emax)
        if abs(E - Eold) < conv and emax < 1e-7:
            break
        Eold = E
        if damp_above is not None and emax > damp_above:
            Fs, Es = [], []
            P = 0.7 * P + 0.3 * density(F)
            continue
        Fs.append(F); Es.append(err)
        Fs, Es = Fs[-8:], Es[-8:]
        P = density(_diis(Fs, Es, F))
    return {"E": E, "E1": E1, "Coulomb": Ecoul, "Exx": Exx, "P": P, "J": J, "K": K, "iterations": it + 1}


def _diis(Fs, Es, F):
    m = len(Fs)
    B = -np.ones((m + 1, m + 1)); B[m, m] = 0.0
    for a in range(m):
        for b in range(m):
            B[a, b] = np.sum(Es[a] * Es[b])
    rhs = np.zeros(m + 1); rhs[m] = -1.0
    try:
        c = np.linalg.solve(B, rhs)[:m]
        return sum(ci * Fi for ci, Fi in zip(c, Fs))
    except np.linalg.LinAlgError:
        return F


def rks(S, H0, coulomb, vxc, nocc_by_block, blocks, maxit=80, conv=1e-10, verbose=False):
    """Restricted closed-shell Kohn-Sham with fixed occupations per symmetry block (pure functionals).
    vxc(P) -> (Hxc, Exc, Nel): the reference's grid.eval_Fxc on the total density (src/atomic/main.cpp:385-393).
    Energy bookkeeping as src/atomic/main.cpp:410-446: E = Tr P (T + V) + 1/2 Tr P J + Exc."""
    n = S.shape[0]
    X = [sinvh(S[np.ix_(b, b)]) for b in blocks]

    def density(F):
        P = np.zeros((n, n))
        for b, Xb, no in zip(blocks, X, nocc_by_block):
            if no == 0:
                continue
            w, C = np.linalg.eigh(Xb.T @ F[np.ix_(b, b)] @ Xb)
            Co = Xb @ C[:, :no]
            P[np.ix_(b, b)] += 2.0 * Co @ Co.T
        return P

    P = density(H0)
    Fs, Es = [], []
    Eold = 0.0
    for it in range(maxit):
        J = coulomb(P)
        Hxc, Exc, Nel = vxc(P)
        F = H0 + J + Hxc
        Ecoul = 0.5 * np.sum(P * J)
        E1 = np.sum(P * H0)
        E = E1 + Ecoul + Exc
        err = F @ P @ S - S @ P @ F
        Fs.append(F); Es.append(err)
        Fs, Es = Fs[-8:], Es[-8:]
        if verbose:
            print(it, E, np.max(np.abs(err)), Nel)
        if abs(E - Eold) < conv and np.max(np.abs(err)) < 1e-7:
            break
        Eold = E
        P = density(_diis(Fs, Es, F))
    return {"E": E, "E1": E1, "Coulomb": Ecoul, "XC": Exc, "Nel": Nel, "P": P, "iterations": it + 1}


def atomic_vxc(grid, n, func_ids, thr=1e-12):
    """vxc(P) for rks() on an oracle grid with eval_density / eval_fxc (atomic 3D, diatomic pure-m / 3D)."""
    from . import xc
    gga = any(xc.is_gga(f) for f in func_ids)
    mgga = any(xc.is_mgga(f) for f in func_ids)

    def vxc(P):
        if mgga:   # density, gradient and tau; the functional returns vtau as well (src/atomic/dftgrid.cpp:304-359)
            d = grid.eval_density(P, None, True, True, False)
            exc, vrho, vsigma, vtau = xc.evaluate_mgga(func_ids, d["rho"][:, 0], d["sigma"][:, 0], d["tau"][:, 0], thr)
            try:
                Ha, _, Exc = grid.eval_fxc(n, exc, vrho[:, None], vsigma[:, None], vtau[:, None])      # AtomicDFTGrid
            except TypeError:
                Ha, _, Exc = grid.eval_fxc(exc, vrho[:, None], vsigma[:, None], vtau[:, None])
            return Ha, Exc, d["Nel"]
        d = grid.eval_density(P, grad=gga)
        exc, vrho, vsigma = xc.evaluate_sum(func_ids, d["rho"][:, 0], d["sigma"][:, 0] if gga else None, thr)
        args = (exc, vrho[:, None], vsigma[:, None] if gga else None)
        try:
            Ha, _, Exc = grid.eval_fxc(n, *args)      # AtomicDFTGrid takes the matrix size
        except TypeError:
            Ha, _, Exc = grid.eval_fxc(*args)
        return Ha, Exc, d["Nel"]

    return vxc


def sadatom_rks(sb, grid, S, T, Tl, Vnuc, occ_by_l, func_ids, exx=False, thr=1e-12, maxit=80, conv=1e-10, verbose=False):
    """Restricted (spin-averaged) SCF of the spherically averaged atom with FIXED occupations per l, following the
    Fock build of src/sadatom/scf.cpp:145-283 line by line: P_l = C occ C^T; XC from the per-l cube / 4 pi, scaled
    back by 1 / 4 pi; J = coulomb(Prad / 4 pi); K_l = exchange(P_l / (2 (2l+1))); F_l = T + Vnuc + J + l(l+1) Tl
    (+ XC_l) (+ K_l).  occ_by_l[l] = list of occupation numbers of the lowest orbitals of that l (electrons).
    sb: oracle SadatomBasis (coulomb / exchange), grid: oracle SadatomDFTGrid (or None for pure HF)."""
    from . import xc
    L = len(occ_by_l)
    N = S.shape[0]
    X = sinvh(S)
    angfac = 4.0 * np.pi
    gga = any(xc.is_gga(f) for f in func_ids)

    def fock_l(l, J, XC, K):
        F = T + Vnuc + J + l * (l + 1) * Tl
        if XC is not None:
            F = F + XC[l]
        if K is not None:
            F = F + K[l]
        return F

    def densities(Fl):
        Pl = []
        for l in range(L):
            occ = np.asarray(occ_by_l[l], dtype=float)
            if len(occ) == 0 or np.abs(occ).max() == 0.0:
                Pl.append(np.zeros((N, N)))
                continue
            w, C = np.linalg.eigh(X.T @ Fl[l] @ X)
            Co = X @ C[:, :len(occ)]
            Pl.append(Co @ np.diag(occ) @ Co.T)
        return Pl

    Pl = densities([fock_l(l, 0.0, None, None) for l in range(L)])
    hist = []
    Eold = 0.0
    for it in range(maxit):
        Prad = sum(Pl)
        Ekin = sum(np.sum(Pl[l] * T) + l * (l + 1) * np.sum(Pl[l] * Tl) for l in range(L))
        XC, Exc, Nel = None, 0.0, None
        if func_ids:
            d = grid.eval_density([P / angfac for P in Pl], grad=gga)
            exc, vrho, vsigma = xc.evaluate_sum(func_ids, d["rho"][:, 0], d["sigma"][:, 0] if gga else None, thr)
            Ha, _, Exc = grid.eval_fxc(exc, vrho, vsigma)
            XC = [H / angfac for H in Ha]
            Nel = d["Nel"]
        Enuc = np.sum(Prad * Vnuc)
        J = sb.coulomb(Prad / angfac)
        Ecoul = 0.5 * np.sum(Prad * J)
        K, Exx = None, 0.0
        if exx:
            K = sb.exchange([Pl[l] / (2.0 * (2 * l + 1)) for l in range(L)])
            Exx = sum(0.5 * np.sum(K[l] * Pl[l]) for l in range(L))
        E = Ekin + Enuc + Ecoul + Exc + Exx
        Fl = [fock_l(l, J, XC, K) for l in range(L)]
        err = np.concatenate([(Fl[l] @ Pl[l] @ S - S @ Pl[l] @ Fl[l]).ravel() for l in range(L)])
        hist.append((np.stack(Fl), err))
        hist = hist[-8:]
        if verbose:
            print(it, E, np.abs(err).max(), Nel)
        if abs(E - Eold) < conv and np.abs(err).max() < 1e-7:
            break
        Eold = E
        Fd = _diis([h[0] for h in hist], [h[1] for h in hist], np.stack(Fl))
        Pl = densities(list(Fd))
    return {"E": E, "Ekin": Ekin, "Enuc": Enuc, "Coulomb": Ecoul, "XC": Exc, "Exx": Exx, "Nel": Nel, "Pl": Pl,
            "iterations": it + 1}
