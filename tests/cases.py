"""Shared small basis configurations: oracle objects and the matching product tables."""
import functools

import numpy as np

from oracle import atomic as oat
from oracle import diatomic as odi
from oracle import fem as ofem


@functools.lru_cache(maxsize=None)
def oracle_atomic(Z, lmax, mmax, nelem, nnodes=15, Rmax=40.0, igrid=4, zexp=2.0):
    bval = ofem.get_grid(Rmax, nelem, igrid, zexp)
    lv, mv = oat.angular_basis(lmax, mmax)
    b = oat.TwoDBasis(Z, nnodes, 5 * nnodes, bval, lv, mv)
    b.compute_tei()
    return b


@functools.lru_cache(maxsize=None)
def oracle_diatomic(Z1, Z2, Rbond, lmax_per_m, nelem, nnodes=15, Rmax=40.0, igrid=4, zexp=1.0):
    Rh = 0.5 * Rbond
    bval = ofem.get_grid(np.arccosh(Rmax / Rh), nelem, igrid, zexp)
    lv, mv = odi.lm_to_l_m(list(lmax_per_m))
    b = odi.TwoDBasis(Z1, Z2, Rh, nnodes, 5 * nnodes, bval, lv, mv)
    b.compute_tei()
    return b


def tables_from_oracle_atomic(hb, b, pref=None):
    """Hand the ORACLE's caches to the product through hfq_tables_from_arrays."""
    Nel = b.radial.Nel()
    blocks = []
    for L in range(b.N_L):
        for e in range(Nel):
            sm = b.disjoint_L[L * Nel + e]
            bg = b.disjoint_m1L[L * Nel + e]
            if bg is None:
                bg = np.zeros_like(sm)
            Bf = b.prim_chol[L * Nel + e]
            blocks.append(([sm], [bg], Bf, np.ones(Bf.shape[1])))
    efirst = [b.radial.get_idx(e)[0] for e in range(Nel)]
    en = [b.radial.Nprim(e) for e in range(Nel)]
    if pref is None:
        pref = [4 * np.pi / (2 * L + 1) for L in range(b.N_L)]
    return hb.Tables.from_arrays(0, b.Nrad(), efirst, en, b.lval, b.mval, list(range(b.N_L)), [-1] * b.N_L, pref, blocks)


def tables_from_oracle_diatomic(hb, b):
    Nel = b.radial.Nel()
    blocks = []
    for ilm in range(len(b.lm_map)):
        for e in range(Nel):
            i = ilm * Nel + e
            blocks.append(([b.disjoint_P0[i], b.disjoint_P2[i]], [b.disjoint_Q0[i], b.disjoint_Q2[i]],
                           b.cd_B[i], b.cd_sigma[i]))
    efirst = [b.radial.get_idx(e)[0] for e in range(Nel)]
    en = [b.radial.fem.nprim(e) for e in range(Nel)]
    return hb.Tables.from_arrays(1, b.Nrad(), efirst, en, b.lval, b.mval, [p[0] for p in b.lm_map],
                                 [p[1] for p in b.lm_map], b.LMfac_abs(), blocks, Rhalf=b.Rhalf)


def random_density(n, nocc, seed, blocks=None):
    """Symmetric P = C occ C^T from a seeded random orthonormal C; block diagonal over
    `blocks` (lists of indices) if given."""
    rng = np.random.default_rng(seed)
    P = np.zeros((n, n))
    if blocks is None:
        blocks = [np.arange(n)]
    for b in blocks:
        b = np.asarray(b)
        k = min(nocc, len(b))
        Q, _ = np.linalg.qr(rng.standard_normal((len(b), k)))
        P[np.ix_(b, b)] = 2.0 * Q @ Q.T
    return P


def m_blocks(mval, Nrad, drop_first):
    """Index lists of the dense basis grouped by m (diatomic boundary rule optional)."""
    out = {}
    off = 0
    for m in mval:
        n = Nrad - (1 if (drop_first and m != 0) else 0)
        out.setdefault(int(m), []).extend(range(off, off + n))
        off += n
    return [np.array(v) for _, v in sorted(out.items())]


def relerr(a, b):
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)
