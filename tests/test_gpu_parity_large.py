"""GPU parity on the branches the headline configuration (N2, lmax=30) takes.

The small cases of test_gpu_parity.py only reach sectors of <= 8 functions (NT = 1).  Here the m-shells are
large enough for: (m, l-parity) sector splitting with NT = 1 and with NT = 2 (`k_fold_reg<2,2>`, the bench's own
instantiation, in-element GEMM with 208..256 columns, S-split partial accumulators, `sec_span` padding), the
unsplit fallbacks `k_fold<3,2,false>` / parity-ordered `k_fold<4,2,true>` (HFQ_SECTOR_SPLIT=0), and the
multi-batch path (R buffer smaller than the task list).

Oracle: the C restatement of src/diatomic/basis.cpp:1627-2089 (oracle/csrc/jk_oracle.c, itself pinned on the numpy
oracle in tests/test_oracle.py) on IDENTICAL caches (the product's tables exported through hfq_tables_get_block).
K is compared on a seeded sample of output blocks (the C oracle computes selected blocks), J in full.
Tolerance 1e-12 relative Frobenius (BASELINE.json north_star).
"""
import os

import numpy as np
import pytest

from tests import cases

pytestmark = pytest.mark.gpu
TOL = 1e-12


def gu_density(T, seed, occ=((0, 0, 3), (0, 1, 2), (1, 1, 1))):
    """g/u-symmetric closed-shell density with the N2 occupation pattern (the bench density, bench.n2_density):
    orbitals live inside one (m, l-parity) subspace; +-m blocks identical."""
    rng = np.random.default_rng(seed)
    n = T.Nbf
    sub, off = {}, 0
    for m, l in zip(T.mval, T.lval):
        k = T.Nrad - (1 if m != 0 else 0)
        sub.setdefault((int(m), int(l) & 1), []).extend(range(off, off + k))
        off += k
    P = np.zeros((n, n))
    for (m, par, k) in occ:
        if (m, par) not in sub:
            continue
        idx = np.array(sub[(m, par)])
        Q, _ = np.linalg.qr(rng.standard_normal((len(idx), k)))
        blk = 2.0 * Q @ Q.T
        P[np.ix_(idx, idx)] = blk
        if m:
            jdx = np.array(sub[(-m, par)])
            P[np.ix_(jdx, jdx)] = blk
    return P


def check_against_c_oracle(basis, T, P, nsample, seed, tol=TOL, check_j=True, candidates=None):
    """K on `nsample` output blocks (all blocks if fewer) + full J against the C oracle."""
    from oracle import cjk
    cjk.use_all_cores()
    C = cjk.DiatomicCaches.from_tables(T)
    pi = C.pure_idx()
    nd = C.Nang * C.Nrad
    K = basis.exchange(P)
    Kd = np.zeros((nd, nd))
    Kd[np.ix_(pi, pi)] = K
    pure = np.zeros(nd, dtype=bool)
    pure[pi] = True
    na, N = C.Nang, C.Nrad
    if candidates is None:
        candidates = [(j, k) for j in range(na) for k in range(na)]
    rng = np.random.default_rng(seed)
    if len(candidates) > nsample:
        sel = [candidates[i] for i in rng.choice(len(candidates), nsample, replace=False)]
    else:
        sel = list(candidates)
    blk = C.exchange_blocks(C.expand(P), [s[0] for s in sel], [s[1] for s in sel])
    num = den = 0.0
    for b, (j, k) in enumerate(sel):
        rows, cols = pure[j * N:(j + 1) * N], pure[k * N:(k + 1) * N]
        ref = blk[b][np.ix_(rows, cols)]
        got = Kd[j * N:(j + 1) * N, k * N:(k + 1) * N][np.ix_(rows, cols)]
        num += np.sum((got - ref) ** 2)
        den += np.sum(ref ** 2)
    assert den > 0.0, "sampled blocks are all zero: the sample does not test anything"
    ek = np.sqrt(num / den)
    assert ek < tol, "K relative error %.3e on %d sampled blocks" % (ek, len(sel))
    ej = None
    if check_j:
        ej = cases.relerr(basis.coulomb(P), C.coulomb(P))
        assert ej < tol, "J relative error %.3e" % ej
    return ek, ej, K


def nonzero_candidates(T, P):
    """Output blocks (j, k) that can be non-zero for an m-diagonal density: mj == mk."""
    mv = T.mval
    return [(j, k) for j in range(len(mv)) for k in range(len(mv)) if mv[j] == mv[k]]


def test_parity_split_nt1(hb):
    """m-shells of 9..16 functions: (m, l-parity) split sectors, NT = 1 kernels."""
    basis = hb.DiatomicTwoDBasis(7, 7, 2.07, [10, 9, 8], 2).compute_tei()
    T = basis.tables
    P = gu_density(T, 5)
    check_against_c_oracle(basis, T, P, 4000, 1, candidates=nonzero_candidates(T, P))
    Pd = cases.random_density(T.Nbf, 4, 6)          # every block non-zero
    check_against_c_oracle(basis, T, Pd, 400, 2)


def test_headline_instantiation_nt2(hb):
    """m-shells of 19 / 17 functions -> parity-split sectors of 9..10 functions, NP = 16: k_fold_reg<2,2>,
    k_tgemm_ws with the symmetric half storage (g/u density, what bench.py times) and the general M = Ni^2 path
    (non-symmetric density), k_offdiag_mma<2>, S-split partials."""
    basis = hb.DiatomicTwoDBasis(7, 7, 2.07, [18, 17], 2).compute_tei()
    T = basis.tables
    P = gu_density(T, 7)
    _, _, K = check_against_c_oracle(basis, T, P, 4000, 3, candidates=nonzero_candidates(T, P))
    assert cases.relerr(K, K.T) < TOL
    # fused entry point (what the bench calls) == separate calls
    J2, K2 = basis.coulomb_exchange(P, 0.5)
    assert cases.relerr(K2, 0.5 * K) < TOL and cases.relerr(J2, basis.coulomb(P)) < TOL
    # dense symmetric density: every sector pair active
    Pd = cases.random_density(T.Nbf, 5, 8)
    check_against_c_oracle(basis, T, Pd, 300, 4)
    # non-symmetric m-diagonal density: general path (all Ni^2 rows, every element pair)
    rng = np.random.default_rng(9)
    Pn = np.zeros((T.Nbf, T.Nbf))
    for b in cases.m_blocks(T.mval, T.Nrad, True):
        Pn[np.ix_(b, b)] = rng.standard_normal((len(b), len(b)))
    check_against_c_oracle(basis, T, Pn, 300, 5, candidates=nonzero_candidates(T, Pn))


def test_three_elements_lmax20(hb):
    """Three radial elements (the bench's nelem; cross-element pairs on both sides of the diagonal) at lmax 20/19/18."""
    basis = hb.DiatomicTwoDBasis(7, 7, 2.07, [20, 19, 18], 3).compute_tei()
    T = basis.tables
    P = gu_density(T, 11)
    check_against_c_oracle(basis, T, P, 400, 6, candidates=nonzero_candidates(T, P))


@pytest.mark.parametrize("lmax", [18, 24])
def test_unsplit_fallback_kernels(hb, lmax):
    """HFQ_SECTOR_SPLIT=0: sectors of 19 functions run k_fold<3,2,false>, of 25 functions the parity-ordered
    k_fold<4,2,true>; must agree with the oracle and with the split (default) engine."""
    tabs = hb.Tables.diatomic(7, 7, 2.07, [lmax, lmax - 1], 2)
    P = gu_density(tabs, 13)
    os.environ["HFQ_SECTOR_SPLIT"] = "0"
    try:
        unsplit = hb.TablesBasis(tabs)
        _, _, K0 = check_against_c_oracle(unsplit, tabs, P, 400, 7, candidates=nonzero_candidates(tabs, P))
        Pd = cases.random_density(tabs.Nbf, 4, 14)
        Kd0 = unsplit.exchange(Pd)
    finally:
        del os.environ["HFQ_SECTOR_SPLIT"]
    split = hb.TablesBasis(tabs)
    assert cases.relerr(split.exchange(P), K0) < TOL
    assert cases.relerr(split.exchange(Pd), Kd0) < TOL


def test_multi_batch(hb):
    """R buffer capped (HFQ_R_BUDGET_MB) so that the task list runs in several batches accumulating into Kacc."""
    tabs = hb.Tables.diatomic(7, 7, 2.07, [10, 9, 8], 2)
    P = cases.random_density(tabs.Nbf, 4, 21)
    full = hb.TablesBasis(tabs)
    Kref = full.exchange(P)
    os.environ["HFQ_R_BUDGET_MB"] = "48"
    try:
        capped = hb.TablesBasis(tabs)
        K = capped.exchange(P)
        Kg = capped.exchange(gu_density(tabs, 22))
    finally:
        del os.environ["HFQ_R_BUDGET_MB"]
    assert cases.relerr(K, Kref) < TOL
    assert cases.relerr(Kg, full.exchange(gu_density(tabs, 22))) < TOL
    check_against_c_oracle(full, tabs, P, 200, 8, check_j=False)


def test_nan_density_propagates(hb):
    """The reference skips a block only when norm < 10 eps (basis.cpp:1858); a NaN block is not skipped, so a
    diverged SCF is visible in K."""
    basis = hb.DiatomicTwoDBasis(3, 1, 1.8, [2, 1], 2).compute_tei()
    n = basis.Nbf()
    P = cases.random_density(n, 2, 3, cases.m_blocks(basis.tables.mval, basis.tables.Nrad, True))
    P[1, 1] = np.nan
    assert np.isnan(basis.exchange(P)).any()
    assert np.isnan(basis.coulomb(P)).any()


def test_n2_headline_size_sampled_blocks(hb):
    """The bench workload itself (N2, lmax = 30, |m| <= 6, 3 elements, Nbf = 14 832, the bench density): K on 256
    sampled output blocks and the complete J against the C oracle."""
    import bench
    from oracle import cjk
    basis = hb.DiatomicTwoDBasis(7, 7, 2.07, [30] * 7, 3).compute_tei()
    T = basis.tables
    P = bench.n2_density(T)
    J, K = basis.coulomb_exchange(P, 0.5)
    cjk.use_all_cores()
    C = cjk.DiatomicCaches.from_tables(T)
    cand = nonzero_candidates(T, P)
    rng = np.random.default_rng(30)
    sel = [cand[i] for i in rng.choice(len(cand), 256, replace=False)]
    blk = C.exchange_blocks(0.5 * C.expand(P), [s[0] for s in sel], [s[1] for s in sel])
    par = bench.parity_of_blocks(C, sel, blk, K, J, P)
    assert par["max_relerr_K"] < TOL and par["max_relerr_J"] < TOL, par
    # blocks off the m-diagonal are exactly zero (the reference never touches them for this density)
    mv = np.repeat(T.mval, [T.Nrad - (1 if m else 0) for m in T.mval])
    assert not K[mv[:, None] != mv[None, :]].any()


def test_n2_rhf_device_scf_matches_oracle_scf(hb):
    """N2 RHF at the shape of the reference's diatomic-N2-hf-r case (tests/cases.json: Rbond 2.07, nelem 3,
    lmax = 13,9, 7 doubly occupied orbitals; the case is weekly-tier and has no recorded value): the device-resident
    SCF (helfem_b200/scf.py: GPU J, K, grid Nel, solver on the device) against the oracle SCF (numpy setup, C J/K)
    to 1e-10 Eh (BASELINE.json north_star)."""
    from oracle import cjk, scf
    from helfem_b200.scf import DeviceRHF
    ob = cases.oracle_diatomic(7, 7, 2.07, (13, 9), 3)
    C = cjk.DiatomicCaches.from_oracle(ob)
    cjk.use_all_cores()
    S, T, V = ob.overlap(), ob.kinetic(), ob.nuclear()
    blocks = cases.m_blocks(ob.mval, ob.Nrad(), True)
    ms = sorted(set(int(m) for m in ob.mval))
    ro = scf.rhf(S, T + V, C.coulomb, C.exchange, [5 if m == 0 else 1 for m in ms], blocks, damp_above=0.3, maxit=150)
    basis = hb.DiatomicTwoDBasis(7, 7, 2.07, [13, 9], 3).compute_tei()
    rd = DeviceRHF(basis, 7, Enucr=49.0 / 2.07, occ_by_m={0: 5, 1: 1, -1: 1}).run()
    assert abs(rd["E_electronic"] - ro["E"]) < 1e-10 * abs(ro["E"]), (rd["E_electronic"], ro["E"])
    assert abs(rd["Coulomb"] - ro["Coulomb"]) < 1e-6 and abs(rd["Exx"] - ro["Exx"]) < 1e-6
    assert abs(rd["Nel"] - 14.0) < 1e-9
    # Aufbau over the blocks finds the same state (1 sigma_g .. 3 sigma_g + pi_u)
    ra = DeviceRHF(basis, 7, Enucr=49.0 / 2.07).run()
    assert abs(ra["E_electronic"] - ro["E"]) < 1e-9 * abs(ro["E"])
