"""GPU parity: the CUDA J/K path (through the C ABI) against the CPU oracle.

Tolerance: 1e-12 relative Frobenius (BASELINE.json north_star); energies 1e-10 Eh.
"""
import numpy as np
import pytest

from tests import cases

pytestmark = pytest.mark.gpu
TOL = 1e-12


def _check_jk(basis, ob, P, tol=TOL):
    J = basis.coulomb(P)
    K = basis.exchange(P)
    Jo = ob.coulomb(P)
    Ko = ob.exchange(P)
    ej, ek = cases.relerr(J, Jo), cases.relerr(K, Ko)
    assert ej < tol, "J relative error %.3e" % ej
    assert ek < tol, "K relative error %.3e" % ek
    return J, K


@pytest.mark.parametrize("lmax,mmax,nelem", [(0, 0, 5), (1, 1, 3), (2, 2, 2), (2, 1, 3)])
def test_atomic_oracle_caches(hb, lmax, mmax, nelem):
    """Identical caches (the oracle's) on both sides: pure kernel parity."""
    ob = cases.oracle_atomic(4, lmax, mmax, nelem)
    basis = hb.TablesBasis(cases.tables_from_oracle_atomic(hb, ob))
    n = ob.Nbf()
    # m-block-diagonal density (typical SCF) and a fully dense random one
    P1 = cases.random_density(n, 3, 11, cases.m_blocks(ob.mval, ob.Nrad(), False))
    _check_jk(basis, ob, P1)
    P2 = cases.random_density(n, 5, 12)
    _check_jk(basis, ob, P2)


@pytest.mark.parametrize("lmax,mmax,nelem", [(0, 0, 5), (2, 2, 2)])
def test_atomic_own_setup(hb, lmax, mmax, nelem):
    """Product's own compute_tei + kernels against the oracle end to end."""
    ob = cases.oracle_atomic(4, lmax, mmax, nelem)
    basis = hb.AtomicTwoDBasis(4, lmax, mmax, nelem).compute_tei()
    P = cases.random_density(ob.Nbf(), 4, 5, cases.m_blocks(ob.mval, ob.Nrad(), False))
    _check_jk(basis, ob, P)


@pytest.mark.parametrize("lmax_per_m,nelem", [((2,), 2), ((3, 2), 2), ((2, 2, 2), 2), ((4,), 3)])
def test_diatomic_oracle_caches(hb, lmax_per_m, nelem):
    ob = cases.oracle_diatomic(3, 1, 1.8, lmax_per_m, nelem)
    basis = hb.TablesBasis(cases.tables_from_oracle_diatomic(hb, ob))
    n = ob.Nbf()
    P1 = cases.random_density(n, 3, 21, cases.m_blocks(ob.mval, ob.Nrad(), True))
    _check_jk(basis, ob, P1)
    P2 = cases.random_density(n, 4, 22)
    _check_jk(basis, ob, P2)


def test_diatomic_own_setup(hb):
    ob = cases.oracle_diatomic(7, 7, 2.07, (3, 2), 2)
    basis = hb.DiatomicTwoDBasis(7, 7, 2.07, [3, 2], 2).compute_tei()
    P = cases.random_density(ob.Nbf(), 3, 31, cases.m_blocks(ob.mval, ob.Nrad(), True))
    _check_jk(basis, ob, P)


def test_diatomic_absm_symmetric(hb):
    """+-m mirrored build (src/diatomic/basis.cpp:1886-1891,2067-2086) == full build for a
    +-m symmetric density, and == the oracle's mirrored build."""
    ob = cases.oracle_diatomic(7, 7, 2.07, (3, 3, 2), 2)
    n = ob.Nbf()
    blocks = cases.m_blocks(ob.mval, ob.Nrad(), True)
    ms = sorted(set(int(m) for m in ob.mval))
    rng = np.random.default_rng(3)
    P = np.zeros((n, n))
    for mabs in range(max(ms) + 1):
        bp = blocks[ms.index(mabs)]
        Q, _ = np.linalg.qr(rng.standard_normal((len(bp), 2)))
        blk = 2.0 * Q @ Q.T
        P[np.ix_(bp, bp)] = blk
        if mabs > 0:
            bm = blocks[ms.index(-mabs)]
            P[np.ix_(bm, bm)] = blk
    basis = hb.TablesBasis(cases.tables_from_oracle_diatomic(hb, ob))
    Kfull = basis.exchange(P)
    basis.set_absm_symmetric(True)
    Ksym = basis.exchange(P)
    assert cases.relerr(Ksym, Kfull) < TOL
    ob.absm_symmetric = True
    try:
        Ko = ob.exchange(P)
    finally:
        ob.absm_symmetric = False
    assert cases.relerr(Ksym, Ko) < TOL


def test_nonsymmetric_density(hb):
    """The reference's J/K are defined for any P (TwoDBasis.cpp:879-999, basis.cpp:1818-2089).  A symmetric
    P takes the half-storage exchange path, a non-symmetric one the general path; both must match the oracle
    and K[P]^T == K[P^T]."""
    rng = np.random.default_rng(77)
    ob = cases.oracle_diatomic(3, 1, 1.8, (3, 2), 2)
    basis = hb.TablesBasis(cases.tables_from_oracle_diatomic(hb, ob))
    n = ob.Nbf()
    blocks = cases.m_blocks(ob.mval, ob.Nrad(), True)
    P = np.zeros((n, n))
    for b in blocks:
        P[np.ix_(b, b)] = rng.standard_normal((len(b), len(b)))
    Kg = basis.exchange(P)
    assert cases.relerr(Kg, ob.exchange(P)) < TOL
    assert cases.relerr(basis.exchange(P.T), Kg.T) < TOL
    assert cases.relerr(basis.coulomb(P), ob.coulomb(P)) < TOL
    Ps = P + P.T
    Ks = basis.exchange(Ps)
    assert cases.relerr(Ks, ob.exchange(Ps)) < TOL
    assert cases.relerr(Ks, Kg + Kg.T) < TOL
    oa = cases.oracle_atomic(4, 2, 1, 3)
    ba = hb.TablesBasis(cases.tables_from_oracle_atomic(hb, oa))
    Pa = rng.standard_normal((oa.Nbf(), oa.Nbf()))
    assert cases.relerr(ba.exchange(Pa), oa.exchange(Pa)) < TOL
    assert cases.relerr(ba.exchange(Pa + Pa.T), oa.exchange(Pa + Pa.T)) < TOL


def test_linearity_and_symmetry(hb):
    """Size-independent properties: J,K linear in P; symmetric P -> symmetric J,K."""
    basis = hb.DiatomicTwoDBasis(7, 7, 2.07, [5, 4, 3], 3).compute_tei()
    n = basis.Nbf()
    t = basis.tables
    blocks = cases.m_blocks(t.mval, t.Nrad, True)
    Pa = cases.random_density(n, 3, 1, blocks)
    Pb = cases.random_density(n, 2, 2, blocks)
    Ka, Kb, Kab = basis.exchange(Pa), basis.exchange(Pb), basis.exchange(Pa + 0.5 * Pb)
    assert cases.relerr(Kab, Ka + 0.5 * Kb) < TOL
    Ja, Jb, Jab = basis.coulomb(Pa), basis.coulomb(Pb), basis.coulomb(Pa + 0.5 * Pb)
    assert cases.relerr(Jab, Ja + 0.5 * Jb) < TOL
    assert cases.relerr(Ka, Ka.T) < TOL and cases.relerr(Ja, Ja.T) < TOL
    # zero density -> zero matrices (empty screening lists)
    Z = np.zeros((n, n))
    assert np.all(basis.exchange(Z) == 0.0) and np.all(basis.coulomb(Z) == 0.0)


def test_he_scf_energy_on_gpu(hb):
    """He RHF through the GPU J/K: total -2.8616799956 Eh (tests/refs/ci.json atomic-He-hf-r)."""
    from oracle import scf
    basis = hb.AtomicTwoDBasis(2, 0, 0, 5).compute_tei()
    S, T, V = basis.tables.one_electron()
    r = scf.rhf(S, T + V, basis.coulomb, basis.exchange, [1], [np.arange(basis.Nbf())])
    assert abs(r["E"] - (-2.8616799956)) < 1e-9
    assert abs(r["Coulomb"] - 2.0515380305) < 2e-6 and abs(r["Exx"] - (-1.0257690153)) < 2e-6


def test_h2_scf_energy_on_gpu(hb):
    """H2 RHF (diatomic-H2-hf-r): total -1.1336295702 Eh."""
    from oracle import scf
    basis = hb.DiatomicTwoDBasis(1, 1, 1.4, [4], 3).compute_tei()
    S, T, V = basis.tables.one_electron()
    r = scf.rhf(S, T + V, basis.coulomb, basis.exchange, [1], [np.arange(basis.Nbf())])
    assert abs(r["E"] + 1.0 / 1.4 - (-1.1336295702)) < 1e-9


def test_errors(hb):
    b = hb.AtomicTwoDBasis(2, 0, 0, 2)
    with pytest.raises(ValueError):      # reference: "Primitive teis have not been computed!"
        b.coulomb(np.zeros((1, 1)))
    b.compute_tei()
    with pytest.raises(ValueError):      # size mismatch
        b.exchange(np.zeros((3, 3)))


def test_exchange_shards_sum_to_full(hb):
    """Owner-computes sharding on one device: every shard writes only the contributions of the units
    (output sector pair, element pair) it owns; the per-shard partial K (and J) matrices sum to the unsharded
    result, and the reported output pattern is the pattern of the COMPLETE matrix on every shard."""
    import torch
    basis = hb.DiatomicTwoDBasis(7, 7, 2.07, [4, 3, 2], 2).compute_tei()
    n = basis.Nbf()
    t = basis.tables
    P = cases.random_density(n, 3, 7, cases.m_blocks(t.mval, t.Nrad, True))
    Kfull, Jfull = basis.exchange(P), basis.coulomb(P)
    dP = torch.from_numpy(np.ascontiguousarray(P.T)).cuda()
    for nsh in (2, 3, 8):
        tot, totj = torch.zeros_like(dP), torch.zeros_like(dP)
        patterns = set()
        for sh in range(nsh):
            dK, dJ = torch.empty_like(dP), torch.empty_like(dP)
            basis.coulomb_exchange_device(dP.data_ptr(), dJ.data_ptr(), dK.data_ptr(), 1.0, sh, nsh)
            tot += dK
            totj += dJ
            patterns.add((tuple(basis.exchange_output_pattern()[1]), tuple(basis.exchange_output_pattern(True)[1])))
        assert cases.relerr(tot.cpu().numpy().T, Kfull) < TOL
        assert cases.relerr(totj.cpu().numpy().T, Jfull) < TOL
        assert len(patterns) == 1, "output pattern differs between shards"


def test_shards_one_task_per_output_pair(hb):
    """An s-only atomic density with lmax >= 1 gives ONE task per output pair: shards then own disjoint sets of
    output pairs, yet the reported pattern (what a compact collective is built from) must be the same everywhere."""
    import torch
    oa = cases.oracle_atomic(4, 2, 2, 3)
    basis = hb.TablesBasis(cases.tables_from_oracle_atomic(hb, oa))
    n = oa.Nbf()
    N = oa.Nrad()
    P = np.zeros((n, n))
    P[:N, :N] = cases.random_density(N, 2, 5)       # (l, m) = (0, 0) block only
    Kfull = basis.exchange(P)
    assert cases.relerr(Kfull, oa.exchange(P)) < TOL
    dP = torch.from_numpy(np.ascontiguousarray(P.T)).cuda()
    for nsh in (2, 4):
        tot = torch.zeros_like(dP)
        patterns = set()
        for sh in range(nsh):
            dK = torch.empty_like(dP)
            basis.exchange_device(dP.data_ptr(), dK.data_ptr(), sh, nsh)
            tot += dK
            patterns.add(tuple(basis.exchange_output_pattern()[1]))
        assert cases.relerr(tot.cpu().numpy().T, Kfull) < TOL
        assert len(patterns) == 1


def test_shards_headline_kernels(hb):
    """Sharding by (output pair, element pair) on the NP = 16 kernels of the bench: the region mask of the fold,
    K-split partials per shard, 8 shards."""
    import torch
    from tests.test_gpu_parity_large import gu_density
    basis = hb.DiatomicTwoDBasis(7, 7, 2.07, [18, 17], 3).compute_tei()
    t = basis.tables
    for P in (gu_density(t, 3), cases.random_density(t.Nbf, 3, 4, cases.m_blocks(t.mval, t.Nrad, True))):
        Kfull = basis.exchange(P)
        dP = torch.from_numpy(np.ascontiguousarray(P.T)).cuda()
        for nsh in (2, 8):
            tot = torch.zeros_like(dP)
            for sh in range(nsh):
                dK = torch.empty_like(dP)
                basis.exchange_device(dP.data_ptr(), dK.data_ptr(), sh, nsh)
                tot += dK
            assert cases.relerr(tot.cpu().numpy().T, Kfull) < TOL


def test_sadatom_coulomb_exchange(hb):
    """Spherically averaged atom (gensap path, src/sadatom/basis.cpp:186-312)."""
    from oracle import sadatom as osad
    lmax = 2
    ob = cases.oracle_atomic(10, lmax, 0, 3)          # radial caches L = 0..2*lmax, m = 0 angular list
    so = osad.SadatomBasis(ob, lmax)
    sb = hb.SadatomTwoDBasis(10, lmax, 3).compute_tei()
    N = ob.Nrad()
    rng = np.random.default_rng(4)
    cube = []
    for l in range(lmax + 1):
        Q, _ = np.linalg.qr(rng.standard_normal((N, 2)))
        cube.append((2 * l + 1) * Q @ Q.T)
    cube[1] = np.zeros((N, N))                        # an empty shell is skipped (P[lin].norm()==0)
    Ko = so.exchange(cube)
    Kg = sb.exchange(cube)
    for l in range(lmax + 1):
        assert cases.relerr(Kg[l], Ko[l]) < TOL, l
    Prad = sum(cube) / (4 * np.pi)
    assert cases.relerr(sb.coulomb(Prad), so.coulomb(Prad)) < TOL
    with pytest.raises(ValueError):
        sb.exchange(cube[:2])


@pytest.mark.parametrize("kind", ["yukawa", "erfc"])
def test_sadatom_rs_exchange(hb, kind):
    """Range-separated exchange of the spherically averaged atom, src/sadatom/basis.cpp:154-184, :314-420."""
    from oracle import sadatom as osad
    lmax, par = 1, 0.37
    ob = cases.oracle_atomic(10, lmax, 0, 2)
    sb = hb.SadatomTwoDBasis(10, lmax, 2).compute_tei()
    if kind == "yukawa":
        ob.compute_yukawa(par)
        sb.compute_yukawa(par)
    else:
        ob.compute_erfc(par)
        sb.compute_erfc(par)
    so = osad.SadatomBasis(ob, lmax)
    N = ob.Nrad()
    rng = np.random.default_rng(14)
    cube = []
    for l in range(lmax + 1):
        Q, _ = np.linalg.qr(rng.standard_normal((N, 2)))
        cube.append((2 * l + 1) * Q @ Q.T)
    Ko, Kg = so.rs_exchange(cube), sb.rs_exchange(cube)
    K0 = sb.exchange(cube)
    for l in range(lmax + 1):
        assert cases.relerr(Kg[l], Ko[l]) < 1e-11, l
        assert np.linalg.norm(Kg[l]) < np.linalg.norm(K0[l])


def test_fused_coulomb_exchange(hb):
    """hfq_coulomb_exchange == hfq_coulomb + hfq_exchange(kscale*P); sharded partial J and K sum to the full matrices."""
    import torch
    basis = hb.DiatomicTwoDBasis(7, 7, 2.07, [4, 3, 2], 2).compute_tei()
    n = basis.Nbf()
    t = basis.tables
    P = cases.random_density(n, 3, 7, cases.m_blocks(t.mval, t.Nrad, True))
    J0, K0 = basis.coulomb(P), basis.exchange(0.5 * P)
    J1, K1 = basis.coulomb_exchange(P, 0.5)
    assert cases.relerr(J1, J0) < 1e-14 and cases.relerr(K1, K0) < 1e-14
    dP = torch.from_numpy(np.ascontiguousarray(P.T)).cuda()
    for nsh in (1, 3):
        tj, tk = torch.zeros_like(dP), torch.zeros_like(dP)
        for sh in range(nsh):
            dJ, dK = torch.empty_like(dP), torch.empty_like(dP)
            basis.coulomb_exchange_device(dP.data_ptr(), dJ.data_ptr(), dK.data_ptr(), 0.5, sh, nsh)
            tj += dJ
            tk += dK
        assert cases.relerr(tj.cpu().numpy().T, J0) < TOL and cases.relerr(tk.cpu().numpy().T, K0) < TOL
    # non-zero patterns reported for compact collectives cover everything that is non-zero
    for coul, M in ((False, K0), (True, J0)):
        bs, pairs = basis.exchange_output_pattern(coul)
        mask = np.zeros((n, n), bool)
        for (sj, sk) in pairs:
            mask[np.ix_(bs == sj, bs == sk)] = True
        assert np.all(M[~mask] == 0.0)


def test_atomic_rs_exchange_yukawa(hb):
    """Range-separated exchange with the Yukawa kernel (CAMY-type functionals):
    compute_yukawa + rs_exchange, src/atomic/TwoDBasis.cpp:737-758, :1001-1131."""
    lam = 0.34
    ob = cases.oracle_atomic(4, 1, 1, 2)
    ob.compute_yukawa(lam)
    basis = hb.AtomicTwoDBasis(4, 1, 1, 2).compute_tei().compute_yukawa(lam)
    n = ob.Nbf()
    P = cases.random_density(n, 3, 41, cases.m_blocks(ob.mval, ob.Nrad(), False))
    Ko = ob.rs_exchange(P)
    Kg = basis.rs_exchange(P)
    assert cases.relerr(Kg, Ko) < 1e-11       # Bessel functions come from two libraries (libstdc++ vs scipy)
    # the screened exchange is weaker than the bare one, and tends to it for small lambda
    K0 = basis.exchange(P)
    assert np.linalg.norm(Kg) < np.linalg.norm(K0)
    # host setup parity of the screened caches
    Ty = basis._rs.tables
    Nel = 2
    for L in range(Ty.nlm):
        for e in range(Nel):
            sm, bg, B, _ = Ty.block(L, e)
            assert np.abs(sm[0] - ob.disjoint_iL[L * Nel + e]).max() <= 1e-12 * np.abs(sm[0]).max()
            W, Wo = B @ B.T, ob.rs_chol[L * Nel + e] @ ob.rs_chol[L * Nel + e].T
            assert np.abs(W - Wo).max() <= 1e-10 * np.abs(Wo).max()


def test_atomic_rs_exchange_erfc(hb):
    """Range-separated exchange with the erfc kernel (CAM / omega-B97 type functionals): compute_erfc +
    rs_exchange, src/atomic/TwoDBasis.cpp:762-771, :1001-1131 (erfc branch), dense pair tensors for every
    element pair (CoulombExchangeFE.h:275-297, :396-422)."""
    mu = 0.3
    ob = cases.oracle_atomic(4, 2, 1, 3)
    ob.compute_erfc(mu)
    n = ob.Nbf()
    Nel = ob.radial.Nel()
    rng = np.random.default_rng(8)
    Psym = cases.random_density(n, 3, 43, cases.m_blocks(ob.mval, ob.Nrad(), False))
    Pdense = cases.random_density(n, 4, 44)
    Pgen = rng.standard_normal((n, n))                       # non-symmetric: general (full storage) path
    # (a) the oracle's rs_ktei handed over: isolates the GPU contraction
    pref = [4 * np.pi * mu / (2 * L + 1) for L in range(ob.N_L)]
    T = cases.tables_from_oracle_atomic(hb, ob, pref=pref).set_pair_tensors(ob.rs_ktei)
    basis = hb.TablesBasis(T)
    for P in (Psym, Pdense, Pgen):
        assert cases.relerr(basis.exchange(P), ob.rs_exchange(P)) < TOL
    with pytest.raises(ValueError):
        basis.coulomb(Psym)                                  # no range-separated Coulomb build in the reference
    # (b) the product's own compute_erfc (Phi, quadrature, pair tensors) end to end
    own = hb.AtomicTwoDBasis(4, 2, 1, 3).compute_tei().compute_erfc(mu)
    To = own._rs.tables
    for L in range(ob.N_L):
        for ie in range(Nel):
            for je in range(Nel):
                a, b = To.pair_tensor(L, ie, je), ob.rs_ktei[(L * Nel + ie) * Nel + je]
                assert np.abs(a - b).max() <= 1e-12 * np.abs(b).max()
    for P in (Psym, Pgen):
        assert cases.relerr(own.rs_exchange(P), ob.rs_exchange(P)) < 1e-11
    # attenuated exchange is weaker than the bare one
    assert np.linalg.norm(own.rs_exchange(Psym)) < np.linalg.norm(own.exchange(Psym))


def test_fused_host_speculative_upload(hb):
    """hfq_coulomb_exchange with page-locked buffers: the second call with the same block structure runs on
    the predicted sparse upload (verified against the full one on the device), a call whose structure grew
    falls back to the full upload; all three match the oracle (include/helfem_b200.h, out[19])."""
    import ctypes
    import torch
    ob = cases.oracle_diatomic(3, 1, 1.8, (3, 2), 2)
    basis = hb.TablesBasis(cases.tables_from_oracle_diatomic(hb, ob))
    n = ob.Nbf()
    blocks = cases.m_blocks(ob.mval, ob.Nrad(), True)
    hP = torch.zeros((n, n), dtype=torch.float64).pin_memory()
    hJ = torch.empty((n, n), dtype=torch.float64).pin_memory()
    hK = torch.empty((n, n), dtype=torch.float64).pin_memory()

    def run(P):
        hP.copy_(torch.from_numpy(np.ascontiguousarray(P.T)))   # column-major for the C ABI
        hJ.fill_(float("nan"))
        hK.fill_(float("nan"))
        hb._check(hb.lib().hfq_coulomb_exchange(basis._context(), hP.data_ptr(), n, 0.5, hJ.data_ptr(), n,
                                                hK.data_ptr(), n))
        return hJ.numpy().T.copy(), hK.numpy().T.copy(), int(basis.last_timings()["speculative_hits"])

    P1 = cases.random_density(n, 2, 5, blocks[:1])          # only the first m block
    P2 = cases.random_density(n, 3, 6, blocks[:1])          # same structure, other values
    P3 = cases.random_density(n, 3, 7, blocks)              # structure grows: prediction fails
    hits = []
    for P in (P1, P2, P3, P3):
        J, K, h = run(P)
        hits.append(h)
        assert cases.relerr(J, ob.coulomb(P)) < TOL
        assert cases.relerr(K, ob.exchange(0.5 * P)) < TOL
    assert hits == [0, 1, 1, 2]
