"""Host side of the product on CPU: the C ABI loads and exports what include/helfem_b200.h
declares; the C++ setup (basis + compute_tei) matches the oracle; error behaviour."""
import ctypes
import os
import re

import numpy as np
import pytest

from tests import cases

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_abi_exports_every_declared_symbol(hb):
    header = open(os.path.join(ROOT, "include", "helfem_b200.h")).read()
    declared = set(re.findall(r"\b(hfq_[a-z_0-9]+)\s*\(", header))
    declared -= {"hfq_tables_desc", "hfq_tables_info"}
    assert len(declared) >= 19
    L = ctypes.CDLL(os.path.join(ROOT, "helfem_b200", "libhelfemqc_b200.so"))
    for name in sorted(declared):
        assert hasattr(L, name), name
    assert declared == set(hb.EXPORTED_SYMBOLS)


def test_atomic_setup_matches_oracle(hb):
    ob = cases.oracle_atomic(4, 1, 1, 3)
    T = hb.Tables.atomic(4, 1, 1, 3)
    assert (T.Nbf, T.Nrad, T.Nel, T.Nang) == (ob.Nbf(), ob.Nrad(), 3, ob.Nang())
    assert list(T.lval) == list(ob.lval) and list(T.mval) == list(ob.mval)
    Nel = 3
    for L in range(T.nlm):
        for e in range(Nel):
            sm, bg, B, sig = T.block(L, e)
            o_sm, o_bg, o_B = ob.disjoint_L[L * Nel + e], ob.disjoint_m1L[L * Nel + e], ob.prim_chol[L * Nel + e]
            assert np.abs(sm[0] - o_sm).max() <= 1e-13 * np.abs(o_sm).max()
            if e > 0:
                assert np.abs(bg[0] - o_bg).max() <= 1e-13 * np.abs(o_bg).max()
            assert B.shape == o_B.shape and np.all(sig == 1.0)
            assert np.abs(B @ B.T - o_B @ o_B.T).max() <= 1e-12 * np.abs(o_B @ o_B.T).max()
    S, Tk, V = T.one_electron()
    assert cases.relerr(S, ob.overlap()) < 1e-13
    assert cases.relerr(Tk, ob.kinetic()) < 1e-13
    assert cases.relerr(V, ob.nuclear()) < 1e-13


def test_diatomic_setup_matches_oracle(hb):
    ob = cases.oracle_diatomic(3, 1, 1.8, (3, 2), 2)
    T = hb.Tables.diatomic(3, 1, 1.8, [3, 2], 2)
    assert (T.Nbf, T.Nrad, T.Nang) == (ob.Nbf(), ob.Nrad(), ob.Nang())
    assert list(zip(T.lmL, T.lmM)) == ob.lm_map
    assert np.allclose(T.pref, ob.LMfac_abs(), rtol=1e-15)
    Nel = 2
    for ilm in range(T.nlm):
        for e in range(Nel):
            i = ilm * Nel + e
            sm, bg, B, sig = T.block(ilm, e)
            for c, (os_, ob_) in enumerate([(ob.disjoint_P0[i], ob.disjoint_Q0[i]), (ob.disjoint_P2[i], ob.disjoint_Q2[i])]):
                assert np.abs(sm[c] - os_).max() <= 1e-13 * np.abs(os_).max()
                if e > 0:
                    assert np.abs(bg[c] - ob_).max() <= 1e-13 * np.abs(ob_).max()
            W = (B * sig) @ B.T
            Wo = (ob.cd_B[i] * ob.cd_sigma[i]) @ ob.cd_B[i].T
            assert np.abs(W - Wo).max() <= 1e-11 * np.abs(Wo).max()
    S, Tk, V = T.one_electron()
    assert cases.relerr(S, ob.overlap()) < 1e-13
    assert cases.relerr(Tk, ob.kinetic()) < 1e-13
    assert cases.relerr(V, ob.nuclear()) < 1e-13


def test_from_arrays_round_trip(hb):
    ob = cases.oracle_diatomic(3, 1, 1.8, (2,), 2)
    T = cases.tables_from_oracle_diatomic(hb, ob)
    sm, bg, B, sig = T.block(3, 1)
    assert np.array_equal(sm[1], ob.disjoint_P2[3 * 2 + 1]) and np.array_equal(B, ob.cd_B[3 * 2 + 1])
    assert np.array_equal(sig, ob.cd_sigma[3 * 2 + 1])


def test_argument_errors(hb):
    with pytest.raises(ValueError):
        hb.Tables.atomic(2, 0, 1, 5)            # mmax > lmax
    with pytest.raises(ValueError):
        hb.Tables.diatomic(1, 1, 1.4, [0, 0], 2)  # lmax(|m|=1) < 1
    b = hb.AtomicTwoDBasis(2, 0, 0, 2)
    with pytest.raises(ValueError, match="Primitive teis"):
        b.coulomb(np.zeros((2, 2)))


def test_no_cpu_fallback(hb):
    """Without a CUDA device the Fock-build path must fail loudly, not fall back."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    b = hb.AtomicTwoDBasis(2, 0, 0, 2).compute_tei()
    with pytest.raises(hb.HfqError, match="CUDA"):
        b.coulomb(np.zeros((b.Nbf(), b.Nbf())))


def test_device_compute_tei_needs_its_gpu(hb):
    """hfq_tables_diatomic_device has no host fallback: without that CUDA device it fails."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(hb.HfqError):
        hb.Tables.diatomic(7, 7, 2.07, [2, 2], 1, device=0)


def test_sap_occupation_tables():
    """Host logic of the batched SAP driver: the tabulated per-l configurations (extracted from the reference's
    pbe_ground_states, src/diatomic/twodquadrature.cpp:26-145) hold Z electrons each, shells fill by capacity with a
    fractional remainder, and the by-element distribution deals every element to exactly one rank."""
    from helfem_b200 import sap
    occ, sym = sap.ground_state_occupations()
    assert len(occ) == 118 and sym[1] == "H" and sym[86] == "Rn" and occ[10] == [4, 6, 0, 0] and occ[36] == [8, 18, 10, 0]
    for z, ol in occ.items():
        assert sum(ol) == z, z
    o = sap.shell_occupations(7.5, 2, 6)            # d shells: capacity 10
    assert list(o) == [7.5, 0, 0, 0, 0, 0]
    o = sap.shell_occupations(24.25, 1, 6)          # p shells: capacity 6
    assert list(o) == [6, 6, 6, 6, 0.25, 0]
    zs = list(range(1, 87))
    parts = [sap.elements_of_rank(zs, r, 8) for r in range(8)]
    assert sorted(z for p in parts for z in p) == zs and max(len(p) for p in parts) - min(len(p) for p in parts) <= 1
    # heaviest-first round-robin: the six lanthanides whose tabulated configuration needs the refill land on six ranks
    assert len({r for r in range(8) for z in parts[r] if 64 <= z <= 69}) == 6


def test_product_does_not_touch_the_oracle():
    pkg = os.path.join(ROOT, "helfem_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cpp", ".cu", ".h", ".cuh")):
                txt = open(os.path.join(dp, f)).read()
                assert "oracle" not in txt.replace("the oracle", "").lower() or f == "__init__.py" and "oracle" not in txt, f


def test_erfc_host_setup_matches_oracle(hb):
    """Product's Phi and compute_erfc pair tensors (host C++) against the oracle."""
    from oracle import erfc
    rng = np.random.default_rng(3)
    L = hb.lib()
    for n in range(5):
        for scale in (0.05, 0.5, 2.0, 6.0):
            Xi, xi = rng.uniform(0, scale, 200), rng.uniform(0, scale, 200)
            got = np.array([L.hfq_erfc_phi(n, float(a), float(b)) for a, b in zip(Xi, xi)])
            ref = erfc.Phi(n, Xi, xi)
            assert np.all(np.abs(got - ref) <= 1e-13 * np.abs(ref) + 2e-12)
    ob = cases.oracle_atomic(4, 1, 1, 2)
    ob.compute_erfc(0.3)
    T = hb.Tables.atomic_erfc(4, 1, 1, 2, mu=0.3)
    Nel = 2
    for Lq in range(3):
        for ie in range(Nel):
            for je in range(Nel):
                a, b = T.pair_tensor(Lq, ie, je), ob.rs_ktei[(Lq * Nel + ie) * Nel + je]
                assert np.abs(a - b).max() <= 1e-12 * np.abs(b).max()


def test_pair_tensor_round_trip(hb):
    """hfq_tables_set_pair_tensors / hfq_tables_get_pair_tensor keep the reference's rs_ktei layout
    (ktei[kk*Ni + jj, ll*Ni + ii], column-major) and reject inconsistent sizes."""
    ob = cases.oracle_atomic(4, 1, 1, 2)
    ob.compute_erfc(0.25)
    pref = [4 * np.pi * 0.25 / (2 * L + 1) for L in range(ob.N_L)]
    T = cases.tables_from_oracle_atomic(hb, ob, pref=pref)
    assert T.pair_tensor(0, 0, 0) is None
    T.set_pair_tensors(ob.rs_ktei)
    Nel = ob.radial.Nel()
    for L in range(ob.N_L):
        for ie in range(Nel):
            for je in range(Nel):
                assert np.array_equal(T.pair_tensor(L, ie, je), ob.rs_ktei[(L * Nel + ie) * Nel + je])
    with pytest.raises(ValueError):
        T.set_pair_tensors(ob.rs_ktei[:-1])


def test_sap_table_matches_oracle_and_physics(hb):
    """hfq_sap_table (host-side, src/sadatom/main.cpp:55-107) against the oracle restatement, for a converged Be RHF
    density: the reference's own printed checks hold (electron count and Coulomb energy by quadrature), Z_eff runs
    from Z at the nucleus to Z - N + (xc tail -> 0) far out."""
    from oracle import sadatom as osad, scf
    Z = 4
    ob = cases.oracle_atomic(Z, 0, 0, 5)
    S, T, V = ob.overlap(), ob.kinetic(), ob.nuclear()
    r = scf.rhf(S, T + V, ob.coulomb, ob.exchange, [2], [np.arange(ob.Nbf())])
    P = r["P"]
    st = osad.SapTable(ob, Z)
    sb = hb.SadatomTwoDBasis(Z, 1, 5).compute_tei()
    N = ob.Nrad()
    Pp = 1e-3 * np.outer(np.linspace(0, 1, N) ** 2, np.linspace(0, 1, N) ** 2)       # a little l = 1 density
    for Pa, Pb in (([P, Pp], None), ([0.6 * P, 0.5 * Pp], [0.4 * P, 0.5 * Pp])):
        to, tg = st.table(Pa, Pb), sb.sap_table(Pa, Pb)
        assert to.shape == tg.shape == (5 * 75 + 1, 9)
        for c in range(9):
            # derivative columns (grad, lapl, tau) are sums of large terms of alternating sign near the nucleus
            tol = 1e-10 if c in (2, 3, 4) else 1e-12
            assert np.abs(to[:, c] - tg[:, c]).max() <= tol * max(np.abs(to[:, c]).max(), 1.0), c
    tab = sb.sap_table([P, np.zeros((N, N))])
    rr, rho, vc, w, zeff = tab[:, 0], tab[:, 1], tab[:, 5], tab[:, 7], tab[:, 8]
    assert abs(np.sum(w * rho * rr * rr) - 4.0) < 1e-10                       # main.cpp:101-102
    assert abs(0.5 * np.sum(rr * rho * w * vc) - r["Coulomb"]) < 1e-9         # main.cpp:103-104
    assert zeff[0] == Z and abs(zeff[-1]) < 1e-6 and abs(vc[-1] - 4.0) < 1e-10
    assert np.all(np.diff(vc[1:]) > -1e-9)                                    # r V_H(r) = charge inside r grows
    with pytest.raises(ValueError):
        sb.sap_table([P])                                                     # wrong number of l blocks is caught below


def test_builtin_functionals_match_oracle(tmp_path):
    """The functional expressions the device kernel evaluates (helfem_b200/csrc/xc_builtin.cuh: dual-number forms of
    libxc ids 1, 7, 101, 130, 202, 231), compiled for the host by tests/cpp/xc_check.cpp, against the oracle's
    symbolically differentiated restatement (oracle/xc.py, pinned on the reference's recorded LDA / PBE / TPSS energies):
    exc, vrho, vsigma, vtau over 12 decades of density, 13 of the gradient and tau from tau_W to 100 tau_W.  The PBE
    and TPSS correlations are differences of terms that cancel at large reduced gradients, so energies and vrho are
    measured against the uniform-gas exchange scale; vsigma and vtau (never near a cancellation of their own) relative."""
    import subprocess
    from oracle import xc
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "xc_check")
    subprocess.check_call(["g++", "-O2", "-std=c++17", os.path.join(root, "tests", "cpp", "xc_check.cpp"), "-o", exe])
    rng = np.random.default_rng(3)
    n = 10.0 ** rng.uniform(-9, 3, 500)
    s = (10.0 ** rng.uniform(-10, 3, 500)) * n ** (8.0 / 3.0)
    tw = s / (8.0 * n)
    tau = tw * (1.0 + 10.0 ** rng.uniform(-3, 2, 500))
    tau[:50] = tw[:50] * (1.0 + 1e-12)          # one-orbital regions: tau = tau_W (z = 1, the cap of the TPSS correlation)
    ex, vx, _ = xc.evaluate(xc.XC_LDA_X, n)
    for fid in (xc.XC_LDA_X, xc.XC_LDA_C_VWN, xc.XC_GGA_X_PBE, xc.XC_GGA_C_PBE, xc.XC_MGGA_X_TPSS, xc.XC_MGGA_C_TPSS):
        inp = "".join("%d %.17e %.17e %.17e\n" % (fid, a, b, c) for a, b, c in zip(n, s, tau))
        out = np.array(subprocess.run([exe], input=inp, capture_output=True, text=True, check=True).stdout.split(), dtype=float)
        out = out.reshape(-1, 4)
        e, v, vs, vt = xc.evaluate_mgga([fid], n, s, tau, thr=0.0)
        # correlation at the lowest densities: the oracle's exp(-ec / gamma) - 1 loses digits (the product uses expm1)
        tol = 1e-10 if fid in (xc.XC_GGA_C_PBE, xc.XC_MGGA_C_TPSS) else 1e-12
        assert np.max(np.abs(out[:, 0] - e) / np.abs(ex)) < tol, fid
        assert np.max(np.abs(out[:, 1] - v) / np.abs(vx)) < tol, fid
        for col, ref in ((2, vs), (3, vt)):
            if not np.any(ref):
                assert np.all(out[:, col] == 0.0), (fid, col)
            else:
                assert np.max(np.abs(out[:, col] - ref) / np.abs(ref)) < tol, (fid, col)
