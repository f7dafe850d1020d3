"""GPU parity of the atomic DFT grid path (density, gradient, tau, Laplacian, XC matrix
assembly) against the CPU oracle, through the C ABI.  Tolerance 1e-12 relative."""
import numpy as np
import pytest

from tests import cases

pytestmark = pytest.mark.gpu
TOL = 1e-12


def _setup(hb, Z, lmax, mmax, nelem):
    from oracle import dftgrid_atomic as dg
    ob = cases.oracle_atomic(Z, lmax, mmax, nelem)
    basis = hb.AtomicTwoDBasis(Z, lmax, mmax, nelem).compute_tei()
    lang = mang = 4 * lmax + 12          # src/atomic/main.cpp:329-331
    return ob, basis, dg.AtomicDFTGrid(ob, lang, mang), hb.AtomicDFTGrid(basis, lang, mang)


def _rel(a, b):
    return np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300)


@pytest.mark.parametrize("lmax,mmax,nelem", [(0, 0, 3), (1, 1, 2), (2, 2, 2)])
def test_density_restricted(hb, lmax, mmax, nelem):
    ob, basis, og, gg = _setup(hb, 4, lmax, mmax, nelem)
    P = cases.random_density(ob.Nbf(), 3, 5)
    o = og.eval_density(P, None, True, True, True)
    g = gg.density(P, None, 7)
    assert gg.N == og.npoints()
    for k in ("rho", "sigma", "tau", "lapl"):
        assert _rel(g[k], o[k]) < TOL, k
    assert _rel(g["w"], o["w"]) < 1e-14
    assert abs(g["Nel"] - o["Nel"]) < 1e-11 * abs(o["Nel"]) and abs(g["Ekin"] - o["Ekin"]) < 1e-11 * abs(o["Ekin"])
    # quadrature integrates the density exactly: Nel = Tr(P S), Ekin = Tr(P T)
    S, T, _ = basis.tables.one_electron()
    assert abs(g["Nel"] - np.sum(P * S)) < 1e-8 and abs(g["Ekin"] - np.sum(P * T)) < 1e-6 * abs(np.sum(P * T))


def test_density_unrestricted(hb):
    ob, basis, og, gg = _setup(hb, 4, 1, 1, 2)
    Pa = cases.random_density(ob.Nbf(), 3, 5)
    Pb = cases.random_density(ob.Nbf(), 2, 6)
    o = og.eval_density(Pa, Pb, True, True, True)
    g = gg.density(Pa, Pb, 7)
    for k in ("rho", "sigma", "tau", "lapl"):
        assert g[k].shape == o[k].shape and _rel(g[k], o[k]) < TOL, k


@pytest.mark.parametrize("kind", ["lda", "gga", "mgga_t", "mgga_tl"])
def test_fxc_restricted(hb, kind):
    """Functional-independent parity: synthetic exc/v arrays exercise every accumulator."""
    ob, basis, og, gg = _setup(hb, 4, 1, 1, 2)
    n = ob.Nbf()
    P = cases.random_density(n, 3, 5)
    og.eval_density(P, None, True, True, True)
    gg.density(P, None, 7)
    rng = np.random.default_rng(7)
    N = gg.N
    exc = rng.uniform(-1, 0, N)
    vrho = rng.uniform(-1, 0, (N, 1))
    vsigma = rng.uniform(0, 1e-2, (N, 1)) if kind != "lda" else None
    vtau = rng.uniform(0, 1e-2, (N, 1)) if kind.startswith("mgga") else None
    vlapl = rng.uniform(0, 1e-2, (N, 1)) if kind == "mgga_tl" else None
    Ho, _, Eo = og.eval_fxc(n, exc, vrho, vsigma, vtau, vlapl)
    Hg, _, Eg = gg.fxc(exc, vrho, vsigma, vtau, vlapl)
    assert cases.relerr(Hg, Ho) < TOL
    assert abs(Eg - Eo) < 1e-12 * abs(Eo)
    assert cases.relerr(Hg, Hg.T) < 1e-13


def test_fxc_unrestricted(hb):
    ob, basis, og, gg = _setup(hb, 4, 1, 1, 2)
    n = ob.Nbf()
    Pa = cases.random_density(n, 3, 5)
    Pb = cases.random_density(n, 2, 6)
    og.eval_density(Pa, Pb, True, True, True)
    gg.density(Pa, Pb, 7)
    rng = np.random.default_rng(8)
    N = gg.N
    exc = rng.uniform(-1, 0, N)
    vrho = rng.uniform(-1, 0, (N, 2)); vsigma = rng.uniform(0, 1e-2, (N, 3))
    vtau = rng.uniform(0, 1e-2, (N, 2)); vlapl = rng.uniform(0, 1e-2, (N, 2))
    Hao, Hbo, Eo = og.eval_fxc(n, exc, vrho, vsigma, vtau, vlapl, polarized=True)
    Hag, Hbg, Eg = gg.fxc(exc, vrho, vsigma, vtau, vlapl)
    assert cases.relerr(Hag, Hao) < TOL and cases.relerr(Hbg, Hbo) < TOL
    assert abs(Eg - Eo) < 1e-12 * abs(Eo)


def test_eval_fxc_slater_exchange(hb):
    """Built-in XC_LDA_X through the reference-shaped eval_Fxc call; He LDA-exchange-only energy
    functional value checked against the oracle evaluating the same closed formula."""
    ob, basis, og, gg = _setup(hb, 2, 0, 0, 3)
    n = ob.Nbf()
    P = cases.random_density(n, 1, 9)
    o = og.eval_density(P)
    rho = o["rho"][:, 0]
    cx = -0.75 * (3.0 / np.pi) ** (1.0 / 3.0)
    exc = np.where(rho >= 1e-12, cx * np.cbrt(np.maximum(rho, 0)), 0.0)
    vr = np.where(rho >= 1e-12, 4.0 / 3.0 * cx * np.cbrt(np.maximum(rho, 0)), 0.0)[:, None]
    Ho, _, Eo = og.eval_fxc(n, exc, vr)
    H, Exc, Nel, Ekin = gg.eval_Fxc(1, 0, P)
    assert cases.relerr(H, Ho) < TOL and abs(Exc - Eo) < 1e-12 * abs(Eo) and abs(Nel - o["Nel"]) < 1e-11
    # HF drivers call eval_Fxc with x_func = -1 only to integrate Nel (src/atomic/main.cpp:240)
    H0, Exc0, Nel0, _ = gg.eval_Fxc(-1, 0, P)
    assert np.all(H0 == 0.0) and Exc0 == 0.0 and abs(Nel0 - o["Nel"]) < 1e-11
    with pytest.raises(ValueError):
        gg.eval_Fxc(263, 267, P)     # SCAN ids are not built in: use density + libxc + fxc
    with pytest.raises(ValueError):
        gg.eval_Fxc(7, 0, P)         # a correlation id in the exchange slot


@pytest.mark.parametrize("x_func,c_func", [(1, 7), (101, 130), (101, 0), (0, 130)])
def test_eval_fxc_builtin_functionals(hb, x_func, c_func):
    """Functionals evaluated on the device (csrc/xc_builtin.cuh: libxc ids 1, 7, 101, 130) through eval_Fxc against
    the oracle grid fed with the oracle's symbolically differentiated functionals (oracle/xc.py, pinned on the
    reference's recorded LDA / PBE energies)."""
    from oracle import xc
    ob, basis, og, gg = _setup(hb, 4, 1, 1, 2)
    n = ob.Nbf()
    P = cases.random_density(n, 2, 11)
    fids = [f for f in (x_func, c_func) if f > 0]
    gga = any(xc.is_gga(f) for f in fids)
    o = og.eval_density(P, None, gga)
    exc, vrho, vsigma = xc.evaluate_sum(fids, o["rho"][:, 0], o["sigma"][:, 0] if gga else None, 1e-12)
    Ho, _, Eo = og.eval_fxc(n, exc, vrho[:, None], vsigma[:, None] if gga else None)
    H, Exc, Nel, Ekin = gg.eval_Fxc(x_func, c_func, P)
    assert cases.relerr(H, Ho) < 1e-11 and abs(Exc - Eo) < 1e-12 * abs(Eo) and abs(Nel - o["Nel"]) < 1e-11 and Ekin == 0.0


def test_eval_fxc_builtin_tpss(hb):
    """TPSS meta-GGA (libxc ids 202 + 231) evaluated on the device: density, gradient and tau from the grid engine, the
    functional point-wise, v_rho / v_sigma / v_tau assembled -- against the oracle grid fed with the oracle's TPSS."""
    from oracle import xc
    ob, basis, og, gg = _setup(hb, 4, 1, 1, 2)
    n = ob.Nbf()
    P = cases.random_density(n, 2, 11)
    o = og.eval_density(P, None, True, True, False)
    exc, vrho, vsigma, vtau = xc.evaluate_mgga([xc.XC_MGGA_X_TPSS, xc.XC_MGGA_C_TPSS], o["rho"][:, 0], o["sigma"][:, 0],
                                               o["tau"][:, 0], 1e-12)
    Ho, _, Eo = og.eval_fxc(n, exc, vrho[:, None], vsigma[:, None], vtau[:, None])
    H, Exc, Nel, Ekin = gg.eval_Fxc(202, 231, P)
    assert cases.relerr(H, Ho) < 1e-10 and abs(Exc - Eo) < 1e-11 * abs(Eo) and abs(Nel - o["Nel"]) < 1e-11
    assert abs(Ekin - o["Ekin"]) < 1e-10 * abs(o["Ekin"])      # tau is integrated for a meta-GGA


def test_eval_fxc_builtin_polarised_exchange(hb):
    """Polarised PBE exchange through the spin-scaling relation E_x[na, nb] = (E_x[2 na] + E_x[2 nb]) / 2: the oracle
    evaluates the unpolarised functional at (2 n_s, 4 sigma_ss) and assembles with libxc's polarised layout."""
    from oracle import xc
    ob, basis, og, gg = _setup(hb, 4, 1, 1, 2)
    n = ob.Nbf()
    Pa, Pb = cases.random_density(n, 3, 5), cases.random_density(n, 2, 6)
    o = og.eval_density(Pa, Pb, True)
    N = og.npoints()
    vrho, vsigma, edens = np.zeros((N, 2)), np.zeros((N, 3)), np.zeros(N)
    for s_, col in ((0, 0), (1, 2)):
        e, v, vs = xc.evaluate(xc.XC_GGA_X_PBE, 2.0 * o["rho"][:, s_], 4.0 * o["sigma"][:, col], 0.0)
        edens += o["rho"][:, s_] * e
        vrho[:, s_], vsigma[:, col] = v, 2.0 * vs
    tot = o["rho"].sum(axis=1)
    ok = tot >= 1e-12
    exc = np.where(ok, edens / np.where(ok, tot, 1.0), 0.0)
    vrho[~ok], vsigma[~ok] = 0.0, 0.0
    Hao, Hbo, Eo = og.eval_fxc(n, exc, vrho, vsigma, polarized=True)
    (Ha, Hb), Exc, Nel, _ = gg.eval_Fxc(101, 0, Pa, Pb)
    assert cases.relerr(Ha, Hao) < 1e-11 and cases.relerr(Hb, Hbo) < 1e-11 and abs(Exc - Eo) < 1e-12 * abs(Eo)
    with pytest.raises(ValueError):
        gg.eval_Fxc(101, 130, Pa, Pb)     # polarised correlation is not built in


@pytest.mark.parametrize("method,Eref,XCref", [("lda", -2.8348356241, -0.9733148392), ("pbe", -2.8929348668, -1.0461619634),
                                               ("tpss", -2.9096638609, -1.0712420321)])
def test_he_ks_energy_on_gpu(hb, method, Eref, XCref):
    """atomic-He-lda-r / -gga-r / -mgga-r of the reference's tests/refs/ci.json: a Kohn-Sham SCF whose Fock build runs
    on the GPU (J = hfq_coulomb, XC = hfq_eval_fxc with the functional evaluated on the device) lands on the recorded
    total energy to 1e-9 Eh."""
    from oracle import scf
    ob, basis, og, gg = _setup(hb, 2, 0, 0, 5)
    S, T, V = basis.tables.one_electron()
    n = ob.Nbf()
    xf, cf = {"lda": (1, 7), "pbe": (101, 130), "tpss": (202, 231)}[method]

    def vxc(P):
        H, Exc, Nel, _ = gg.eval_Fxc(xf, cf, P)
        return H, Exc, Nel

    r = scf.rks(S, T + V, basis.coulomb, vxc, [1], [np.arange(n)])
    assert abs(r["E"] - Eref) < 1e-9 and abs(r["XC"] - XCref) < 2e-6 and abs(r["Nel"] - 2.0) < 1e-10


# ---------------------------------------------------------------------------------------------
# diatomic pure-m grid (src/diatomic/dftgrid_purem.cpp)
# ---------------------------------------------------------------------------------------------
def _setup_purem(hb, lmax_per_m, nelem, Z1=7, Z2=7, R=2.07):
    from oracle import dftgrid_purem as dp
    ob = cases.oracle_diatomic(Z1, Z2, R, tuple(lmax_per_m), nelem)
    basis = hb.DiatomicTwoDBasis(Z1, Z2, R, list(lmax_per_m), nelem).compute_tei()
    lang = 4 * max(lmax_per_m) + 12        # src/diatomic/main.cpp: ldft = 4*lmax+12
    return ob, basis, dp.PureMDFTGrid(ob, lang), hb.PureMDFTGrid(basis, lang)


def test_purem_density(hb):
    ob, basis, og, gg = _setup_purem(hb, (3, 2), 2)
    n = ob.Nbf()
    blocks = cases.m_blocks(ob.mval, ob.Nrad(), True)
    P = cases.random_density(n, 3, 5, blocks)
    o = og.eval_density(P, None, True, True, True)
    g = gg.density(P, None, 7)
    assert gg.N == og.npoints()
    for k in ("rho", "sigma", "tau", "lapl"):
        assert _rel(g[k], o[k]) < TOL, k
    assert _rel(g["w"], o["w"]) < 1e-14
    S, T, _ = basis.tables.one_electron()
    assert abs(g["Nel"] - np.sum(P * S)) < 1e-8 * abs(np.sum(P * S))
    assert abs(g["Ekin"] - np.sum(P * T)) < 1e-6 * abs(np.sum(P * T))
    # unrestricted
    Pb = cases.random_density(n, 2, 6, blocks)
    o = og.eval_density(P, Pb, True, True, True)
    g = gg.density(P, Pb, 7)
    for k in ("rho", "sigma", "tau", "lapl"):
        assert g[k].shape == o[k].shape and _rel(g[k], o[k]) < TOL, k


@pytest.mark.parametrize("kind", ["lda", "gga", "mgga_tl"])
def test_purem_fxc(hb, kind):
    ob, basis, og, gg = _setup_purem(hb, (3, 2), 2)
    n = ob.Nbf()
    blocks = cases.m_blocks(ob.mval, ob.Nrad(), True)
    P = cases.random_density(n, 3, 5, blocks)
    og.eval_density(P, None, True, True, True)
    gg.density(P, None, 7)
    rng = np.random.default_rng(17)
    N = gg.N
    exc = rng.uniform(-1, 0, N); vrho = rng.uniform(-1, 0, (N, 1))
    vsigma = rng.uniform(0, 1e-2, (N, 1)) if kind != "lda" else None
    vtau = rng.uniform(0, 1e-2, (N, 1)) if kind == "mgga_tl" else None
    vlapl = rng.uniform(0, 1e-2, (N, 1)) if kind == "mgga_tl" else None
    Ho, _, Eo = og.eval_fxc(exc, vrho, vsigma, vtau, vlapl)
    Hg, _, Eg = gg.fxc(exc, vrho, vsigma, vtau, vlapl)
    assert cases.relerr(Hg, Ho) < TOL and abs(Eg - Eo) < 1e-12 * abs(Eo)
    # the pure-m matrix is m-block diagonal
    for i, bi in enumerate(blocks):
        for j, bj in enumerate(blocks):
            if i != j:
                assert np.all(Hg[np.ix_(bi, bj)] == 0.0)


def test_purem_fxc_unrestricted(hb):
    ob, basis, og, gg = _setup_purem(hb, (2, 2), 2)
    n = ob.Nbf()
    blocks = cases.m_blocks(ob.mval, ob.Nrad(), True)
    Pa = cases.random_density(n, 3, 5, blocks); Pb = cases.random_density(n, 2, 6, blocks)
    og.eval_density(Pa, Pb, True, True, True)
    gg.density(Pa, Pb, 7)
    rng = np.random.default_rng(18)
    N = gg.N
    exc = rng.uniform(-1, 0, N)
    vrho = rng.uniform(-1, 0, (N, 2)); vsigma = rng.uniform(0, 1e-2, (N, 3))
    vtau = rng.uniform(0, 1e-2, (N, 2)); vlapl = rng.uniform(0, 1e-2, (N, 2))
    Hao, Hbo, Eo = og.eval_fxc(exc, vrho, vsigma, vtau, vlapl, polarized=True)
    Hag, Hbg, Eg = gg.fxc(exc, vrho, vsigma, vtau, vlapl)
    assert cases.relerr(Hag, Hao) < TOL and cases.relerr(Hbg, Hbo) < TOL and abs(Eg - Eo) < 1e-12 * abs(Eo)


def test_diatomic_3d_grid(hb):
    """General 3D diatomic grid (--symmetry=0): m-mixing density, GGA + tau, restricted and unrestricted."""
    from oracle import dftgrid_atomic as dg
    ob = cases.oracle_diatomic(3, 1, 1.8, (2, 2), 2)
    basis = hb.DiatomicTwoDBasis(3, 1, 1.8, [2, 2], 2).compute_tei()
    lang = mang = 4 * 2 + 12
    og, gg = dg.Diatomic3DGrid(ob, lang, mang), hb.DFTGrid(basis, lang, mang)
    n = ob.Nbf()
    Pa = cases.random_density(n, 3, 5); Pb = cases.random_density(n, 2, 6)     # dense: couples different m
    o = og.eval_density(Pa, Pb, True, True)
    g = gg.density(Pa, Pb, 3)
    for k in ("rho", "sigma", "tau"):
        assert _rel(g[k], o[k]) < TOL, k
    S, T, _ = basis.tables.one_electron()
    assert abs(g["Nel"] - np.sum((Pa + Pb) * S)) < 1e-8 * abs(np.sum((Pa + Pb) * S))
    rng = np.random.default_rng(27)
    N = gg.N
    exc = rng.uniform(-1, 0, N)
    vrho = rng.uniform(-1, 0, (N, 2)); vsigma = rng.uniform(0, 1e-2, (N, 3)); vtau = rng.uniform(0, 1e-2, (N, 2))
    Hao, Hbo, Eo = og.eval_fxc(exc, vrho, vsigma, vtau, polarized=True)
    Hag, Hbg, Eg = gg.fxc(exc, vrho, vsigma, vtau)
    assert cases.relerr(Hag, Hao) < TOL and cases.relerr(Hbg, Hbo) < TOL and abs(Eg - Eo) < 1e-12 * abs(Eo)
    # purem on == off for an m-diagonal density (invariant of tests/cases.json)
    Pm = cases.random_density(n, 3, 9, cases.m_blocks(ob.mval, ob.Nrad(), True))
    d3 = gg.density(Pm, None, 3)
    gp = hb.DFTGrid(basis, lang)          # pure-m grid on the same context replaces the 3D one
    dp_ = gp.density(Pm, None, 3)
    assert abs(d3["Nel"] - dp_["Nel"]) < 1e-9 * abs(dp_["Nel"]) and abs(d3["Ekin"] - dp_["Ekin"]) < 1e-9 * abs(dp_["Ekin"])


def _sad_setup(hb, lmax=2, nelem=3):
    from oracle import sadatom as osad
    ob = cases.oracle_atomic(10, lmax, 0, nelem)          # radial caches only: angular list irrelevant for the grid
    basis = hb.SadatomTwoDBasis(10, lmax, nelem).compute_tei()
    return ob, basis, osad.SadatomDFTGrid(ob, lmax), hb.SadatomDFTGrid(basis)


def _sad_cube(N, lmax, seed, nocc=2):
    rng = np.random.default_rng(seed)
    cube = []
    for l in range(lmax + 1):
        Q, _ = np.linalg.qr(rng.standard_normal((N, nocc)))
        cube.append((2 * l + 1) * Q @ Q.T)
    return cube


@pytest.mark.parametrize("pol", [False, True])
def test_sadatom_grid_density(hb, pol):
    """Radial-only grid of the spherically averaged atom, src/sadatom/dftgrid.cpp:45-241, :464-486."""
    ob, basis, og, gg = _sad_setup(hb)
    N = ob.Nrad()
    Pa, Pb = _sad_cube(N, 2, 1), (_sad_cube(N, 2, 2, 1) if pol else None)
    o = og.eval_density(Pa, Pb, True, True, True)
    g = gg.density(Pa, Pb, 7)
    assert gg.N == og.npoints()
    for k in ("rho", "sigma", "tau"):
        assert g[k].shape == o[k].shape and _rel(g[k], o[k]) < TOL, k
    # the Laplacian sum over basis-function pairs cancels by ~5 digits next to the nucleus for these
    # unphysical random densities (f_u f_v'' of alternating sign): parity is measured against the sum of the
    # absolute values of the summands, the quantity any summation order is accurate to
    assert g["lapl"].shape == o["lapl"].shape
    assert np.all(np.abs(g["lapl"] - o["lapl"]) <= TOL * o["lapl_scale"] + 1e-300)
    assert _rel(g["w"], o["w"]) < 1e-14
    assert abs(g["Nel"] - o["Nel"]) < 1e-11 * abs(o["Nel"])


@pytest.mark.parametrize("kind,pol", [("lda", False), ("gga", False), ("mgga_t", False), ("mgga_tl", False),
                                      ("gga", True), ("mgga_tl", True)])
def test_sadatom_grid_fxc(hb, kind, pol):
    """Fock-cube assembly incl. the l(l+1) tau term and the Laplacian cross terms, src/sadatom/dftgrid.cpp:256-460."""
    ob, basis, og, gg = _sad_setup(hb)
    N = ob.Nrad()
    Pa, Pb = _sad_cube(N, 2, 3), (_sad_cube(N, 2, 4, 1) if pol else None)
    og.eval_density(Pa, Pb, True, True, True)
    gg.density(Pa, Pb, 7)
    rng = np.random.default_rng(11)
    Np, ns = gg.N, (2 if pol else 1)
    exc = rng.uniform(-1, 0, Np)
    vrho = rng.uniform(-1, 0, (Np, ns))
    vs = rng.uniform(0, 1e-2, (Np, 3 if pol else 1)) if kind != "lda" else None
    vt = rng.uniform(0, 1e-2, (Np, ns)) if kind.startswith("mgga") else None
    vl = rng.uniform(0, 1e-2, (Np, ns)) if kind == "mgga_tl" else None
    Hao, Hbo, Eo = og.eval_fxc(exc, vrho, vs, vt, vl)
    Hag, Hbg, Eg = gg.fxc(exc, vrho, vs, vt, vl)
    for l in range(3):
        assert _rel(Hag[l], Hao[l]) < TOL, l
        if pol:
            assert _rel(Hbg[l], Hbo[l]) < TOL, l
    assert abs(Eg - Eo) < 1e-11 * abs(Eo)


def test_sadatom_eval_fxc_slater(hb):
    """DFTGrid::eval_Fxc with LDA exchange (the SAP functional, src/general/sap.h:40-43) for the sadatom cube."""
    ob, basis, og, gg = _sad_setup(hb)
    N = ob.Nrad()
    P = _sad_cube(N, 2, 5)
    H, Exc, Nel = gg.eval_Fxc(1, 0, P)
    o = og.eval_density(P)
    rho = o["rho"][:, 0]
    cx = -0.75 * (3.0 / np.pi) ** (1.0 / 3.0)
    exc = np.where(rho > 1e-12, cx * np.cbrt(np.maximum(rho, 0)), 0.0)
    vrho = np.where(rho > 1e-12, 4.0 / 3.0 * cx * np.cbrt(np.maximum(rho, 0)), 0.0)
    Ho, _, Eo = og.eval_fxc(exc, vrho[:, None])
    assert abs(Nel - o["Nel"]) < 1e-11 * abs(o["Nel"]) and abs(Exc - Eo) < 1e-11 * abs(Eo)
    for l in range(3):
        assert _rel(H[l], Ho[l]) < 1e-11, l


def test_fused_fock_build_device(hb):
    """hfq_fock_build_device: XC + J + K of one restricted build, device-resident, the grid density chain running
    next to the J/K kernels.  HF (x_func = -1): the reference's fock_builder still calls eval_Fxc, which then only
    integrates Nel (src/diatomic/main.cpp:396-401) -> Nel = Tr(P S), zero XC matrix.  x_func = 1: Slater exchange on
    the device == hfq_eval_fxc."""
    import torch
    basis = hb.DiatomicTwoDBasis(7, 7, 2.07, [6, 5, 4], 2).compute_tei()
    t = basis.tables
    n = basis.Nbf()
    grid = hb.DFTGrid(basis, 4 * 6 + 12)
    P = cases.random_density(n, 3, 17, cases.m_blocks(t.mval, t.Nrad, True))
    S = basis.overlap()
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a.T)).cuda()
    dP = dev(P)
    dJ, dK, dH = torch.empty_like(dP), torch.empty_like(dP), torch.full_like(dP, float("nan"))
    exc, nel = basis.fock_build_device(dP.data_ptr(), dJ.data_ptr(), dK.data_ptr(), 0.5, -1, 0, dH.data_ptr())
    J, K = basis.coulomb_exchange(P, 0.5)
    assert _rel(dJ.cpu().numpy().T, J) < 1e-13 and _rel(dK.cpu().numpy().T, K) < 1e-13
    assert abs(nel - np.sum(P * S)) < 1e-10 * abs(nel) and exc == 0.0
    assert not dH.cpu().numpy().any()
    exc, nel = basis.fock_build_device(dP.data_ptr(), dJ.data_ptr(), dK.data_ptr(), 0.5, 1, 0, dH.data_ptr())
    Href, eref, nref, _ = grid.eval_Fxc(1, 0, P)
    assert _rel(dH.cpu().numpy().T, Href) < 1e-13 and abs(exc - eref) < 1e-12 * abs(eref) and abs(nel - nref) < 1e-12 * nref
    assert _rel(dK.cpu().numpy().T, K) < 1e-13


@pytest.mark.parametrize("x_func,c_func", [(101, 130), (202, 231)])
def test_purem_eval_fxc_builtin(hb, x_func, c_func):
    """PBE and TPSS evaluated on the device on the diatomic pure-m grid (gradient / tau incl. the analytic phi terms)
    against the oracle grid fed with the oracle's functionals."""
    from oracle import xc
    ob, basis, og, gg = _setup_purem(hb, (3, 2), 2)
    n = ob.Nbf()
    P = cases.random_density(n, 3, 5, cases.m_blocks(ob.mval, ob.Nrad(), True))
    mgga = x_func == 202
    o = og.eval_density(P, None, True, mgga, False)
    if mgga:
        exc, vrho, vsigma, vtau = xc.evaluate_mgga([x_func, c_func], o["rho"][:, 0], o["sigma"][:, 0], o["tau"][:, 0], 1e-12)
        Ho, _, Eo = og.eval_fxc(exc, vrho[:, None], vsigma[:, None], vtau[:, None])
    else:
        exc, vrho, vsigma = xc.evaluate_sum([x_func, c_func], o["rho"][:, 0], o["sigma"][:, 0], 1e-12)
        Ho, _, Eo = og.eval_fxc(exc, vrho[:, None], vsigma[:, None])
    H, Exc, Nel, Ekin = gg.eval_Fxc(x_func, c_func, P)
    assert cases.relerr(H, Ho) < 1e-10 and abs(Exc - Eo) < 1e-11 * abs(Eo) and abs(Nel - o["Nel"]) < 1e-10
