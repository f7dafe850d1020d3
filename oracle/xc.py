"""Exchange-correlation functionals for the ORACLE (test infrastructure only; the product leaves the functional
evaluation to the caller's libxc, as the reference does: src/general/dftgrid_common.cpp:98-255).

libxc is not available in this image, so the handful of functionals that the reference's recorded energies use
(tests/refs/ci.json: lda_x-lda_c_vwn, gga_x_pbe-gga_c_pbe) are restated from their published definitions, spin
unpolarised, and differentiated symbolically (sympy) -- exc, vrho, vsigma in libxc's conventions
(exc per particle; vsigma = d(n exc)/d sigma, sigma = |grad n|^2; zero output below the density threshold that the
reference sets with xc_func_set_dens_threshold, dftgrid_common.cpp:131).  Constants are libxc's
(lda_c_vwn.c: VWN5 paramagnetic set; gga_x_pbe.c: kappa = 0.8040, mu = beta pi^2/3; gga_c_pbe.c: beta =
0.06672455060314922, gamma = (1 - ln 2)/pi^2 on top of the "modified" PW92 of lda_c_pw.c).

Pinned through the oracle SCF on the reference's recorded energies (tests/test_oracle.py): He / Be LDA, He PBE
(atomic grid), He LDA (spherically averaged grid), H2 LDA (pure-m grid).
"""
import functools

import numpy as np

# libxc functional ids (xc_funcs.h)
XC_LDA_X, XC_LDA_C_VWN, XC_GGA_X_PBE, XC_GGA_C_PBE = 1, 7, 101, 130
XC_MGGA_X_TPSS, XC_MGGA_C_TPSS = 202, 231


@functools.lru_cache(maxsize=None)
def _compiled(func_id):
    import sympy as sp
    n, s = sp.symbols("n sigma", positive=True)
    pi = sp.pi
    rs = (3 / (4 * pi * n)) ** sp.Rational(1, 3)
    ex_unif = -sp.Rational(3, 4) * (3 / pi) ** sp.Rational(1, 3) * n ** sp.Rational(1, 3)   # per particle
    gga = False
    if func_id == XC_LDA_X:
        e = ex_unif
    elif func_id == XC_LDA_C_VWN:
        A, b, c, x0 = sp.Float("0.0310907"), sp.Float("3.72744"), sp.Float("12.9352"), sp.Float("-0.10498")
        x = sp.sqrt(rs)
        X = lambda y: y * y + b * y + c
        Q = sp.sqrt(4 * c - b * b)
        at = sp.atan(Q / (2 * x + b))
        e = A * (sp.log(x * x / X(x)) + 2 * b / Q * at
                 - b * x0 / X(x0) * (sp.log((x - x0) ** 2 / X(x)) + 2 * (b + 2 * x0) / Q * at))
    elif func_id == XC_GGA_X_PBE:
        gga = True
        kappa = sp.Float("0.8040")
        mu = sp.Float("0.06672455060314922") * pi ** 2 / 3
        kF = (3 * pi ** 2 * n) ** sp.Rational(1, 3)
        s2 = s / (2 * kF * n) ** 2
        e = ex_unif * (1 + kappa - kappa / (1 + mu * s2 / kappa))
    elif func_id == XC_GGA_C_PBE:
        gga = True
        beta = sp.Float("0.06672455060314922")
        gamma = (1 - sp.log(2)) / pi ** 2
        a, a1 = sp.Float("0.0310907"), sp.Float("0.21370")
        b1, b2, b3, b4 = sp.Float("7.5957"), sp.Float("3.5876"), sp.Float("1.6382"), sp.Float("0.49294")
        ec = -2 * a * (1 + a1 * rs) * sp.log(1 + 1 / (2 * a * (b1 * sp.sqrt(rs) + b2 * rs + b3 * rs ** sp.Rational(3, 2) + b4 * rs ** 2)))
        kF = (3 * pi ** 2 * n) ** sp.Rational(1, 3)
        ks = sp.sqrt(4 * kF / pi)
        t2 = s / (2 * ks * n) ** 2
        Aa = beta / gamma / (sp.exp(-ec / gamma) - 1)
        H = gamma * sp.log(1 + beta / gamma * t2 * (1 + Aa * t2) / (1 + Aa * t2 + Aa ** 2 * t2 ** 2))
        e = ec + H
    elif func_id in (XC_MGGA_X_TPSS, XC_MGGA_C_TPSS):
        return _compiled_tpss(func_id)
    else:
        raise ValueError("functional id %d is not restated in the oracle" % func_id)
    f = n * e
    outs = [e, sp.diff(f, n)] + ([sp.diff(f, s)] if gga else [])
    fn = sp.lambdify((n, s), outs, modules="numpy")
    return fn, gga


def _compiled_tpss(func_id):
    """TPSS meta-GGA (Tao, Perdew, Staroverov, Scuseria, PRL 91, 146401 (2003)), spin unpolarised, in libxc's
    parametrisation (mgga_x_tpss.c: b = 0.40, c = 1.59096, e = 1.537, kappa = 0.804, mu = 0.21951; mgga_c_tpss.c:
    d = 2.8, C(0, 0) = 0.53, z = min(tau_W / tau, 1), PBE correlation with beta = 0.06672455060314922 on the "modified"
    PW92 incl. its fully polarised branch for the one-spin term).  Returns (fn(n, sigma, tau) -> [exc, vrho, vsigma,
    vtau], "mgga").  Pinned through the oracle SCF on atomic-He-mgga-r of the reference's tests/refs/ci.json."""
    import sympy as sp
    n, s, tau = sp.symbols("n sigma tau", positive=True)
    pi, R = sp.pi, sp.Rational
    if func_id == XC_MGGA_X_TPSS:
        kappa, b, c, e, mu = sp.Float("0.804"), sp.Float("0.40"), sp.Float("1.59096"), sp.Float("1.537"), sp.Float("0.21951")
        mu_ge = R(10, 81)
        p = s / (4 * (3 * pi ** 2) ** R(2, 3) * n ** R(8, 3))
        z = s / (8 * n * tau)
        alpha = (tau - s / (8 * n)) / (R(3, 10) * (3 * pi ** 2) ** R(2, 3) * n ** R(5, 3))
        qb = R(9, 20) * (alpha - 1) / sp.sqrt(1 + b * alpha * (alpha - 1)) + 2 * p / 3
        x = ((mu_ge + c * z ** 2 / (1 + z ** 2) ** 2) * p + R(146, 2025) * qb ** 2
             - R(73, 405) * qb * sp.sqrt(R(1, 2) * (R(9, 25) * z ** 2 + p ** 2)) + mu_ge ** 2 / kappa * p ** 2
             + 2 * sp.sqrt(e) * mu_ge * R(9, 25) * z ** 2 + e * mu * p ** 3) / (1 + sp.sqrt(e) * p) ** 2
        en = -R(3, 4) * (3 / pi) ** R(1, 3) * n ** R(1, 3) * (1 + kappa - kappa / (1 + x / kappa))
    else:
        beta = sp.Float("0.06672455060314922")
        gamma = (1 - sp.log(2)) / pi ** 2

        def pw92(rs, a, a1, b1, b2, b3, b4):
            return -2 * a * (1 + a1 * rs) * sp.log(1 + 1 / (2 * a * (b1 * sp.sqrt(rs) + b2 * rs + b3 * rs ** R(3, 2) + b4 * rs ** 2)))

        def pbe_c(ntot, sig, polarised):
            rs = (3 / (4 * pi * ntot)) ** R(1, 3)
            if polarised:   # zeta = 1: ferromagnetic PW92 branch, phi = 2^(-1/3)
                ec = pw92(rs, *[sp.Float(v) for v in ("0.01554535", "0.20548", "14.1189", "6.1977", "3.3662", "0.62517")])
                phi = 2 ** (-R(1, 3))
            else:
                ec = pw92(rs, *[sp.Float(v) for v in ("0.0310907", "0.21370", "7.5957", "3.5876", "1.6382", "0.49294")])
                phi = 1
            kF = (3 * pi ** 2 * ntot) ** R(1, 3)
            t2 = sig / (2 * phi * sp.sqrt(4 * kF / pi) * ntot) ** 2
            A = beta / gamma / (sp.exp(-ec / (gamma * phi ** 3)) - 1)
            return ec + gamma * phi ** 3 * sp.log(1 + beta / gamma * t2 * (1 + A * t2) / (1 + A * t2 + A ** 2 * t2 ** 2))

        z = sp.Min(s / (8 * n * tau), 1)
        e_pbe = pbe_c(n, s, False)
        e_tilde = sp.Max(pbe_c(n / 2, s / 4, True), e_pbe)
        rev = e_pbe * (1 + sp.Float("0.53") * z ** 2) - (1 + sp.Float("0.53")) * z ** 2 * e_tilde
        en = rev * (1 + sp.Float("2.8") * rev * z ** 3)
    f = n * en
    fn = sp.lambdify((n, s, tau), [en, sp.diff(f, n), sp.diff(f, s), sp.diff(f, tau)], modules="numpy")
    return fn, "mgga"


def is_gga(func_id):
    """needs the density gradient (GGAs and meta-GGAs)"""
    return func_id in (XC_GGA_X_PBE, XC_GGA_C_PBE, XC_MGGA_X_TPSS, XC_MGGA_C_TPSS)


def is_mgga(func_id):
    return func_id in (XC_MGGA_X_TPSS, XC_MGGA_C_TPSS)


def evaluate_mgga(func_ids, rho, sigma, tau, thr=1e-12):
    """Sum of the functionals at the points (spin unpolarised): (exc, vrho, vsigma, vtau); LDAs and GGAs contribute
    zero vtau.  Zero below the density threshold."""
    rho = np.asarray(rho, dtype=float).ravel()
    ok = rho >= thr
    r = np.where(ok, rho, 1.0)
    sg = np.maximum(np.where(ok, np.asarray(sigma, dtype=float).ravel(), 0.0), 1e-40)
    tt = np.where(ok, np.asarray(tau, dtype=float).ravel(), 1.0)
    tot = [np.zeros_like(r) for _ in range(4)]
    for fid in func_ids:
        fn, kind = _compiled(fid)
        out = fn(r, sg, tt) if kind == "mgga" else fn(r, sg)
        for k, o in enumerate(out):
            tot[k] += np.where(ok, np.broadcast_to(np.asarray(o, dtype=float), r.shape), 0.0)
    return tuple(tot)


def evaluate(func_id, rho, sigma=None, thr=1e-12):
    """(exc, vrho, vsigma or None) at the points, spin unpolarised (rho = total density)."""
    fn, gga = _compiled(func_id)
    rho = np.asarray(rho, dtype=float).ravel()
    ok = rho >= thr
    r = np.where(ok, rho, 1.0)
    sg = np.where(ok, np.asarray(sigma, dtype=float).ravel(), 0.0) if gga else np.zeros_like(r)
    sg = np.maximum(sg, 1e-40)   # the symbolic derivative is regular at sigma -> 0; avoid 0/0 in intermediate terms
    out = fn(r, sg)
    res = [np.where(ok, np.broadcast_to(np.asarray(o, dtype=float), r.shape), 0.0) for o in out]
    return res[0], res[1], (res[2] if gga else None)


def evaluate_sum(func_ids, rho, sigma=None, thr=1e-12):
    """Sum over the functionals of a method (x + c), as compute_xc accumulates them."""
    exc = np.zeros(np.asarray(rho).size)
    vrho = np.zeros_like(exc)
    vsigma = None
    for fid in func_ids:
        e, v, vs = evaluate(fid, rho, sigma, thr)
        exc += e
        vrho += v
        if vs is not None:
            vsigma = vs if vsigma is None else vsigma + vs
    return exc, vrho, vsigma
