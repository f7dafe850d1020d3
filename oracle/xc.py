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
    else:
        raise ValueError("functional id %d is not restated in the oracle" % func_id)
    f = n * e
    outs = [e, sp.diff(f, n)] + ([sp.diff(f, s)] if gga else [])
    fn = sp.lambdify((n, s), outs, modules="numpy")
    return fn, gga


def is_gga(func_id):
    return func_id in (XC_GGA_X_PBE, XC_GGA_C_PBE)


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
