"""Secondary measurement: DFT-grid path (density + XC-matrix assembly) on the GPU next to the
reference algorithm (numpy oracle: materialised tables + BLAS GEMMs, as src/*/dftgrid*.cpp).
Prints one JSON line per configuration.  Functional-independent: synthetic v arrays.
  python tests/measure/bench_vxc.py [kr|n2]"""
import json, sys, time
import numpy as np
sys.path.insert(0, '.')
import helfem_b200 as hb
from tests import cases


def run(name):
    if name == "kr":
        desc = "Kr atom, lmax=mmax=2, 5 elements x 15-node LIP, meta-GGA (tau+laplacian) grid 20x20 angular x 75 radial"
        ob = cases.oracle_atomic(36, 2, 2, 5)
        basis = hb.AtomicTwoDBasis(36, 2, 2, 5).compute_tei()
        from oracle import dftgrid_atomic as dg
        lang = 4 * 2 + 12
        og, gg = dg.AtomicDFTGrid(ob, lang, lang), hb.DFTGrid(basis, lang, lang)
        P = cases.random_density(ob.Nbf(), 18, 1, cases.m_blocks(ob.mval, ob.Nrad(), False))
        ofx = lambda *a: og.eval_fxc(ob.Nbf(), *a)
    else:
        lmax, mmax = (30, 6) if name == "n2" else (12, 3)
        desc = "N2 diatomic pure-m grid, lmax=%d |m|<=%d, 3 elements, meta-GGA (tau+laplacian), %d nu x 225 mu points" % (lmax, mmax, 4 * lmax + 12)
        basis = hb.DiatomicTwoDBasis(7, 7, 2.07, [lmax] * (mmax + 1), 3).compute_tei()
        from oracle import dftgrid_purem as dp, diatomic as odi, fem as ofem
        T = basis.tables
        ob = odi.TwoDBasis(7, 7, 1.035, 15, 75, T.bval, T.lval, T.mval)    # oracle object for the grid only (no TEIs needed)
        lang = 4 * lmax + 12
        og, gg = dp.PureMDFTGrid(ob, lang), hb.DFTGrid(basis, lang)
        from bench import n2_density
        P = n2_density(T)
        ofx = lambda *a: og.eval_fxc(*a)
    n = basis.Nbf()
    rng = np.random.default_rng(7)
    N = gg.N
    exc = rng.uniform(-1, 0, N); vrho = rng.uniform(-1, 0, (N, 1)); vs = rng.uniform(0, 1e-2, (N, 1))
    vt = rng.uniform(0, 1e-2, (N, 1)); vl = rng.uniform(0, 1e-2, (N, 1))
    times = []
    for it in range(5):
        t0 = time.perf_counter()
        d = gg.density(P, None, 7)
        H, _, E = gg.fxc(exc, vrho, vs, vt, vl)
        times.append(time.perf_counter() - t0)
    tg = float(np.median(times[1:]))
    t0 = time.perf_counter()
    do = og.eval_density(P, None, True, True, True)
    Ho, _, Eo = ofx(exc, vrho, vs, vt, vl)
    tc = time.perf_counter() - t0
    err = cases.relerr(H, Ho)
    print(json.dumps({"config": desc, "Nbf": n, "points": N, "gpu_s_per_eval": tg, "cpu_oracle_s_per_eval": tc,
                      "speedup": tc / tg, "relerr_H": err, "relerr_rho": float(np.max(np.abs(d["rho"] - do["rho"])) / np.max(np.abs(do["rho"]))),
                      "note": "GPU time includes H2D of P and D2H of densities and H (host-pointer C ABI); CPU = numpy oracle with BLAS on all cores"}))


if __name__ == "__main__":
    for name in (sys.argv[1:] or ["kr", "n2small"]):
        run(name)
