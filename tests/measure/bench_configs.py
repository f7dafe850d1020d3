"""Secondary measurements on BASELINE.json's other configurations: C1 Ne RHF, C2 Xe UHF, C3 Kr
(bare K, erfc rs_exchange).  GPU: device-resident builds through the C ABI (CUDA-event timings of the library);
CPU: the C restatement of the reference's loops (oracle/csrc/jk_oracle.c, all host cores) on the same caches.
One JSON line per configuration.   python tests/measure/bench_configs.py [ne xe kr]"""
import json, sys, time
import numpy as np, torch
sys.path.insert(0, '.')
import helfem_b200 as hb
from oracle import cjk
from tests import cases

CFG = {"ne": ("C1 Ne RHF, lmax=mmax=0, 5 elements", 10, 0, 0, 5, 5),
       "xe": ("C2 Xe UHF, lmax=mmax=2, 20 elements", 54, 2, 2, 20, 27),
       "kr": ("C3 Kr, lmax=mmax=2, 5 elements", 36, 2, 2, 5, 18)}


def gpu_ms(fn, reps=10):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return 1e3 * (time.perf_counter() - t0) / reps


def run(name):
    desc, Z, lmax, mmax, nelem, nocc = CFG[name]
    t0 = time.time()
    basis = hb.AtomicTwoDBasis(Z, lmax, mmax, nelem).compute_tei()
    T = basis.tables
    n = basis.Nbf()
    setup = time.time() - t0
    blocks = cases.m_blocks(T.mval, T.Nrad, False)
    Pa = cases.random_density(n, max(1, nocc // len(blocks)), 1234, blocks)
    dP = torch.from_numpy(np.ascontiguousarray(Pa.T)).cuda()
    dJ, dK = torch.empty_like(dP), torch.empty_like(dP)
    st = torch.cuda.current_stream().cuda_stream
    tJ = gpu_ms(lambda: basis.coulomb_device(dP.data_ptr(), dJ.data_ptr(), st))
    tK = gpu_ms(lambda: basis.exchange_device(dP.data_ptr(), dK.data_ptr(), 0, 1, st))
    tJK = gpu_ms(lambda: basis.coulomb_exchange_device(dP.data_ptr(), dJ.data_ptr(), dK.data_ptr(), 1.0, 0, 1, st))
    tm = basis.last_timings()
    C = cjk.AtomicCaches.from_tables(T)
    C.exchange(Pa)
    t0 = time.perf_counter(); Kc = C.exchange(Pa); tc = time.perf_counter() - t0
    err = cases.relerr(dK.cpu().numpy().T, Kc)
    out = {"config": desc, "Nbf": n, "setup_s": setup, "gpu_ms": {"coulomb": tJ, "exchange": tK, "fused_J+K": tJK},
           "kernel_ms_of_last_fused": {k: float(tm[k]) for k in ("ms_pack", "ms_fold", "ms_tgemm", "ms_offdiag", "ms_unpack")},
           "cpu_exchange_s": tc, "cpu_cores": cjk.num_threads(), "exchange_speedup": 1e3 * tc / tK, "relerr_K_vs_cpu": err,
           "note": "device-resident P/J/K, wall clock per call incl. launch + host plan lookup; CPU = C port of the "
                   "reference loops on the product's own caches"}
    if name == "kr":
        mu = 0.3
        t0 = time.time(); basis.compute_erfc(mu); out["erfc_setup_s"] = time.time() - t0
        rs = basis._rs
        trs = gpu_ms(lambda: rs.exchange_device(dP.data_ptr(), dK.data_ptr(), 0, 1, st))
        out["gpu_ms"]["rs_exchange_erfc"] = trs
        ob = cases.oracle_atomic(Z, lmax, mmax, nelem)
        t0 = time.time(); ob.compute_erfc(mu); out["oracle_erfc_setup_s"] = time.time() - t0
        t0 = time.perf_counter(); Ko = ob.rs_exchange(Pa); out["numpy_oracle_rs_exchange_s"] = time.perf_counter() - t0
        out["relerr_rsK_vs_oracle"] = cases.relerr(dK.cpu().numpy().T, Ko)
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    for nm in (sys.argv[1:] or ["ne", "xe", "kr"]):
        run(nm)
