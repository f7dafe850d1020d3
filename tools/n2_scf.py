"""SCF time-to-convergence of N2 RHF (second half of BASELINE.json's metric) with everything on the device:
helfem_b200/scf.py::DeviceRHF around hfq_fock_build_device.  Prints one JSON line.
    python tools/n2_scf.py [--lmax 30 --mmax 6 --nelem 3]"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--lmax", type=int, default=30)
    ap.add_argument("--mmax", type=int, default=6)
    ap.add_argument("--nelem", type=int, default=3)
    args = ap.parse_args()
    import helfem_b200 as hb
    from helfem_b200.scf import DeviceRHF
    t0 = time.perf_counter()
    basis = hb.DiatomicTwoDBasis(7, 7, 2.07, [args.lmax] * (args.mmax + 1), args.nelem, tei_on_device=True).compute_tei()
    basis._context()
    t_setup = time.perf_counter() - t0
    t0 = time.perf_counter()
    scf = DeviceRHF(basis, 7, Enucr=49.0 / 2.07)
    t_init = time.perf_counter() - t0
    r = scf.run(verbose=True)
    print(json.dumps({"workload": "N2 RHF, Rbond 2.07, lmax=%d |m|<=%d, nelem=%d, Nbf=%d, core-Hamiltonian guess, damped Roothaan + DIIS, "
                                  "convergence 1e-10 Eh / 1e-7 commutator" % (args.lmax, args.mmax, args.nelem, basis.Nbf()),
                      "E_total": r["E"], "iterations": r["iterations"], "scf_seconds": r["seconds"],
                      "fock_build_seconds": r["fock_build_seconds"], "setup_compute_tei_on_device_and_upload_s": t_setup,
                      "one_electron_and_sinvh_s": t_init, "occupations_per_m_block": scf.occ_per_block, "Nel": r["Nel"]}))


if __name__ == "__main__":
    main()
