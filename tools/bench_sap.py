"""Config 5 of BASELINE.json: batched SAP potentials for all elements Z = 1-86 (gen_sap_table workload), distributed
by element over the GPUs (replicas only: no collective).  Prints one JSON line.

    python tools/bench_sap.py [--zmax 86] [--out DIR]                       one GPU
    torchrun --nproc-per-node N tools/bench_sap.py                          N GPUs, elements dealt round-robin

One unit = one converged spin-restricted LDA-exchange SCF of a spherically averaged atom + its effective-potential
table (376 x 9, result_<El>.dat).  The CPU comparator is the oracle restatement of the reference's Fock build
(oracle/scf.py::sadatom_rks, numpy/BLAS) on a sample of atoms.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--zmax", type=int, default=86)
    ap.add_argument("--out", default=None)
    ap.add_argument("--cpu-sample", type=int, default=3, help="atoms of the CPU comparator sample (0: skip)")
    ap.add_argument("--profile", action="store_true", help="phase timers (synchronise the device around every phase)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    import torch
    import torch.distributed as dist
    import helfem_b200 as hb
    from helfem_b200 import sap
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    hb.set_host_threads(world=world)
    zs_all = list(range(1, args.zmax + 1))
    zs = sap.elements_of_rank(zs_all, rank, world)
    t0 = time.perf_counter()
    batch = sap.SadatomBatchSCF(zs, device=local)
    torch.cuda.synchronize()
    t_setup = time.perf_counter() - t0
    batch.profile = args.profile
    batch.run(maxit=3)            # warm-up (allocations, kernel attributes)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    res = batch.run()
    torch.cuda.synchronize()
    t_scf = time.perf_counter() - t0
    t0 = time.perf_counter()
    out = args.out or os.path.join(ROOT, "gpurun_out", "sap_results")
    paths = batch.write_results(out)
    t_tab = time.perf_counter() - t0
    tt = torch.tensor([t_scf, t_tab, float(batch.iterations), float(int(batch.converged.all()))], dtype=torch.float64, device="cuda")
    if world > 1:
        mx = tt.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        mn = tt.clone()
        dist.all_reduce(mn, op=dist.ReduceOp.MIN)
        t_scf, t_tab, iters, conv = float(mx[0]), float(mx[1]), int(mx[2]), bool(mn[3] > 0)
    else:
        iters, conv = batch.iterations, bool(batch.converged.all())
    if rank != 0:
        dist.destroy_process_group()
        return
    cpu = None
    if args.cpu_sample:
        sys.path.insert(0, ROOT)
        from tests.test_gpu_sap import _oracle_atom
        occ, sym = sap.ground_state_occupations()
        sample = [z for z in (2, 36, min(86, args.zmax)) if z in zs_all][:args.cpu_sample]
        per, errs = [], []
        for z in sample:
            t0 = time.perf_counter()
            ob, lmax, ro = _oracle_atom(z, occ[z])
            per.append(time.perf_counter() - t0)
            if z in zs:
                a = zs.index(z)
                errs.append(abs(res["E"][a] - ro["E"]) / max(1.0, abs(ro["E"])))
        cpu = {"atoms_per_s": 1.0 / float(np.mean(per)), "sample": "oracle SCF (numpy/BLAS, all host cores) of Z = %s, basis setup included" % sample,
               "max_relerr_E_vs_gpu": max(errs) if errs else None}
    line = {"metric": "SAP atoms/s (spherically averaged LDA-x SCF + effective-potential table)", "value": len(zs_all) / (t_scf + t_tab),
            "unit": "atoms/s", "n_gpus": world, "elements": len(zs_all), "scf_s": t_scf, "tables_s": t_tab, "setup_s": t_setup,
            "iterations": iters, "all_converged": conv,
            "refilled_rank0": {str(z): [round(x, 4) for x in v] for z, v in batch.refilled.items()},
            "refill_note": "atoms whose tabulated (PBE) frozen configuration has no bound LDA-x Aufbau solution: per-l counts from a finite-temperature SCF with one chemical potential (kT 0.02 -> 0.005 Eh), then frozen again at T = 0", "scaling": "strong (elements dealt round-robin, no collective)",
            "kernel_launches_native": batch.launches, "phase_seconds": batch.timing if batch.profile else None, "cpu_baseline": cpu, "results": os.path.relpath(out, ROOT),
            "E_Rn" if args.zmax >= 86 and 86 in zs else "E_last": float(res["E"][zs.index(max(zs))])}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
