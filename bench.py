#!/usr/bin/env python
"""bench.py -- J/K/Vxc Fock builds per second for N2 HF in prolate spheroidal coordinates.

Workload (BASELINE.json metric, configs[3]): N2, Rbond 2.07, lmax=30 for |m|<=6, 3 radial
elements x 15-node LIP (src/diatomic/main.cpp defaults, tests/cases.json diatomic-N2-hf-r
scaled to the north-star lmax/mmax), closed-shell density with the N2 occupation pattern
(5 sigma + pi+- doubly occupied; seeded synthetic orbitals).  One step = one Fock build as the
reference's fock_builder issues it (src/diatomic/main.cpp:385-426): XC = pmgrid.eval_Fxc(P) -- for
HF (x_func = -1) the density on the pure-m grid and its integral Nel, zero XC matrix --,
J = coulomb(P), K = exchange(P/2).

  python bench.py --gpus N --steps K --warmup W            our CUDA path
  python bench.py --impl reference ...                     the reference algorithm on host cores
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


class _BasisShape:
    def __init__(self, Nrad, lval, mval):
        self.Nrad, self.lval, self.mval = int(Nrad), np.asarray(lval), np.asarray(mval)
        self.Nbf = int(sum(self.Nrad - (1 if m != 0 else 0) for m in self.mval))


def n2_density(T, seed=42):
    """Synthetic closed-shell density with the N2 ground-state structure: 3 sigma_g (even l) and
    2 sigma_u (odd l) doubly-occupied orbitals in m=0, one pi_u orbital (odd l) in each of m=+1
    and m=-1 (same coefficients: +-m symmetric).  Orbitals are seeded random orthonormal vectors
    inside their (m, l-parity) subspace, like the g/u-symmetric SCF orbitals of a homonuclear
    diatomic."""
    rng = np.random.default_rng(seed)
    n = T.Nbf
    mval, lval = T.mval, T.lval
    sub = {}
    off = 0
    for m, l in zip(mval, lval):
        k = T.Nrad - (1 if m != 0 else 0)
        sub.setdefault((int(m), int(l) & 1), []).extend(range(off, off + k))
        off += k
    P = np.zeros((n, n), order="F")
    for (m, par, k) in ((0, 0, 3), (0, 1, 2), (1, 1, 1)):
        idx = np.array(sub[(m, par)])
        Q, _ = np.linalg.qr(rng.standard_normal((len(idx), k)))
        blk = 2.0 * Q @ Q.T
        P[np.ix_(idx, idx)] = blk
        if m:
            jdx = np.array(sub[(-m, par)])
            P[np.ix_(jdx, jdx)] = blk
    return P


class ClockSampler(threading.Thread):
    def __init__(self, gpu=0):
        super().__init__(daemon=True)
        self.gpu, self.rows, self.stop_flag = gpu, [], False

    def run(self):
        q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unsampled"]}
        sm = sorted(float(r[0]) for r in self.rows)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[3 + i] == "Active" for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(self.rows[0][1]), "reasons": reasons,
                "power_w_max": max(float(r[2]) for r in self.rows)}


def cpu_reference_build(C, P, kscale, blocks_per_thread=8, jstride=8, threads=None):
    """One Fock build (J = coulomb(P), K = exchange(kscale P)) of the reference algorithm on the host cores, on a
    bounded sample, extrapolated to the full build.

    K (src/diatomic/basis.cpp:1818-2089, OpenMP over output blocks): the block-norm screening runs once per build
    and is timed in full (call with zero output blocks); then `blocks_per_thread` x threads of the output blocks that
    receive density are computed (evenly spaced over the m-diagonal blocks, omp dynamic), and the marginal cost per
    block is scaled to all of them: t_K = t_screen + (t_sample - t_screen) * nblocks / nsample.
    J (basis.cpp:1627-1816, serial and unscreened in the reference): every phase is linear in the angular rows /
    channels visited, a 1/jstride sample is timed: t_J = jstride * t_sample.
    Returns (seconds per build, dict with the parts, the sampled K blocks for the parity check)."""
    from oracle import cjk
    nthr = cjk.use_all_cores() if threads is None else threads
    cjk.lib().jk_set_num_threads(int(nthr))
    Pd = C.expand(P)
    Pk = kscale * Pd
    mv = C.mval
    cand = [(j, k) for j in range(C.Nang) for k in range(C.Nang) if mv[j] == mv[k]]
    nsel = min(len(cand), max(1, blocks_per_thread * nthr))
    pick = np.linspace(0, len(cand) - 1, nsel).astype(int)
    sel = [cand[i] for i in pick]
    t0 = time.perf_counter()
    C.exchange_blocks(Pk, [], [])
    t_screen = time.perf_counter() - t0
    t0 = time.perf_counter()
    blk = C.exchange_blocks(Pk, [s[0] for s in sel], [s[1] for s in sel])
    t_sample = time.perf_counter() - t0
    tK = t_screen + max(t_sample - t_screen, 0.0) * len(cand) / nsel
    tJ = jstride * C.coulomb_timing_sample(Pd, jstride)
    parts = {"threads": int(nthr), "k_blocks_sampled": nsel, "k_blocks_total": len(cand), "k_screen_s": t_screen,
             "k_sample_s": t_sample, "k_build_s": tK, "j_stride": jstride, "j_build_s": tJ}
    return tK + tJ, parts, (sel, blk)


def cpu_sample_text(parts):
    return ("J+K of the reference algorithm (C restatement of src/diatomic/basis.cpp:1627-2089, -O3 -march=native): K = full "
            "block-norm screening + %d of %d density-carrying output blocks (%d per thread, omp dynamic over blocks like the "
            "reference), marginal block cost extrapolated; J (serial and unscreened in the reference) = 1/%d sample x %d"
            % (parts["k_blocks_sampled"], parts["k_blocks_total"], parts["k_blocks_sampled"] // max(parts["threads"], 1),
               parts["j_stride"], parts["j_stride"]))


def cpu_vxc_seconds(bval, lval, mval, P, lang, elements=None):
    """The HF build's eval_Fxc on the reference's pure-m grid (src/diatomic/dftgrid_purem.cpp:87-442, 675-700:
    materialised basis tables per element + density GEMMs; numpy/BLAS restatement oracle/dftgrid_purem.py) --
    seconds for all elements, extrapolated from `elements` if given; also returns Nel."""
    from oracle import diatomic as odi
    from oracle import dftgrid_purem as dp
    ob = odi.TwoDBasis(7, 7, 1.035, 15, 75, np.asarray(bval), np.asarray(lval), np.asarray(mval))   # grid tables only, no TEIs
    og = dp.PureMDFTGrid(ob, lang)
    Pd = ob.expand_boundaries(P)
    nel_tot = ob.radial.Nel()
    els = list(range(nel_tot)) if elements is None else list(elements)
    t0 = time.perf_counter()
    nel = 0.0
    for iel in els:
        og.compute_bf(iel)
        d = og.density(Pd, False, False, False)
        nel += float(np.sum(og.wtot * d["rho"]))
    return (time.perf_counter() - t0) * nel_tot / len(els), nel


def parity_of_blocks(C, sel, blk, Kdense, Jdense=None, P=None):
    """Relative Frobenius error of the GPU K on the CPU-sampled output blocks (and of J against the complete
    single-M Coulomb build of the oracle)."""
    pi = C.pure_idx()
    N = C.Nrad
    pure = np.zeros(C.Nang * N, dtype=bool)
    pure[pi] = True
    pos = np.cumsum(pure) - 1   # dummy index -> dense index
    num = den = 0.0
    for b, (j, k) in enumerate(sel):
        rows = pos[j * N:(j + 1) * N][pure[j * N:(j + 1) * N]]
        cols = pos[k * N:(k + 1) * N][pure[k * N:(k + 1) * N]]
        ref = blk[b][np.ix_(pure[j * N:(j + 1) * N], pure[k * N:(k + 1) * N])]
        got = Kdense[np.ix_(rows, cols)]
        num += float(np.sum((got - ref) ** 2))
        den += float(np.sum(ref ** 2))
    out = {"max_relerr_K": (num / den) ** 0.5 if den > 0 else None, "k_blocks": len(sel)}
    if Jdense is not None:
        Jref = C.coulomb(P, single_M=0)
        out["max_relerr_J"] = float(np.linalg.norm(Jdense - Jref) / np.linalg.norm(Jref))
        out["j_blocks"] = "all (complete J of the m-diagonal density)"
    return out


def make_density(kind, T):
    if kind == "n2":
        return n2_density(T), "synthetic closed-shell N2 structure (3 sigma_g + 2 sigma_u + pi_u+-, g/u-symmetric orbitals), seed 42"
    rng = np.random.default_rng(42)
    n = T.Nbf
    P = np.zeros((n, n), order="F")
    off = 0
    blocks = {}
    for m in T.mval:
        k = T.Nrad - (1 if m != 0 else 0)
        blocks.setdefault(int(m), []).extend(range(off, off + k))
        off += k
    for m, idx in blocks.items():
        idx = np.array(idx)
        if kind == "dense":   # dense random m-block-diagonal density (worst case of the screening)
            Q, _ = np.linalg.qr(rng.standard_normal((len(idx), 3)))
            P[np.ix_(idx, idx)] = 2.0 * Q @ Q.T
        else:                 # non-symmetric: general exchange path (no half storage)
            A = rng.standard_normal((len(idx), 3))
            B = rng.standard_normal((len(idx), 3))
            P[np.ix_(idx, idx)] = A @ B.T
    return P, {"dense": "dense random symmetric m-block-diagonal density, every (m, parity) sector pair of equal m active, seed 42",
               "nonsym": "non-symmetric random m-block-diagonal density (general exchange path), seed 42"}[kind]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--lmax", type=int, default=30)
    ap.add_argument("--mmax", type=int, default=6)
    ap.add_argument("--nelem", type=int, default=3)
    ap.add_argument("--cpu-blocks-per-thread", type=int, default=16,
                    help="exchange output blocks per host thread in the CPU sample (>= 4: no idle threads)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--density", default="n2", choices=["n2", "dense", "nonsym"],
                    help="n2 = the headline density; dense / nonsym = secondary figures (no CPU leg)")
    ap.add_argument("--profile-mode", action="store_true",
                    help="for runs under ncu only: device-resident steps, no e2e/peak/CPU legs, prints no bench line")
    args = ap.parse_args()
    if args.impl == "ours" and not args.profile_mode:
        args.warmup = max(args.warmup, 3)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    lang = 4 * args.lmax + 12   # pure-m grid of the reference driver (src/diatomic/main.cpp:316)

    metric = "J/K/Vxc Fock builds/s (N2 HF)"
    workload = "N2 HF diatomic Fock build (eval_Fxc on the pure-m grid [HF: density + Nel] + coulomb + exchange), Rbond=2.07, " \
               "lmax=%d |m|<=%d, nelem=%d x 15-node LIP, grid %d nu x %d mu" % (args.lmax, args.mmax, args.nelem, lang, 75 * args.nelem)
    config = {"workload": workload, "density": None,
              "symmetry": "per-m (reference default --symmetry=1, absm_symmetric off)",
              "l2": "inputs larger than L2 (P/J/K 1.76 GB each, work buffers > 10 GB)",
              "sharding": "owner computes: exchange units (output sector pair, radial element pair) dealt longest-first to the least "
                          "loaded rank; ONE in-place ncclAllGather of the compact result issued by the library (hfq_comm_init); J and "
                          "the grid density replicated"}

    if args.impl == "reference":
        if rank != 0:
            return
        # Inputs: the integral caches (what compute_tei() leaves in memory) come from tools/export_caches.py run in
        # a subprocess -- this process holds the oracle only.  See that file for why the numpy restatement is not
        # used for them at lmax = 30.
        import tempfile
        from oracle import cjk
        npz = os.path.join(tempfile.gettempdir(), "hfq_caches_%d_%d_%d.npz" % (args.lmax, args.mmax, args.nelem))
        if not os.path.exists(npz):
            subprocess.check_call([sys.executable, os.path.join(ROOT, "tools", "export_caches.py"), npz, "7", "7", "2.07",
                                   str(args.lmax), str(args.mmax), str(args.nelem)])
        C = cjk.DiatomicCaches.from_npz(npz)
        bval = np.load(npz)["bval"]
        shape = _BasisShape(C.Nrad, C.lval, C.mval)
        P = n2_density(shape)
        config["density"] = "synthetic closed-shell N2 structure (3 sigma_g + 2 sigma_u + pi_u+-, g/u-symmetric orbitals), seed 42"
        # every step = one bounded sample of the build, sized so that 25 steps end within a few minutes
        times, parts = [], None
        for it in range(args.warmup + args.steps):
            tB, parts, _ = cpu_reference_build(C, P, 0.5, blocks_per_thread=8, jstride=16)
            tX, _ = cpu_vxc_seconds(bval, C.lval, C.mval, P, lang, elements=[args.nelem // 2])
            parts["vxc_build_s"] = tX
            if it >= args.warmup:
                times.append(tB + tX)
        tB = float(np.median(times)) if times else float("nan")
        val = 1.0 / tB
        line = {"metric": metric, "value": val, "unit": "builds/s", "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tB, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "impl": "reference",
                "config": config,
                "cpu_baseline": {"value": val, "unit": "builds/s", "cores": parts["threads"], "kind": "port",
                                 "sample": cpu_sample_text(parts) + "; Vxc (HF: grid density + Nel) = 1 of %d elements x %d, numpy/BLAS"
                                           % (args.nelem, args.nelem), "parts": parts,
                                 "inputs": "integral caches written by tools/export_caches.py in a subprocess"},
                "e2e": {"value": val, "unit": "builds/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return

    import helfem_b200 as hb
    from helfem_b200 import build as hb_build
    import torch
    import torch.distributed as dist
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    hb_build.build()

    hb.set_host_threads(world=world)   # torchrun exports OMP_NUM_THREADS=1: give each rank its share of the cores
    t0 = time.time()
    T = hb.Tables.diatomic(7, 7, 2.07, [args.lmax] * (args.mmax + 1), args.nelem)
    t_setup = time.time() - t0
    tei_dev = None
    if world == 1 and not args.profile_mode:
        # the same setup with the in-element kernels computed on the GPU (hfq_tables_diatomic_device), timed next to
        # the host path and compared with it channel by channel (reconstructed kernels W = B sigma B^T)
        hb.Tables.diatomic(7, 7, 2.07, [2, 2], 1, device=local)     # CUDA context + module load outside the timing
        t0 = time.time()
        Td = hb.Tables.diatomic(7, 7, 2.07, [args.lmax] * (args.mmax + 1), args.nelem, device=local)
        t_dev = time.time() - t0
        worst = 0.0
        for ilm in range(0, T.nlm, max(1, T.nlm // 12)):
            for iel in range(T.Nel):
                _, _, Bh, sh = T.block(ilm, iel)
                _, _, Bd, sd = Td.block(ilm, iel)
                Wh = (Bh * sh[None, :]) @ Bh.T
                worst = max(worst, float(np.abs(Wh - (Bd * sd[None, :]) @ Bd.T).max() / np.abs(Wh).max()))
        tei_dev = {"seconds": t_dev, "max_relerr_W_vs_host": worst, "channels_compared": len(range(0, T.nlm, max(1, T.nlm // 12))) * T.Nel}
        del Td
    basis = hb.TablesBasis(T, device=local)
    n = T.Nbf
    P, config["density"] = make_density(args.density, T)
    t0 = time.time()
    basis._context()
    t_upload = time.time() - t0
    if world > 1:
        basis.comm_init()   # NCCL communicator inside the library; torch.distributed only hands the id around
    hb.DFTGrid(basis, lang)
    nel_ref = float(np.sum(P * T.one_electron()[0]))

    # device-resident inputs/outputs (column-major n x n == transposed row-major torch tensors)
    dP = torch.from_numpy(np.ascontiguousarray(P.T)).cuda()
    dJ = torch.empty_like(dP)
    dK = torch.empty_like(dP)
    stream = torch.cuda.current_stream().cuda_stream

    # host matrices of the end-to-end leg: one set, visible to every rank (multi-GPU: POSIX shared memory), page-locked
    def shared_host(name, fill=None):
        if world == 1:
            t = torch.empty((n, n), dtype=torch.float64).pin_memory()
        else:
            path = "/dev/shm/hfq_bench_%s_%s" % (os.environ.get("MASTER_PORT", "0"), name)
            if rank == 0:
                torch.from_file(path, shared=True, size=n * n, dtype=torch.float64)   # creates the segment
            dist.barrier()
            t = torch.from_file(path, shared=True, size=n * n, dtype=torch.float64).view(n, n)
            assert torch.cuda.cudart().cudaHostRegister(t.data_ptr(), n * n * 8, 0) in (0, None, torch.cuda.cudart().cudaError.success)
        if fill is not None and rank == 0:
            t.copy_(fill)
        return t

    hP = shared_host("P", torch.from_numpy(np.ascontiguousarray(P.T)))
    hJ = shared_host("J")
    hK = shared_host("K")
    if world > 1:
        dist.barrier()

    acc = {"ms_pack": 0.0, "ms_unpack": 0.0, "ms_total": 0.0, "ms_fold": 0.0, "ms_tgemm": 0.0, "ms_offdiag": 0.0, "alg_fold": 0.0, "alg_tgemm": 0.0, "alg_offdiag": 0.0,
           "launches": 0.0, "launches_tgemm": 0.0, "flops_tgemm": 0.0, "flops_fold": 0.0, "flops_offdiag": 0.0, "n": 0}

    def step_device(collect=False):
        # one Fock build: XC (HF: grid density + Nel), J = coulomb(P), K = exchange(P/2); with a communicator the
        # exchange is sharded and completed by the library: results are whole on every rank
        exc, nel = basis.fock_build_device(dP.data_ptr(), dJ.data_ptr(), dK.data_ptr(), 0.5, -1, 0, None, 1e-12, stream)
        assert abs(nel - nel_ref) < 1e-9 * abs(nel_ref), (nel, nel_ref)
        if collect:
            tm = basis.last_timings()
            for k in acc:
                if k in tm:
                    acc[k] += tm[k]
            acc["launches"] += 4   # grid chain: pack, two GEMM stages, point kernel
            acc["n"] += 1

    nel_h, exc_h = ctypes.c_double(), ctypes.c_double()

    def step_host():
        # the fock_builder's three calls through the fused host entry point (host matrices in and out)
        hb._check(hb.lib().hfq_fock_build(basis._context(), hP.data_ptr(), n, 0.5, hJ.data_ptr(), n, hK.data_ptr(), n, -1, 0,
                                          None, n, ctypes.byref(exc_h), ctypes.byref(nel_h), 1e-12))
        if world > 1:
            dist.barrier()   # the shared matrices are complete when every rank has returned

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing
    for _ in range(args.warmup):
        step_device()
    sampler = ClockSampler(local)
    sampler.start()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step_device(collect=True)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    sampler.stop_flag = True
    sampler.join(timeout=2)
    tms = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
    ms_step = float(tms.item()) / args.steps
    value = 1e3 / ms_step

    # ---- secondary figure: the DFT flavour of the grid work (density -> Slater exchange on the device -> assembly of
    #      the XC matrix), device-resident, with the HBM model of SURVEY.md section 8d (separable evaluation)
    vxc_dft = None
    if world == 1:
        dH = torch.empty_like(dP)
        ex, ne, ek = ctypes.c_double(), ctypes.c_double(), ctypes.c_double()

        def vxc_call(xf=1, cf=0):
            hb._check(hb.lib().hfq_eval_fxc(basis._context(), xf, cf, dP.data_ptr(), n, None, 0, dH.data_ptr(), n, None, 0,
                                            ctypes.byref(ex), ctypes.byref(ne), ctypes.byref(ek), 1, 1e-12))

        def time_vxc(xf, cf):
            for _ in range(3):
                vxc_call(xf, cf)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(args.steps):
                vxc_call(xf, cf)
            torch.cuda.synchronize()
            return (time.perf_counter() - t0) / args.steps, ex.value

        t_pbe, exc_pbe = time_vxc(101, 130)      # GGA: density + gradient, PBE x + c on the device, v_rho and v_sigma assembled
        t_tpss, exc_tpss = time_vxc(202, 231)    # meta-GGA: + tau and v_tau
        t_vxc, _ = time_vxc(1, 0)
        npts = lang * 75 * args.nelem
        mv = np.asarray(T.mval)
        coupled = int(sum((mv == m).sum() ** 2 for m in set(mv.tolist())))
        alg_bytes = (coupled * T.Nrad ** 2 * 8          # the coupled (same-m) blocks of P
                     + n * n * 8                         # H: the call defines the whole dense matrix
                     + args.nelem * 75 * 15 * 8 * 2      # radial tables (value, derivative-free LDA: F only, both stages)
                     + T.Nang * lang * 8                 # angular table
                     + (1 + 1) * 8 * npts + 4 * 8 * npts)   # rho in / v_rho out, weights and scale factors
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
        hbm = peaks.get("hbm_gbs", 6650.0)
        vxc_dft = {"workload": "N2 LDA-exchange Vxc build on the pure-m grid (hfq_eval_fxc, x_func = 1), %d points, device-resident" % npts,
                   "ms_per_build": 1e3 * t_vxc, "builds_per_s": 1.0 / t_vxc,
                   "roofline": {"bound": "hbm", "achieved": alg_bytes / t_vxc / 1e9, "peak": hbm, "unit": "GB/s",
                                "frac": alg_bytes / t_vxc / 1e9 / hbm, "alg_bytes": alg_bytes,
                                "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback of B200_PROFILING.md",
                                "note": "separable-evaluation byte model of SURVEY.md 8d; 93 % of the bytes are the dense H the call "
                                        "must define (10 071 of 130 321 blocks are non-zero); the chain is 10 small launches "
                                        "(latency-bound), not a bandwidth-bound stream"},
                   "Exc": ex.value, "Nel": ne.value,
                   "other_functionals_ms_per_build": {"gga_x_pbe+gga_c_pbe": 1e3 * t_pbe, "mgga_x_tpss+mgga_c_tpss": 1e3 * t_tpss},
                   "Exc_pbe": exc_pbe, "Exc_tpss": exc_tpss}
        del dH

    if args.profile_mode:
        print("profile-mode: %.2f ms/step (not a bench value)" % ms_step)
        return
    # ---- end-to-end through the host-pointer C ABI (page-locked host matrices, copies inside the timed region)
    if rank == 0:
        hJ.fill_(float("nan"))   # the call must define every element (copied blocks + host zero-fill)
        hK.fill_(float("nan"))
    barrier()
    for _ in range(max(args.warmup, 1)):   # the same W untimed steps as the device-resident leg
        step_host()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_host()
    torch.cuda.synchronize()
    tt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    e2e_val = args.steps / float(tt.item())
    tm = basis.last_timings()
    hd = torch.tensor([tm["h2d_bytes"], tm["d2h_bytes"]], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(hd)   # bytes moved by all ranks
    h2d_bytes, d2h_bytes = int(hd[0].item()), int(hd[1].item())
    spec_hits = int(tm["speculative_hits"]) if world == 1 else None
    if rank == 0:
        ek = float((hK.cuda() - dK).abs().max() / dK.abs().max())
        ej = float((hJ.cuda() - dJ).abs().max() / dJ.abs().max())
        assert ek < 1e-12 and ej < 1e-12, "host and device paths disagree: %g %g" % (ek, ej)
        assert abs(nel_h.value - nel_ref) < 1e-9 * abs(nel_ref)

    per_rank = None
    if world > 1:
        # sharded kernels per rank (ms per build and executed Gflop): shows how level the owner-computes assignment is
        nst_ = max(acc["n"], 1)
        mine = torch.tensor([acc["ms_fold"] / nst_, acc["ms_tgemm"] / nst_, acc["ms_offdiag"] / nst_, acc["ms_total"] / nst_,
                             acc["flops_fold"] / nst_ / 1e9, acc["flops_tgemm"] / nst_ / 1e9, acc["flops_offdiag"] / nst_ / 1e9],
                            dtype=torch.float64, device="cuda")
        allr = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(allr, mine)
        per_rank = {"columns": ["fold_ms", "gemm_ms", "cross_element_ms", "exchange_path_ms", "fold_gflop", "gemm_gflop", "cross_gflop"],
                    "ranks": [[round(float(x), 3) for x in t.tolist()] for t in allr]}
    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    # ---- FP64 peak on this box, same method as MEASURED_PEAKS.json's bf16 number (cuBLAS, burst)
    a = torch.randn(8192, 8192, dtype=torch.float64, device="cuda")
    b = torch.randn(8192, 8192, dtype=torch.float64, device="cuda")
    for _ in range(2):
        a @ b
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(5):
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
        a @ b
        s1.record()
        torch.cuda.synchronize()
        best = min(best, s0.elapsed_time(s1))
    fp64_peak = 2 * 8192 ** 3 / best / 1e9  # TFLOP/s
    del a, b

    nst = max(acc["n"], 1)
    nl = max(acc["launches_tgemm"], 1)
    ach = acc["alg_tgemm"] / (acc["ms_tgemm"] * 1e-3) / 1e12 if acc["ms_tgemm"] > 0 else 0.0
    # DRAM traffic of the dominant kernel is a profiler number, not measurable here: it is read from the committed
    # per-state file written from the last ncu --set full capture of this workload (1 GPU, headline density)
    traffic, traffic_note = None, None
    tf = os.path.join(ROOT, "profiles", "current_traffic.json")
    if os.path.exists(tf) and world == 1 and args.density == "n2" and (args.lmax, args.mmax, args.nelem) == (30, 6, 3):
        tj = json.load(open(tf))
        traffic, traffic_note = tj.get("k_tgemm_ws_dram_bytes_per_launch"), tj.get("note")
    roofline = {"bound": "tensor", "kernel": "k_tgemm_ws (in-element exchange, FP64 DMMA)", "achieved": ach, "peak": fp64_peak,
                "unit": "TFLOP/s", "frac": ach / fp64_peak, "traffic": traffic, "traffic_note": traffic_note,
                "peak_source": "cuBLAS DGEMM 8192^3 measured in this run (MEASURED_PEAKS.json has no FP64 entry; the DMMA "
                               "microbenchmark of profiles/r01_fp64_peak_microbench.txt reads 37.1)",
                "alg_flops_per_launch": acc["alg_tgemm"] / nl, "ms_per_launch": acc["ms_tgemm"] / nl,
                "executed_tflops": acc["flops_tgemm"] / (acc["ms_tgemm"] * 1e-3) / 1e12 if acc["ms_tgemm"] > 0 else 0.0,
                "step_share": {"fold_ms": acc["ms_fold"] / nst, "gemm_ms": acc["ms_tgemm"] / nst,
                               "cross_element_ms": acc["ms_offdiag"] / nst,
                               "norms_pack_ms": acc["ms_pack"] / nst, "reduce_gather_unpack_ms": acc["ms_unpack"] / nst,
                               "exchange_path_ms": acc["ms_total"] / nst, "step_ms": ms_step},
                "alg_tflops_all_kernels": (acc["alg_fold"] + acc["alg_tgemm"] + acc["alg_offdiag"]) / nst / (ms_step * 1e-3) / 1e12}

    cpu_baseline, parity = None, None
    if not args.no_cpu_baseline and args.density == "n2":
        from oracle import cjk
        C = cjk.DiatomicCaches.from_tables(T)
        tB, parts, (sel, blk) = cpu_reference_build(C, P, 0.5, blocks_per_thread=args.cpu_blocks_per_thread, jstride=8)
        tX, nel_cpu = cpu_vxc_seconds(T.bval, T.lval, T.mval, P, lang)
        parts["vxc_build_s"] = tX
        # the sample is also the parity check of THIS run's result at full size: the GPU K on the sampled blocks, the
        # complete J and the grid integral Nel against the oracle (1e-12, BASELINE.json north_star)
        Kh = dK.cpu().numpy().T
        Jh = dJ.cpu().numpy().T
        parity = parity_of_blocks(C, sel, blk, Kh, Jh, P)
        parity["relerr_Nel"] = abs(nel_h.value - nel_cpu) / abs(nel_cpu)
        parity["tolerance"] = 1e-12
        assert parity["max_relerr_K"] < 1e-12 and parity["max_relerr_J"] < 1e-12 and parity["relerr_Nel"] < 1e-12, \
            "GPU result differs from the oracle: %s" % parity
        # one-thread figure (the reference's own test setting, tests/cases.json defaults.env OMP_NUM_THREADS=1)
        tB1, parts1, _ = cpu_reference_build(C, P, 0.5, blocks_per_thread=48, jstride=16, threads=1)
        cpu_baseline = {"value": 1.0 / (tB + tX), "unit": "builds/s", "cores": parts["threads"], "kind": "port",
                        "sample": cpu_sample_text(parts) + "; Vxc (HF: grid density + Nel) in full, numpy/BLAS restatement of "
                                  "src/diatomic/dftgrid_purem.cpp", "parts": parts,
                        "one_thread_value": 1.0 / (tB1 + tX), "one_thread_parts": parts1}

    line = {"metric": metric, "value": value, "unit": "builds/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
            "e2e": {"value": e2e_val, "unit": "builds/s", "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": d2h_bytes,
                    "call": ("hfq_fock_build, dense page-locked host matrices in and out. P: the row ranges that were non-zero in "
                             "the previous call are uploaded and the build starts on them, while host threads verify that "
                             "every other element of the caller's P is exactly zero (mismatch = rebuild from the complete "
                             "upload; speculative_hits counts the calls that did not need it). J: the complete dense matrix "
                             "(a share of its columns that follows the measured finish times of the two sides) travels over the "
                             "otherwise idle device-to-host link while K is being built; K and the rest of J: only the row "
                             "ranges of the non-zero blocks cross PCIe, the rest of the host matrix is zero-filled by host "
                             "threads while the GPU computes") if world == 1 else
                            ("hfq_fock_build with a communicator: ONE set of host matrices in POSIX shared memory, page-locked by "
                             "every rank; each rank uploads its 1/N column slice of P over its own PCIe link, one in-place "
                             "ncclAllGather over NVLink completes P on every GPU, sharded build, each rank copies back the non-zero "
                             "row ranges of its column slice of J and K and zero-fills the rest; host barrier at the end"),
                    "speculative_hits": spec_hits},
            "gpu_launches": int(acc["launches"]), "roofline": roofline, "cpu_baseline": cpu_baseline, "parity": parity,
            "vxc_dft": vxc_dft, "per_rank": per_rank,
            "clocks": sampler.summary(),
            "setup": {"host_compute_tei_s": t_setup, "device_compute_tei": tei_dev, "device_upload_s": t_upload, "Nbf": n,
                      "channels": T.nlm}}
    print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
