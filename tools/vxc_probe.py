"""Time the DFT-grid calls of the N2 workload with page-locked host buffers (wall clock per call).
  python tools/vxc_probe.py [lmax mmax]"""
import sys, time, ctypes
import numpy as np, torch
sys.path.insert(0, '.')
import helfem_b200 as hb
from bench import n2_density

lmax, mmax = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (30, 6)
T = hb.Tables.diatomic(7, 7, 2.07, [lmax] * (mmax + 1), 3)
basis = hb.TablesBasis(T)
n = T.Nbf
L = hb.lib()
ctx = basis._context()
hb._check(L.hfq_grid_attach(ctx, 4 * lmax + 12, 1))
N = int(L.hfq_grid_npoints(ctx))
P = n2_density(T)
hP = torch.from_numpy(np.ascontiguousarray(P.T)).pin_memory()
hH = torch.empty((n, n), dtype=torch.float64).pin_memory()
pin = lambda k: torch.zeros(k * N, dtype=torch.float64).pin_memory()
rho, sig, tau, lap, w = pin(1), pin(1), pin(1), pin(1), pin(1)
rng = np.random.default_rng(7)
mk = lambda lo, hi: torch.from_numpy(rng.uniform(lo, hi, N)).pin_memory()
exc, vrho, vs, vt, vl = mk(-1, 0), mk(-1, 0), mk(0, 1e-2), mk(0, 1e-2), mk(0, 1e-2)
nel, ekin, Exc = ctypes.c_double(), ctypes.c_double(), ctypes.c_double()
p = lambda t: ctypes.c_void_p(t.data_ptr())
for flags, name in ((0, "LDA"), (1, "GGA"), (7, "meta-GGA+lapl")):
    for it in range(4):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        hb._check(L.hfq_grid_density(ctx, p(hP), n, None, n, flags, p(rho), p(sig), p(tau), p(lap), p(w), ctypes.byref(nel), ctypes.byref(ekin)))
        t1 = time.perf_counter()
        hb._check(L.hfq_grid_fxc(ctx, flags, 1, p(exc), p(vrho), p(vs) if flags & 1 else None, p(vt) if flags & 2 else None,
                                 p(vl) if flags & 4 else None, p(hH), n, None, n, ctypes.byref(Exc)))
        t2 = time.perf_counter()
    print("%s: Nbf %d points %d  density %.1f ms  fxc %.1f ms  Nel %.6f" % (name, n, N, 1e3 * (t1 - t0), 1e3 * (t2 - t1), nel.value))
    # the same with every matrix / point array resident on the device
    dP, dH = hP.cuda(), torch.empty((n, n), dtype=torch.float64, device="cuda")
    d = {k: v.cuda() for k, v in dict(rho=rho, sig=sig, tau=tau, lap=lap, w=w, exc=exc, vrho=vrho, vs=vs, vt=vt, vl=vl).items()}
    for it in range(4):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        hb._check(L.hfq_grid_density(ctx, p(dP), n, None, n, flags, p(d["rho"]), p(d["sig"]), p(d["tau"]), p(d["lap"]), p(d["w"]), ctypes.byref(nel), ctypes.byref(ekin)))
        t1 = time.perf_counter()
        hb._check(L.hfq_grid_fxc(ctx, flags, 1, p(d["exc"]), p(d["vrho"]), p(d["vs"]) if flags & 1 else None, p(d["vt"]) if flags & 2 else None,
                                 p(d["vl"]) if flags & 4 else None, p(dH), n, None, n, ctypes.byref(Exc)))
        torch.cuda.synchronize(); t2 = time.perf_counter()
    err = float((dH.cpu() - hH).abs().max() / hH.abs().max())
    print("   device-resident: density %.2f ms  fxc %.2f ms  (H vs host-path %.1e)" % (1e3 * (t1 - t0), 1e3 * (t2 - t1), err))
