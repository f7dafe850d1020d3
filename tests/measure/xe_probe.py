"""Config 2 of BASELINE.json: Xe atom UHF, 20 radial elements x 15-node LIP, lmax=mmax=2 (Nbf=2511):
GPU J and K for both spin densities against the numpy oracle."""
import sys, time, numpy as np
sys.path.insert(0, '.')
import helfem_b200 as hb
from tests import cases
t0 = time.time()
basis = hb.AtomicTwoDBasis(54, 2, 2, 20).compute_tei()
print('product setup %.1fs Nbf %d' % (time.time() - t0, basis.Nbf()), flush=True)
n = basis.Nbf(); T = basis.tables
blocks = cases.m_blocks(T.mval, T.Nrad, False)
Pa = cases.random_density(n, 6, 1234, blocks); Pb = cases.random_density(n, 5, 1235, blocks)
basis._context()
for it in range(3):
    t0 = time.perf_counter()
    J = basis.coulomb(Pa + Pb); Ka = basis.exchange(Pa); Kb = basis.exchange(Pb)
    dt = time.perf_counter() - t0
print('GPU UHF Fock build (J + 2K, host API) %.1f ms' % (1e3 * dt), basis.last_timings(), flush=True)
if '--check' in sys.argv:
    t0 = time.time(); ob = cases.oracle_atomic(54, 2, 2, 20); print('oracle setup %.1fs' % (time.time() - t0), flush=True)
    t0 = time.time(); Jo = ob.coulomb(Pa + Pb); Ko = ob.exchange(Pa); tc = time.time() - t0
    print('oracle J+K(a) %.1fs  relerr J %.2e K %.2e' % (tc, cases.relerr(J, Jo), cases.relerr(Ka, Ko)))
