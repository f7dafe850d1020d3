import sys, time, numpy as np
sys.path.insert(0, '.')
import helfem_b200 as hb
lmax = int(sys.argv[1]) if len(sys.argv) > 1 else 30
mmax = int(sys.argv[2]) if len(sys.argv) > 2 else 6
t = time.time()
T = hb.Tables.diatomic(7, 7, 2.07, [lmax] * (mmax + 1), 3)
print('tables %.1fs Nbf %d nlm %d ranks %d..%d' % (time.time() - t, T.Nbf, T.nlm, T.ranks.min(), T.ranks.max()), flush=True)
if '--gpu' in sys.argv:
    from tests import cases
    b = hb.TablesBasis(T)
    n = T.Nbf
    blocks = cases.m_blocks(T.mval, T.Nrad, True)
    P = cases.random_density(n, 3, 1, blocks)
    t = time.time(); b._context(); print('context %.1fs' % (time.time() - t), flush=True)
    for sym in (False, True):
        b.set_absm_symmetric(sym)
        for it in range(2):
            t = time.time(); K = b.exchange(P); dt = time.time() - t
            tm = b.last_timings()
            print('absm', sym, 'exchange wall %.3fs' % dt, {k: (round(v, 2) if v < 1e6 else '%.3e' % v) for k, v in tm.items()}, flush=True)
            fl = tm['flops_fold'] + tm['flops_tgemm'] + tm['flops_offdiag']
            print('   TF/s fold %.2f tgemm %.2f offdiag %.2f' % (tm['flops_fold'] / tm['ms_fold'] / 1e9, tm['flops_tgemm'] / tm['ms_tgemm'] / 1e9, tm['flops_offdiag'] / max(tm['ms_offdiag'], 1e-9) / 1e9), flush=True)
    t = time.time(); J = b.coulomb(P); print('coulomb wall %.3fs' % (time.time() - t), b.last_timings(), flush=True)
    print('K sym err', np.abs(K - K.T).max() / np.abs(K).max(), 'J sym', np.abs(J - J.T).max() / np.abs(J).max())
