import sys, time, numpy as np, torch
sys.path.insert(0, '.')
import helfem_b200 as hb
from bench import n2_density
T = hb.Tables.diatomic(7, 7, 2.07, [30] * 7, 3)
b = hb.TablesBasis(T); n = T.Nbf
P = n2_density(T)
dP = torch.from_numpy(np.ascontiguousarray(P.T)).cuda(); dPh = (0.5 * dP).contiguous()
dJ = torch.empty_like(dP); dK = torch.empty_like(dP)
st = torch.cuda.current_stream().cuda_stream
for it in range(4):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    b.coulomb_device(dP.data_ptr(), dJ.data_ptr(), st); torch.cuda.synchronize(); t1 = time.perf_counter()
    tj = b.last_timings()
    b.exchange_device(dPh.data_ptr(), dK.data_ptr(), 0, 1, st); torch.cuda.synchronize(); t2 = time.perf_counter()
    tk = b.last_timings()
    print('J wall %.2f ms (gpu total %.2f: pack %.2f fold %.2f radial %.2f unfold %.2f unpack %.2f) | K wall %.2f ms (gpu total %.2f: pack %.2f fold %.2f gemm %.2f off %.2f unpack %.2f)' % (
        1e3 * (t1 - t0), tj['ms_total'], tj['ms_pack'], tj['ms_fold'], tj['ms_offdiag'], tj['ms_tgemm'], tj['ms_unpack'],
        1e3 * (t2 - t1), tk['ms_total'], tk['ms_pack'], tk['ms_fold'], tk['ms_tgemm'], tk['ms_offdiag'], tk['ms_unpack']))
