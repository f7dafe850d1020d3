import sys, os, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from helfem_b200 import sap
b = sap.SadatomBatchSCF([10, 36, 86, 24, 29, 46, 64])
r = b.run(verbose=True)
print(r["E"], b.converged, b.iterations)
np.save("gpurun_out/dbg_Pl_Ne.npy", b.Pl[0].cpu().numpy())
np.save("gpurun_out/dbg_tab_Ne.npy", b.sap_table(0))
