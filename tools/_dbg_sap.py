import sys, os, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from helfem_b200 import sap
zs = list(range(1, 87))
b = sap.SadatomBatchSCF(zs)
hist = []
orig = b.xc
r = b.run(maxit=int(sys.argv[1]) if len(sys.argv) > 1 else 150, verbose=False)
bad = [z for z, c in zip(zs, b.converged.cpu().numpy()) if not c]
print("iterations", b.iterations, "not converged:", bad, [b.occ_l[zs.index(z)] for z in bad])
print("errs", [float(x) for x in b.last_emax.cpu().numpy()[[zs.index(z) for z in bad]]] if hasattr(b, "last_emax") else None)
