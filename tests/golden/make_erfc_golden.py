"""Golden values of the erfc Green's function Phi_L(Xi, xi) from the REFERENCE's own implementation
(libhelfem/src/erfc_expn.cpp compiled into oracle/_ref/liberfc_ref.so by oracle/Makefile).
Run in the build container (needs /root/reference):  python tests/golden/make_erfc_golden.py
Writes tests/golden/erfc_phi_ref.json."""
import ctypes
import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
so = os.path.join(os.path.dirname(os.path.dirname(HERE)), "oracle", "_ref", "liberfc_ref.so")
lib = ctypes.CDLL(so)
rng = np.random.default_rng(20261017)
rows = []
for n in range(0, 7):
    for scale in (0.05, 0.5, 2.0, 6.0):
        Xi = rng.uniform(0.0, scale, 24)
        xi = rng.uniform(0.0, scale, 24)
        Xi[:4] = [0.0, 1e-9, 0.4, 0.5]          # branch boundaries of Phi (erfc_expn.cpp:225-235)
        xi[:4] = [0.0, 0.0, 0.4, 0.25]
        out = np.empty_like(Xi)
        vp = ctypes.c_void_p
        rc = lib.ref_erfc_phi(out.ctypes.data_as(vp), ctypes.c_uint(n), Xi.ctypes.data_as(vp), xi.ctypes.data_as(vp),
                              ctypes.c_long(len(Xi)))
        assert rc == 0
        rows += [[n, float(a).hex(), float(b).hex(), float(v).hex()] for a, b, v in zip(Xi, xi, out)]
json.dump({"source": "libhelfem/src/erfc_expn.cpp (Phi<double>)", "columns": ["n", "Xi", "xi", "Phi"], "rows": rows},
          open(os.path.join(HERE, "erfc_phi_ref.json"), "w"))
print(len(rows), "values")
