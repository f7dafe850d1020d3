"""Write the integral caches of a diatomic basis to an .npz (inputs of the CPU reference arm of bench.py).

    python tools/export_caches.py OUT.npz Z1 Z2 Rbond lmax mmax nelem

The caches (disjoint P/Q integrals, in-element factor B / sigma, prefactors) are the INPUTS of a Fock build, what the
reference's compute_tei() leaves in memory (src/diatomic/basis.cpp:1382-1547).  They are produced here by this
repository's host-side setup (C++, no GPU involved; checked against the numpy restatement of the reference in
tests/test_host.py::test_diatomic_setup_matches_oracle) because the numpy restatement needs > 20 minutes at
lmax = 30.  bench.py --impl reference runs this file in a SUBPROCESS, so the timed process holds the oracle only.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    out, Z1, Z2, Rbond, lmax, mmax, nelem = sys.argv[1:8]
    import helfem_b200 as hb
    from helfem_b200 import build as hb_build
    hb_build.build()
    T = hb.Tables.diatomic(int(Z1), int(Z2), float(Rbond), [int(lmax)] * (int(mmax) + 1), int(nelem))
    d = {"Nrad": T.Nrad, "efirst": T.efirst, "en": T.en, "lval": T.lval, "mval": T.mval, "lmL": T.lmL, "lmM": T.lmM,
         "pref": T.pref, "ranks": T.ranks, "bval": T.bval}
    sm, bg, B, sg = [], [], [], []
    for ilm in range(T.nlm):
        for e in range(T.Nel):
            b = T.block(ilm, e)
            sm.append(np.concatenate([b[0][c].ravel(order="F") for c in range(2)]))
            bg.append(np.concatenate([b[1][c].ravel(order="F") for c in range(2)]))
            B.append(np.asarray(b[2]).ravel(order="F"))
            sg.append(np.asarray(b[3]).ravel())
    d["small"], d["big"], d["B"], d["sigma"] = map(np.concatenate, (sm, bg, B, sg))
    np.savez(out, **d)


if __name__ == "__main__":
    main()
