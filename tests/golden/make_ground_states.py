"""Extracts the per-l electron counts of the neutral-atom ground states (spin-restricted spherical average) that
the reference freezes in its own sadatom sub-SCFs: `pbe_ground_states[118][4]`, src/diatomic/twodquadrature.cpp:26-145
("PBE ground states determined with 10 radial elements").  Writes helfem_b200/data/ground_states.json (the default
occupations of the batched SAP driver, helfem_b200/sap.py) -- data, cited, reproducible with this script.

    python tests/golden/make_ground_states.py [/root/reference]
"""
import json
import os
import re
import sys

ref = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
src = open(os.path.join(ref, "src/diatomic/twodquadrature.cpp")).read()
body = src.split("int pbe_ground_states[118][4] = {", 1)[1].split("};", 1)[0]
rows = re.findall(r"\{\s*(\d+)\s*,\s*(\d+)\s*,\s*(\d+)\s*,\s*(\d+)\s*\}\s*,?\s*//\s*(\d+)\s+(\w+)", body)
out = {"source": "src/diatomic/twodquadrature.cpp:26-145 (pbe_ground_states)", "occ": {}, "symbol": {}}
for s, p, d, f, Z, sym in rows:
    assert int(s) + int(p) + int(d) + int(f) == int(Z), (Z, s, p, d, f)
    out["occ"][Z] = [int(s), int(p), int(d), int(f)]
    out["symbol"][Z] = sym
assert len(out["occ"]) == 118
dst = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..", "helfem_b200", "data", "ground_states.json")
json.dump(out, open(dst, "w"), indent=0)
print("wrote", os.path.normpath(dst), len(out["occ"]), "elements")
