"""Extract the reference's tabulated Gaunt / modified-Gaunt known answers.

Source: /root/reference/src/general/gaunt_test.cpp (pairs of
``val=helfem::gaunt::<fn>(args); ref=<value>;`` lines).  Run once in the build
container; the output tests/golden/gaunt_ref.json is committed so the tests
never need /root/reference at run time.
"""
import json
import re
import sys

src = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/src/general/gaunt_test.cpp"
txt = open(src).read()
pat = re.compile(r"val=helfem::gaunt::(\w+)\(([-\d, ]+)\);\s*ref=([-+.\deE]+);")
out = []
for fn, args, ref in pat.findall(txt):
    out.append({"fn": fn, "args": [int(a) for a in args.split(",")], "ref": float(ref)})
json.dump(out, open("tests/golden/gaunt_ref.json", "w"))
print(len(out), "entries")
