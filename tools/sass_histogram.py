"""Opcode histogram per kernel of libhelfemqc_b200.so (cuobjdump -sass): the evidence that the FP64 work runs on the
tensor pipe (DMMA.8x8x4), that the GEMM ring uses bulk copies + mbarriers (UBLKCP, SYNCS) and the fold / cross-element
kernels cp.async (LDGSTS).  tcgen05 has no f64 kind, so no UTC*MMA is expected in this library.
    python tools/sass_histogram.py > profiles/rNN_sass_histogram.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "helfem_b200", "libhelfemqc_b200.so")
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
kern, hist = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
        hist[kern] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and kern:
        hist[kern][m.group(1).split(".")[0]] += 1
keys = ["DMMA", "DFMA", "DMUL", "DADD", "UBLKCP", "SYNCS", "LDGSTS", "LDG", "LDS", "STS", "STG", "BAR", "NOP", "HMMA", "UTCHMMA"]
print("# cuobjdump -sass %s : static instruction counts per kernel" % os.path.relpath(so, ROOT))
print("%-52s" % "kernel" + "".join("%8s" % k for k in keys) + "   total")
for k, h in hist.items():
    if not sum(h.values()):
        continue
    print("%-52s" % k.replace("hfq::dev::", "").replace("hfq::", "")[:52] + "".join("%8d" % h.get(x, 0) for x in keys) + "%8d" % sum(h.values()))
