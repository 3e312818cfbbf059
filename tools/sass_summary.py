#!/usr/bin/env python3
"""SASS mnemonic counts per kernel of the product library (cuobjdump -sass) -> profiles/<tag>_sass_summary.txt.
   python tools/sass_summary.py eicos_b200/libeicos_b200.so profiles/r02f_sass_summary.txt"""
import collections, re, subprocess, sys
lib, out = sys.argv[1], sys.argv[2]
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
cols = ["UBLKCP", "SYNCS", "LDGSTS", "LDGDEPBAR", "DEPBAR", "LDS", "STS", "LDG", "STG", "DFMA", "DMUL", "DADD", "MUFU", "DMMA", "HMMA", "BAR", "BRA", "BRX"]
kern, counts = None, collections.OrderedDict()
for line in txt.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = m.group(1)
        d = re.search(r"_ZN5eicos\d+([a-z_0-9]+?)ENS", name)
        kern = d.group(1) if d else name
        counts[kern] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and kern:
        op = m.group(1)
        counts[kern]["total"] += 1
        for c in cols:
            if op == c or op.startswith(c + "."):
                counts[kern][c] += 1
with open(out, "w") as f:
    f.write("SASS mnemonic counts per kernel of %s (cuobjdump -sass).\n" % lib)
    f.write("UBLKCP = cp.async.bulk (TMA bulk copy of the record / load-list chunks), SYNCS = mbarrier, LDGSTS = cp.async (data rows), DMMA: none (no dense fronts).\n\n")
    f.write("%-28s %7s" % ("kernel", "total") + "".join(" %7s" % c for c in cols) + "\n")
    for k in sorted(counts):
        f.write("%-28s %7d" % (k, counts[k]["total"]) + "".join(" %7d" % counts[k][c] for c in cols) + "\n")
print(open(out).read())
