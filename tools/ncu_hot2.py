#!/usr/bin/env python3
"""Hot instructions of one ncu capture: per-instruction stall samples by reason (source page, CSV)."""
import csv, subprocess, sys
rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]
col = {h: i for i, h in enumerate(hdr)}
reasons = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
data = [r for r in rows[2:] if len(r) == len(hdr)]
tot = sum(int(r[col["# Samples"]]) for r in data)
print("samples", tot, " instructions executed", sum(int(r[col["Instructions Executed"]]) for r in data))
agg = {k: sum(int(r[col[k]]) for r in data) for k in reasons}
print("by reason:", {k[6:]: round(v / tot, 3) for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v})
for r in sorted(data, key=lambda r: -int(r[col["# Samples"]]))[:top]:
    rs = {k[6:]: int(r[col[k]]) for k in reasons if int(r[col[k]])}
    rs = dict(sorted(rs.items(), key=lambda kv: -kv[1])[:3])
    print("%6d %5.1f%%  %-60s %s" % (int(r[col["# Samples"]]), 100.0 * int(r[col["# Samples"]]) / tot, r[col["Source"]].strip()[:60], rs))
