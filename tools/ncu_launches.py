#!/usr/bin/env python3
"""Per-kernel totals of an ncu launch list (csv with gpu__time_duration.sum, smsp__inst_executed.sum, dram__bytes_*)."""
import collections
import csv
import json
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = None
agg = collections.defaultdict(lambda: collections.defaultdict(float))
cnt = collections.Counter()
for r in rows:
    if "Kernel Name" in r:
        hdr = r
        continue
    if hdr is None or len(r) != len(hdr):
        continue
    d = dict(zip(hdr, r))
    k = d["Kernel Name"].split("(")[0]
    v = float(d["Metric Value"].replace(",", ""))
    u = d["Metric Unit"]
    if d["Metric Name"] == "gpu__time_duration.sum":
        cnt[k] += 1
        v = v / 1e6 if u in ("nsecond", "ns") else (v / 1e3 if u in ("usecond", "us") else v)
    if d["Metric Name"].startswith("dram"):
        v = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1) * v
    agg[k][d["Metric Name"]] += v
total = sum(m["gpu__time_duration.sum"] for m in agg.values())
out = []
for k, m in sorted(agg.items(), key=lambda kv: -kv[1]["gpu__time_duration.sum"]):
    n = cnt[k]
    rec = {"kernel": k, "launches": n, "total_ms": round(m["gpu__time_duration.sum"], 2), "avg_ms": round(m["gpu__time_duration.sum"] / n, 3),
           "share": round(m["gpu__time_duration.sum"] / total, 4), "inst_per_launch": m["smsp__inst_executed.sum"] / n,
           "dram_read_GB_per_launch": round(m["dram__bytes_read.sum"] / n / 1e9, 2), "dram_write_GB_per_launch": round(m["dram__bytes_write.sum"] / n / 1e9, 2)}
    out.append(rec)
    print("%-24s launches %4d total %9.2f ms avg %8.3f ms share %.3f inst/launch %.3e dram GB/launch r %.2f w %.2f"
          % (k, n, rec["total_ms"], rec["avg_ms"], rec["share"], rec["inst_per_launch"], rec["dram_read_GB_per_launch"], rec["dram_write_GB_per_launch"]))
if len(sys.argv) > 2:
    json.dump(out, open(sys.argv[2], "w"), indent=1)
