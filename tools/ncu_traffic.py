#!/usr/bin/env python3
"""Average DRAM traffic per launch (dram__bytes_read.sum + dram__bytes_write.sum) of each kernel in an
ncu --csv log -> the JSON bench.py reads for roofline.traffic.
   python tools/ncu_traffic.py traffic.csv out.json"""
import collections, csv, json, sys
rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
hdr = rows[hi]
ki, mi, ui, vi = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Unit"), hdr.index("Metric Value")
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
tot, launches = collections.Counter(), collections.defaultdict(set)
for r in rows[hi + 1:]:
    if len(r) <= vi or not r[mi].startswith("dram__bytes"):
        continue
    k = r[ki].split("(")[0].split("::")[-1]
    tot[k] += float(r[vi].replace(",", "")) * scale.get(r[ui], 1.0)
    launches[k].add(r[0])
out = {k: tot[k] / max(len(launches[k]), 1) for k in tot}
out["_note"] = "bytes per launch, averaged over %s launches of `bench.py --steps 1 --warmup 1` under ncu (dram__bytes_read.sum + dram__bytes_write.sum)" % \
    {k: len(v) for k, v in launches.items()}
out["_workload"] = sys.argv[3] if len(sys.argv) > 3 else "mpc02"
out["_batch_per_gpu"] = int(sys.argv[4]) if len(sys.argv) > 4 else 65536
json.dump(out, open(sys.argv[2], "w"), indent=1)
print(json.dumps(out))
