#!/bin/bash
# ncu launch list (durations per kernel) at a given batch. Usage: bash tools/gpu_launches.sh <tag> [batch] [workers]
TAG=$1; B=${2:-4096}; W=${3:-1}
OUT=gpurun_out/$TAG; mkdir -p $OUT
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:eicos_ --csv --log-file $OUT/launches.csv \
  python bench.py --batch $B --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --workers $W > $OUT/launches.log 2>&1
python - <<PY
import csv, collections
rows = list(csv.reader(open("$OUT/launches.csv")))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
hdr = rows[hi]; ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
agg = collections.OrderedDict()
for r in rows[hi + 1:]:
    if len(r) <= vi: continue
    try: v = float(r[vi].replace(",", ""))
    except ValueError: continue
    a = agg.setdefault(r[ki].split("(")[0], [0, 0.0]); a[0] += 1; a[1] += v
tot = sum(a[1] for a in agg.values())
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:28s} n={a[0]:4d} total={a[1]/1e6:9.2f} ms avg={a[1]/a[0]/1e6:8.3f} ms share={a[1]/tot:6.3f}")
PY
