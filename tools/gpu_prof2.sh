#!/bin/bash
# ncu evidence for the row-program kernels: launch list + full sets at a small batch
TAG=${1:-prof}
B=${2:-4096}
OUT=gpurun_out/$TAG
mkdir -p $OUT
SMALL="python bench.py --batch $B --steps 1 --warmup 1 --no-e2e --no-cpu-baseline"
echo "== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -k regex:eicos_ --csv --log-file $OUT/launches.csv $SMALL > $OUT/launches.log 2>&1
python - <<PY
import csv, collections
rows = list(csv.reader(open("$OUT/launches.csv")))
hdr = None
agg = collections.defaultdict(lambda: [0, 0.0, 0.0])
for r in rows:
    if "Kernel Name" in r:
        hdr = r; continue
    if hdr is None or len(r) != len(hdr): continue
    d = dict(zip(hdr, r))
    k = d["Kernel Name"].split("(")[0]
    v = float(d["Metric Value"].replace(",", ""))
    if d["Metric Name"] == "gpu__time_duration.sum":
        agg[k][0] += 1; agg[k][1] += v / 1e6 if d["Metric Unit"] in ("nsecond", "ns") else v
    else:
        agg[k][2] += v
for k, (n, ms, inst) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%-28s launches %4d  total %9.2f ms  avg %8.3f ms  inst/launch %.3e" % (k, n, ms, ms / max(n, 1), inst / max(n, 1)))
PY
for K in eicos_solve_kkt eicos_residuals eicos_ldl_factor; do
  echo "== ncu full: $K"
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:^$K\$ -s 4 -c 1 -f -o $OUT/prof_$K $SMALL > $OUT/prof_$K.log 2>&1
done
ls -la $OUT
