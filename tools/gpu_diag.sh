#!/bin/bash
# diagnostics: paired solves as two CTAs vs one two-job pass at the full batch, and one ncu capture of each form
TAG=${1:-diag}
OUT=gpurun_out/$TAG
mkdir -p $OUT
for PAIR in 0 1; do
  echo "== bench batch 65536 EICOS_PAIR_SOLVES=$PAIR"
  EICOS_PAIR_SOLVES=$PAIR timeout 600 python bench.py --batch 65536 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline 2>$OUT/bench_pair$PAIR.err | tee $OUT/bench_pair$PAIR.json | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: print(l[:300]); continue
    r = d['roofline']
    print({k: d[k] for k in ('value','ms_per_step')}, 'solve frac %.3f avg_ms %.2f' % (r['frac'], r['avg_launch_ms']), d['kernel_ms'], d['kkt_phase_share'])
"
done
B="python bench.py --batch 65536 --steps 1 --warmup 0 --no-e2e --no-cpu-baseline"
echo "== ncu launch list (full batch)"
timeout 900 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:eicos_ --csv --log-file $OUT/launches.csv $B > $OUT/launches.log 2>&1
python - <<PY
import csv, collections
rows = list(csv.reader(open("$OUT/launches.csv")))
hdr = None
agg = collections.defaultdict(lambda: collections.defaultdict(float))
cnt = collections.Counter()
for r in rows:
    if "Kernel Name" in r:
        hdr = r; continue
    if hdr is None or len(r) != len(hdr): continue
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
for k, m in sorted(agg.items(), key=lambda kv: -kv[1]["gpu__time_duration.sum"]):
    n = cnt[k]
    print("%-24s launches %4d total %9.2f ms avg %8.3f ms inst/launch %.3e dram GB/launch r %.2f w %.2f" % (k, n, m["gpu__time_duration.sum"], m["gpu__time_duration.sum"] / n, m["smsp__inst_executed.sum"] / n, m["dram__bytes_read.sum"] / n / 1e9, m["dram__bytes_write.sum"] / n / 1e9))
PY
for K in eicos_solve_kkt_pair eicos_solve_kkt; do
  echo "== ncu full: $K"
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:^$K\$ -s 2 -c 1 -f -o $OUT/prof_$K $B > $OUT/prof_$K.log 2>&1
done
ls -la $OUT
