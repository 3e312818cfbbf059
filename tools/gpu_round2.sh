#!/bin/bash
# Round-2 GPU session A: the whole gpu test tier, workers A/B, the +-5 % recipe of SURVEY.md 8d
TAG=${1:-r02a}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/gpu.txt 2>&1
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -25 | tee $OUT/pytest_gpu.txt
echo "== smoke"; python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee $OUT/smoke.txt
short() {
  python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: print(l[:300]); continue
    r = d['roofline']
    print({k: round(d[k], 1) for k in ('value','ms_per_step')}, 'solve frac %.3f avg_ms %.2f' % (r['frac'], r['avg_launch_ms']), 'factor %.2f ms' % r['ldl_factor']['avg_launch_ms'], {k: round(v) for k, v in d['kernel_ms'].items()}, d['config']['exit_flags'], (d.get('cpu_baseline') or {}).get('exit_flags'))
"
}
for W in 4 8; do for B in 65536 8192; do
  echo "== workers $W batch $B"
  timeout 600 python bench.py --batch $B --workers $W --steps 1 --warmup 1 --no-e2e --no-cpu-baseline 2>$OUT/w${W}_$B.err | tee $OUT/w${W}_$B.json | short
done; done
echo "== mpc02pct5 (h, b +-5 %: the survey's recipe)"
timeout 900 python bench.py --workload mpc02pct5 --steps 2 --warmup 3 2>$OUT/bench_pct5.err | tee $OUT/bench_pct5.json | short
tail -3 $OUT/bench_pct5.err
