#!/bin/bash
# quick GPU check of a new build: parity tests (bounded), then short bench lines at the full batch and at the 8-GPU share
TAG=${1:-try}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 | tee $OUT/pytest_gpu.txt
for B in 65536 8192; do
  echo "== bench batch $B"
  timeout 600 python bench.py --batch $B --steps 2 --warmup 3 --no-e2e --no-cpu-baseline 2>$OUT/bench_$B.err | tee $OUT/bench_$B.json | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: print(l[:300]); continue
    r = d['roofline']
    print({k: d[k] for k in ('value','ms_per_step')}, 'solve frac %.3f avg_ms %.2f' % (r['frac'], r['avg_launch_ms']), 'factor frac %.3f' % r['ldl_factor']['frac'], d['kernel_ms'], d['kkt_phase_share'], d['config']['exit_flags'], d['config']['iterations_mean'])
"
  tail -3 $OUT/bench_$B.err
done
