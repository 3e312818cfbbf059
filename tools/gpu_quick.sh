#!/bin/bash
# quick GPU check of a new build: a bounded slice of the parity tests, then short bench lines at the full batch and at the 8-GPU share
TAG=${1:-quick}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== pytest -m gpu (slice)"; timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "single_instance or batched_perturbed or starved or lanes or forms" 2>&1 | tail -5 | tee $OUT/pytest_gpu.txt
for B in ${2:-65536 8192}; do
  echo "== bench batch $B"
  timeout 600 python bench.py --batch $B --steps 2 --warmup 3 --no-e2e --no-cpu-baseline 2>$OUT/bench_$B.err | tee $OUT/bench_$B.json | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: print(l[:300]); continue
    r = d['roofline']
    print({k: d[k] for k in ('value','ms_per_step')}, 'solve frac %.3f avg_ms %.2f' % (r['frac'], r['avg_launch_ms']), 'factor frac %.3f avg_ms %.2f' % (r['ldl_factor']['frac'], r['ldl_factor']['avg_launch_ms']), d['kernel_ms'], d['kkt_phase_share'], d['config']['exit_flags'], d['config']['iterations_mean'])
"
  tail -3 $OUT/bench_$B.err
done
