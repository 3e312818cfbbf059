#!/bin/bash
# Quick GPU iteration: parity tests + short device-resident bench lines (no CPU baseline, no e2e).
# Usage (under gpurun): bash tools/gpu_quick.sh <tag> [bench args...]
TAG=${1:-q}; shift
OUT=gpurun_out/$TAG
mkdir -p $OUT
[ -n "$SKIP_TESTS" ] || { echo "== pytest -m gpu"; timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -8 | tee $OUT/pytest_gpu.txt; }
for W in ${WORKERS:-1 2}; do
  echo "== bench workers=$W"
  timeout 600 python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --workers $W "$@" 2>$OUT/bench_w$W.err | tee $OUT/bench_w$W.json | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print({k: d[k] for k in ('value', 'ms_per_step', 'kernel_ms', 'kkt_phase_share')}, d['roofline']['frac'], d['roofline']['ldl_factor']['frac'], d['config'].get('exit_flags'))"
  tail -3 $OUT/bench_w$W.err
done
