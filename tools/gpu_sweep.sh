#!/bin/bash
# bench at several batch sizes (device-resident only)
TAG=${1:-s}; shift
OUT=gpurun_out/$TAG; mkdir -p $OUT
for B in ${BATCHES:-4096 16384 65536}; do
  echo "== batch $B"
  timeout 600 python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --workers ${W:-1} --batch $B "$@" 2>$OUT/b$B.err | tee $OUT/b$B.json | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print({k: d[k] for k in ('value', 'ms_per_step', 'kernel_ms', 'kkt_phase_share')}, d['roofline']['frac'], d['roofline']['ldl_factor']['frac'])"
  tail -3 $OUT/b$B.err
done
