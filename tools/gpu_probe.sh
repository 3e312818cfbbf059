#!/bin/bash
# Quick performance probes: parity tests, workers sweep at full batch, optional loaded ncu capture.
TAG=${1:-probe}
WORKERS=${2:-"1 2 4"}
NCU=${3:-no}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== pytest -m gpu"; python -m pytest tests -x -q -m gpu 2>&1 | tail -4
for w in $WORKERS; do
  echo "== workers=$w"
  python bench.py --workers $w --steps 1 --warmup 1 --no-e2e --no-cpu-baseline 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print({k: d[k] for k in ('value','ms_per_step','kernel_ms')}, 'solve frac', d['roofline']['frac'], 'factor frac', d['roofline']['ldl_factor']['frac'])
"
done
if [ "$NCU" != "no" ]; then
echo "== ncu full: solve_kkt at batch 16384"
ncu --set full --clock-control none --import-source on -k regex:eicos_solve_kkt -s 8 -c 1 -f -o $OUT/prof_solve_kkt_16k python bench.py --batch 16384 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > $OUT/prof_solve.log 2>&1
tail -2 $OUT/prof_solve.log
fi
