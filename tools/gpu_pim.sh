#!/bin/bash
# GPU session for the per-instance-matrices path: parity tests, a regression check of the default
# bench (device arm only), the mpc02pim bench line, and an ncu launch list of the pim launch sequence.
# Usage (under gpurun): bash tools/gpu_pim.sh <tag>
TAG=${1:-pim}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== pytest -m gpu"; timeout 420 python -m pytest tests -x -q -m gpu 2>&1 | tail -12 | tee $OUT/pytest_gpu.txt
echo "== bench (default workload, device arm)"
timeout 200 python bench.py --steps 2 --warmup 2 --no-e2e --no-cpu-baseline 2>$OUT/bench_default.err | tee $OUT/bench_default.json | cut -c1-300
tail -3 $OUT/bench_default.err
echo "== bench mpc02pim"
timeout 300 python bench.py --workload mpc02pim --steps 2 --warmup 3 2>$OUT/bench_pim.err | tee $OUT/bench_pim.json | cut -c1-3000
tail -3 $OUT/bench_pim.err
echo "== ncu launch list (pim, batch 4096)"
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:eicos_ --csv --log-file $OUT/launches_pim.csv \
  python bench.py --workload mpc02pim --batch 4096 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > $OUT/launches_pim.log 2>&1
ls -la $OUT
