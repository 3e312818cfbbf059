#!/bin/bash
# round-2 baseline: the metric's per-GPU share at 8 GPUs (8192 instances) and the full batch, round-1 build
TAG=${1:-r02a}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/gpu.txt 2>&1
for B in 8192 65536; do
  echo "== bench batch $B"
  timeout 600 python bench.py --batch $B --steps 2 --warmup 3 --no-e2e --no-cpu-baseline 2>$OUT/bench_$B.err | tee $OUT/bench_$B.json | cut -c1-1200
  tail -3 $OUT/bench_$B.err
done
