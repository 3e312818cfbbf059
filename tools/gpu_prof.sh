#!/bin/bash
# ncu --set full (+source) of one kernel at a small batch. Usage: bash tools/gpu_prof.sh <tag> <kernel-regex> <skip> [batch]
TAG=$1; K=$2; SKIP=${3:-8}; B=${4:-2048}
OUT=gpurun_out/$TAG; mkdir -p $OUT
ncu --set full --clock-control none --import-source on -k regex:$K -s $SKIP -c 1 -f -o $OUT/prof_$K \
  python bench.py --batch $B --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --workers ${W:-1} > $OUT/prof_$K.log 2>&1
tail -3 $OUT/prof_$K.log
ls -la $OUT
