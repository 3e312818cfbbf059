#!/bin/bash
# Round-2 GPU session B: the default bench line (+ reference arm), and the ncu evidence AT THE BENCH BATCH:
# launch list with instruction counts and DRAM bytes, --set full of the four hot kernels
TAG=${1:-r02b}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/gpu.txt 2>&1
echo "== bench (default)"; python bench.py 2>$OUT/bench.err | tee $OUT/bench.json | cut -c1-1800; tail -3 $OUT/bench.err
echo "== bench reference"; python bench.py --impl reference --steps 2 --warmup 1 2>&1 | tail -1 | tee $OUT/bench_reference.json | cut -c1-400
B="python bench.py --steps 1 --warmup 0 --no-e2e --no-cpu-baseline"
echo "== ncu launch list at batch 65536"
timeout 1200 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:eicos_ --csv --log-file $OUT/launches.csv $B > $OUT/launches.log 2>&1
python tools/ncu_launches.py $OUT/launches.csv | tee $OUT/launches_summary.txt
for K in eicos_solve_kkt eicos_ldl_factor eicos_residuals eicos_iter_head; do
  echo "== ncu full: $K"
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:^$K\$ -s 3 -c 1 -f -o $OUT/prof_$K $B > $OUT/prof_$K.log 2>&1
done
ls -la $OUT
