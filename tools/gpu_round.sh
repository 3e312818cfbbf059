#!/bin/bash
# One GPU-box session: parity tests, the bench line, and the ncu evidence for profiles/.
# Usage (from the repo root, under gpurun):  bash tools/gpu_round.sh [tag]
TAG=${1:-r01}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/gpu.txt 2>&1
echo "== pytest -m gpu"; python -m pytest tests -x -q -m gpu 2>&1 | tail -15 | tee $OUT/pytest_gpu.txt
echo "== smoke"; python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee $OUT/smoke.txt
echo "== bench"; python bench.py 2>$OUT/bench.err | tee $OUT/bench.json | cut -c1-1500; tail -5 $OUT/bench.err
echo "== bench reference"; python bench.py --impl reference --steps 2 --warmup 1 2>&1 | tail -1 | tee $OUT/bench_reference.json | cut -c1-400
echo "== bench mpc02pim (per-instance matrices)"; python bench.py --workload mpc02pim 2>$OUT/bench_pim.err | tee $OUT/bench_pim.json | cut -c1-600; tail -3 $OUT/bench_pim.err
SMALL="python bench.py --batch 4096 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline"
echo "== ncu launch list"
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:eicos_ --csv --log-file $OUT/launches.csv $SMALL > $OUT/launches.log 2>&1
echo "== ncu full: solve_kkt"
ncu --set full --clock-control none --import-source on -k regex:eicos_solve_kkt -s 8 -c 2 -f -o $OUT/prof_solve_kkt $SMALL > $OUT/prof_solve.log 2>&1
echo "== ncu full: ldl_factor"
ncu --set full --clock-control none --import-source on -k regex:eicos_ldl_factor -s 3 -c 2 -f -o $OUT/prof_ldl_factor $SMALL > $OUT/prof_factor.log 2>&1
echo "== ncu DRAM traffic per launch at the bench batch (one pass, two metrics)"
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k "regex:eicos_(solve_kkt|ldl_factor)" --csv --log-file $OUT/traffic.csv \
  python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > $OUT/traffic.log 2>&1
python tools/ncu_traffic.py $OUT/traffic.csv $OUT/ncu_traffic.json
ls -la $OUT
