#!/bin/bash
# Last GPU session of the round: the C++-facade and reference-fixture tests on the CUDA library, smoke(),
# and one ncu --set full capture of eicos_equilibrate (per-instance-matrices path, batch 4096).
TAG=${1:-fin}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== pytest -m gpu (facade + reference fixtures)"
timeout 240 python -m pytest tests/test_cpp_facade.py tests/test_reference_tester.py -x -q -m gpu 2>&1 | tail -8 | tee $OUT/pytest_gpu_facade.txt
echo "== smoke"; timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee $OUT/smoke.txt
echo "== ncu full: eicos_equilibrate"
timeout 150 ncu --set full --clock-control none --import-source on -k regex:eicos_equilibrate -c 1 -f -o $OUT/prof_equilibrate \
  python bench.py --workload mpc02pim --batch 4096 --steps 1 --warmup 0 --no-e2e --no-cpu-baseline > $OUT/prof_equilibrate.log 2>&1
tail -2 $OUT/prof_equilibrate.log
ls -la $OUT
