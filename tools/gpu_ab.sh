#!/bin/bash
# A/B of program-form settings at the full batch (and the 8-GPU share): environment overrides, short bench lines
TAG=${1:-ab}
OUT=gpurun_out/$TAG
mkdir -p $OUT
run() {
  name=$1; shift
  for B in $BATCHES; do
    echo "== $name batch $B: $*"
    env "$@" timeout 600 python bench.py --batch $B --steps 1 --warmup 1 --no-e2e --no-cpu-baseline 2>$OUT/${name}_$B.err | tee $OUT/${name}_$B.json | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: print(l[:300]); continue
    r = d['roofline']
    print({k: round(d[k], 1) for k in ('value','ms_per_step')}, 'solve frac %.3f avg_ms %.2f' % (r['frac'], r['avg_launch_ms']), 'factor %.2f ms' % r['ldl_factor']['avg_launch_ms'], {k: round(v) for k, v in d['kernel_ms'].items()})
"
    tail -2 $OUT/${name}_$B.err
  done
}
BATCHES="${BATCHES:-65536 8192}"
echo "== pytest -m gpu (slice)"; timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "single_instance or batched_perturbed or starved or forms" 2>&1 | tail -5 | tee $OUT/pytest_gpu.txt
run base EICOS_NO_TRAFFIC_FLAGS=1 EICOS_L2_PREFETCH=0
run flags EICOS_L2_PREFETCH=0
run pf EICOS_NO_TRAFFIC_FLAGS=1
run both X=1
