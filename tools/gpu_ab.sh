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
BATCHES="${BATCHES:-65536}"
run g3 EICOS_SHALLOW_GROUPS=3
run g5 EICOS_SHALLOW_GROUPS=5
run g6 EICOS_SHALLOW_GROUPS=6
run g3again EICOS_SHALLOW_GROUPS=3
