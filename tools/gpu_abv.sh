#!/bin/bash
# A/B of kernel builds (eicos_b200/variants/*.so, built with different -DEICOS_K_* forms): short bench lines
TAG=${1:-abv}; shift
OUT=gpurun_out/$TAG
mkdir -p $OUT
BATCHES="${BATCHES:-65536 8192}"
for spec in "$@"; do
  name=${spec%%:*}; envs=${spec#*:}; [ "$envs" = "$spec" ] && envs="X=1"
  cp eicos_b200/variants/$name.so eicos_b200/libeicos_b200.so
  for B in $BATCHES; do
    echo "== $name batch $B: $envs"
    env $envs timeout 600 python bench.py --batch $B --steps 1 --warmup 1 --no-e2e --no-cpu-baseline 2>$OUT/${name}_$B.err | tee $OUT/${name}_$B.json | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: print(l[:300]); continue
    r = d['roofline']
    print({k: round(d[k], 1) for k in ('value','ms_per_step')}, 'solve avg_ms %.2f' % (r['avg_launch_ms']), 'factor %.2f ms' % r['ldl_factor']['avg_launch_ms'], {k: round(v) for k, v in d['kernel_ms'].items()})
"
    tail -2 $OUT/${name}_$B.err
  done
done
