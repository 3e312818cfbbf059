#!/bin/bash
# latency floor vs throughput: one tile, a few tiles, full batch; several worker counts
for cfg in "64 1" "64 4" "64 8" "4096 1" "4096 4" "4096 8" "65536 1" "65536 2"; do
  set -- $cfg
  python bench.py --batch $1 --workers $2 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read())
k = d['kernel_ms']; c = d['config']
print('batch', c['batch_per_gpu'], 'workers', c['workers_per_tile'], 'ms_step %.0f' % d['ms_per_step'], 'solve %.0f factor %.0f other %.0f' % (k['solve_kkt'], k['ldl_factor'], k['other']), 'iters max', c['iterations_max'], 'launches', d['gpu_launches'], 'solves/s %.0f' % d['value'])
"
done
