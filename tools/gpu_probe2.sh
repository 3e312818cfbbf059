#!/bin/bash
# latency floor vs throughput: one tile, a few tiles, full batch; several worker counts
run() {
  python bench.py --batch $1 --workers $2 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline 2>&1 | grep '^{' | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read())
k = d['kernel_ms']; c = d['config']
print('batch', c['batch_per_gpu'], 'workers', c['workers_per_tile'], 'ms_step %.0f' % d['ms_per_step'], 'solve %.0f factor %.0f other %.0f' % (k['solve_kkt'], k['ldl_factor'], k['other']), 'iters max', c['iterations_max'], 'launches', d['gpu_launches'], 'solves/s %.0f' % d['value'])
"
}
if [ "$1" == "ncu1" ]; then
  mkdir -p gpurun_out/tile1
  SMALL="python bench.py --batch 64 --workers 1 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline"
  ncu --set full --clock-control none --import-source on -k regex:"eicos_solve_kkt|eicos_iter_tail|eicos_ldl_factor|eicos_iter_head" -s 12 -c 8 -f -o gpurun_out/tile1/prof_tile1 $SMALL > gpurun_out/tile1/log.txt 2>&1
  tail -3 gpurun_out/tile1/log.txt
  exit 0
fi
run 64 2; run 65536 2
