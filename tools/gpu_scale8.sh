#!/bin/bash
# 8 GPUs of one node: the metric's configuration (65 536 instances IN TOTAL, strong scaling) and the weak form
TAG=${1:-r02c}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=index,name --format=csv > $OUT/gpus.txt 2>&1
N=${2:-8}
for S in strong weak; do
  echo "== $N GPUs, $S scaling"
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --scaling $S --steps 2 --warmup 3 --no-cpu-baseline 2>$OUT/scale_${S}_$N.err | grep '^{' | tee $OUT/scale_${S}_$N.json | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l)
    print({k: d[k] for k in ('value','ms_per_step','n_gpus','scaling')}, 'e2e', d['e2e']['value'] if d.get('e2e') else None, d['config']['total_batch'], d['per_rank'])
"
  tail -2 $OUT/scale_${S}_$N.err
done
echo "== multi-GPU handle test on this box"
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "multi_gpu" 2>&1 | tail -3 | tee $OUT/pytest_multi.txt
