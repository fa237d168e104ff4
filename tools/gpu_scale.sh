#!/bin/bash
# strong-scaling lines on an N-GPU box: tools/gpu_scale.sh <tag> "<N list>" workload...
TAG=${1:-s}; NS=${2:-"2 4 8"}; shift; shift
OUT=gpurun_out/$TAG; mkdir -p $OUT
for w in "$@"; do for n in $NS; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 20 --warmup 3 --workload $w 2>$OUT/${w}_n$n.err | tee $OUT/${w}_n$n.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$w', d['n_gpus'], 'value', round(d['value'],1), 'ms', round(d['ms_per_step'],4), 'kernel(max rank)', round(d['roofline']['kernel_ms'],4), 'e2e', round(d['e2e']['value'],1))"
done; done
