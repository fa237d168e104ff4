#!/bin/bash
# bench every BASELINE.json workload on one GPU: tools/gpu_workloads.sh <tag>
TAG=${1:-w}; OUT=gpurun_out/$TAG; mkdir -p $OUT
for w in clouds1080 clouds1080_default100 atmosphere1080 planet2160 raytracer4320 egg256; do
  timeout 600 python bench.py --steps 10 --workload $w --cpu-seconds 8 2>$OUT/$w.err | tee $OUT/$w.json | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('%-22s %9.1f Mpix/s %8.3f ms  e2e %9.1f  cpu %8.4f (%d cores)  regs %3d ctas/sm %d hbm_frac %.4f' % ('$w', d['value'], d['ms_per_step'], d['e2e']['value'], d['cpu_baseline']['value'], d['cpu_baseline']['cores'], d['config']['regs'], d['config']['ctas_per_sm'], d['roofline']['frac']))"
done
