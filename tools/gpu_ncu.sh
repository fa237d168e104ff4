#!/bin/bash
# one ncu --set full capture of the render kernel of a bench workload: tools/gpu_ncu.sh <tag> [workload] [variant]
TAG=${1:-p}; W=${2:-clouds1080}; V=${3:-}
OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sbx_render -s 3 -c 1 -o $OUT/prof \
    python bench.py --steps 2 --warmup 3 --no-cpu --workload $W ${V:+--variant $V} > $OUT/ncu_full.log 2>&1
tail -2 $OUT/ncu_full.log
