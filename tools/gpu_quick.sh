#!/bin/bash
# quick GPU visit: parity tests for the CLOUDS/noise paths, then kernel variants back to back
# usage (under gpurun): bash tools/gpu_quick.sh <tag> "<pytest -k expr>" variant...
TAG=${1:-q}; K=${2:-clouds}; shift; shift
OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 900 python -m pytest tests -x -q -m gpu -k "$K" 2>&1 | tail -12 | tee $OUT/pytest.txt
bash tools/bench_variants.sh clouds1080 "$@" 2>&1 | tee $OUT/variants.txt
