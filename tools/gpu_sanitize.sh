#!/bin/bash
# compute-sanitizer (memcheck + racecheck) over smoke-sized launches of every hand-written kernel image, through the
# C++ launcher (no Python, no torch kernels in the way).  usage (under gpurun): bash tools/gpu_sanitize.sh <tag>
TAG=${1:-san}; OUT=gpurun_out/$TAG; mkdir -p $OUT
CLI=shaderbox_b200/sbx_cli
run() {   # name, tool, args...
    local name=$1 tool=$2; shift 2
    timeout 600 compute-sanitizer --tool $tool --error-exitcode 9 --print-limit 5 $CLI "$@" > $OUT/${name}_${tool}.log 2>&1
    echo "$name $tool rc=$? $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' $OUT/${name}_${tool}.log | tail -1)"
}
for tool in memcheck racecheck; do
    run clouds_native $tool render APP_CLOUDS 256 144 1.5 - --variant native --steps 32
    run clouds_coop4  $tool render APP_CLOUDS 256 144 1.5 - --variant coop --steps 32
    run clouds_coop2  $tool render APP_CLOUDS 256 144 1.5 - --variant coop2 --steps 32
    run clouds_hybrid $tool render APP_CLOUDS 331 203 1.5 - --steps 32
    run clouds_parts3 $tool render APP_CLOUDS 331 203 1.5 - --steps 32 --gpus 3
    run planet_native $tool render APP_PLANET 192 108 2.0 -
    run raytracer_native $tool render APP_RAYTRACER 320 180 1.0 -
    run atmosphere    $tool render APP_ATMOSPHERE 192 108 1.0 -
    run egg           $tool render APP_EGG 128 128 1.0 -
done 2>&1 | tee $OUT/summary.txt
