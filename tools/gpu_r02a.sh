#!/bin/bash
# round 2, visit a (1 GPU): parity suite, part timings for the N-GPU cuts, sanitizer, ncu of ATMOSPHERE / EGG, bench line
OUT=gpurun_out/r02a; mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/gpu.txt 2>&1; nproc > $OUT/nproc.txt
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -25 | tee $OUT/pytest_gpu.txt
echo "== part times"; for n in 8 4 2; do timeout 600 python tools/part_time.py clouds1080 $n --out $OUT/parts_clouds1080_n$n.json 2>&1 | tee $OUT/parts_clouds1080_n$n.txt; done
timeout 300 python tools/part_time.py planet2160 8 --out $OUT/parts_planet2160_n8.json 2>&1 | tee $OUT/parts_planet2160_n8.txt
echo "== bench"; timeout 600 python bench.py --steps 20 --warmup 3 2> $OUT/bench.err | tee $OUT/bench.json; tail -3 $OUT/bench.err
echo "== sanitizer"; bash tools/gpu_sanitize.sh r02a/sanitizer
echo "== ncu atmosphere / egg"
for wl in atmosphere1080 egg256; do
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sbx_render -s 3 -c 1 -o $OUT/prof_$wl \
    python bench.py --steps 2 --warmup 3 --no-cpu --workload $wl > $OUT/ncu_$wl.log 2>&1; tail -1 $OUT/ncu_$wl.log
done
echo done
