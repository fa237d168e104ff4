#!/bin/bash
# round 2, visit n (1 GPU): ncu captures of the two apps that had none (VINYL, SDF_AO); per-workload kernel times of every image
OUT=gpurun_out/r02n; mkdir -p $OUT
for wl in vinyl1080 sdf_ao1080; do
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sbx_render -s 3 -c 1 -o $OUT/prof_$wl \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-extra --workload $wl > $OUT/ncu_$wl.log 2>&1; tail -1 $OUT/ncu_$wl.log
done
for wl in egg256 vinyl1080 sdf_ao1080 atmosphere1080 planet2160 raytracer4320 clouds1080 clouds1080_default100; do python tools/variant_time.py $wl default; done 2>&1 | tee $OUT/workloads.txt
python tools/variant_time.py clouds1080 plugin 2>&1 | tee -a $OUT/workloads.txt
python tools/variant_time.py planet2160 plugin 2>&1 | tee -a $OUT/workloads.txt
echo done
