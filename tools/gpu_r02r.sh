#!/bin/bash
# round 2, last visit (1 GPU): whole GPU suite, ncu capture of the shipped native CLOUDS image (traffic.json), launch list
OUT=gpurun_out/r02r; mkdir -p $OUT
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -4 | tee $OUT/pytest_gpu.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sbx_render -s 3 -c 1 -o $OUT/prof_clouds1080 \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-extra > $OUT/ncu_full.log 2>&1; tail -1 $OUT/ncu_full.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu --no-extra > $OUT/bench_under_ncu.log 2>&1
sha256sum shaderbox_b200/images/APP_CLOUDS.native.cubin | tee $OUT/native_sha.txt
echo done
