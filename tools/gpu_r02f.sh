#!/bin/bash
# round 2, visit f (1 GPU): native ATMOSPHERE / EGG kernels: parity, time vs their plugin images, ncu
OUT=gpurun_out/r02f; mkdir -p $OUT
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -8 | tee $OUT/pytest_gpu.txt
python tools/variant_time.py atmosphere1080 plugin native 2>&1 | tee $OUT/variants.txt
python tools/variant_time.py egg256 plugin native native_c6 2>&1 | tee -a $OUT/variants.txt
python tools/variant_time.py clouds1080 native coop coop2 2>&1 | tee -a $OUT/variants.txt
for wl in atmosphere1080 egg256; do
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sbx_render -s 3 -c 1 -o $OUT/prof_${wl}_native \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-extra --workload $wl --variant native > $OUT/ncu_$wl.log 2>&1; tail -1 $OUT/ncu_$wl.log
done
echo done
