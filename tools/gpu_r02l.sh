#!/bin/bash
# round 2, visit l (1 GPU): noise-texture path parity; full suite; final bench line; ncu launch list + full capture
OUT=gpurun_out/r02l; mkdir -p $OUT
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -8 | tee $OUT/pytest_gpu.txt
python tools/variant_time.py clouds1080 native 2>&1 | tee $OUT/variants.txt
echo "== bench"; timeout 900 python bench.py --steps 20 --warmup 3 2> $OUT/bench.err | tee $OUT/bench.json | cut -c1-400; tail -3 $OUT/bench.err
echo "== smoke"; python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee $OUT/smoke.txt
echo "== ncu launches"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu --no-extra > $OUT/bench_under_ncu.log 2>&1; tail -2 $OUT/bench_under_ncu.log | cut -c1-200
echo "== ncu full"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sbx_render -s 3 -c 1 -o $OUT/prof_clouds1080 \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-extra > $OUT/ncu_full.log 2>&1; tail -1 $OUT/ncu_full.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sbx_render -s 3 -c 1 -o $OUT/prof_clouds_tex1080 \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-extra --workload clouds_tex1080 > $OUT/ncu_full_tex.log 2>&1; tail -1 $OUT/ncu_full_tex.log
echo done
