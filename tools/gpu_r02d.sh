#!/bin/bash
# round 2, visit d (1 GPU): whole GPU suite; octave-1/2 merge and branch-free merge experiments; bench line with extras
OUT=gpurun_out/r02d; mkdir -p $OUT
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -8 | tee $OUT/pytest_gpu.txt
V=native,native_m12,coop,coop_m12,coop_brm
timeout 600 python tools/part_time.py clouds1080 8 --variants $V --splits rows4 --out $OUT/parts_n8.json 2>&1 | tee $OUT/parts_n8.txt
timeout 600 python tools/part_time.py clouds1080 1 --variants $V --splits rows4 --out $OUT/parts_n1.json 2>&1 | tee $OUT/parts_n1.txt
echo "== bench"; timeout 900 python bench.py --steps 20 --warmup 3 2> $OUT/bench.err | tee $OUT/bench.json; tail -5 $OUT/bench.err
echo "== bench reference arm"; timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2> $OUT/bench_ref.err | tee $OUT/bench_ref.json
echo done
