#!/bin/bash
# round 2, visit i (1 GPU): whole suite after the table prefetch; N=8/4/1 part times; launch floor
OUT=gpurun_out/r02i; mkdir -p $OUT
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -5 | tee $OUT/pytest_gpu.txt
timeout 600 python tools/part_time.py clouds1080 8 --variants native,coop,coop2 --splits rows4 --out $OUT/parts_n8.json 2>&1 | tee $OUT/parts_n8.txt
timeout 600 python tools/part_time.py clouds1080 1 --variants native,coop --splits rows4 --out $OUT/parts_n1.json 2>&1 | tee $OUT/parts_n1.txt
timeout 600 python tools/part_time.py planet2160 8 --variants native --splits rows4 2>&1 | tee $OUT/parts_planet_n8.txt
echo done
