#!/bin/bash
# round 2, visit c (1 GPU): cooperative-march variants on one rank's share (CTA size, lanes per pixel, trimmed loop)
OUT=gpurun_out/r02c; mkdir -p $OUT
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "coop or hybrid or native_clouds or golden" 2>&1 | tail -5 | tee $OUT/pytest.txt
timeout 300 python -m pytest tests/test_parts_gpu.py -x -q -m gpu 2>&1 | tail -5 | tee -a $OUT/pytest.txt
V=native,coop,coop_w2,coop_w1,coop8,coop8_w2,coop2,native_w2
timeout 600 python tools/part_time.py clouds1080 8 --variants $V --splits rows4,rows1 --out $OUT/parts_n8.json 2>&1 | tee $OUT/parts_n8.txt
timeout 600 python tools/part_time.py clouds1080 4 --variants $V --splits rows4,rows1 --out $OUT/parts_n4.json 2>&1 | tee $OUT/parts_n4.txt
timeout 600 python tools/part_time.py clouds1080 1 --variants $V --splits rows4 --out $OUT/parts_n1.json 2>&1 | tee $OUT/parts_n1.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sbx_render -s 3 -c 1 -o $OUT/prof_n1_coop \
      python tools/ncu_part.py clouds1080 1 0 rows4 coop > $OUT/ncu_n1_coop.log 2>&1; tail -1 $OUT/ncu_n1_coop.log
echo done
