#!/bin/bash
# round 2, visit b (1 GPU): per-warp timelines of one rank's share, ncu of coop vs native on the share and the frame
OUT=gpurun_out/r02b; mkdir -p $OUT
for v in native_trace coop_trace; do
  python tools/trace_timeline.py clouds1080 8 3 rows4 $v --out $OUT/trace_n8_rows4_$v.npz 2>&1 | tee -a $OUT/trace.txt
  python tools/trace_timeline.py clouds1080 1 0 rows4 $v --out $OUT/trace_n1_$v.npz 2>&1 | tee -a $OUT/trace.txt
done
python tools/trace_timeline.py clouds1080 8 3 tiles hybrid_trace 100 --out $OUT/trace_n8_tiles_hybrid100.npz 2>&1 | tee -a $OUT/trace.txt
for cfg in "1 0 rows4 coop" "8 3 rows4 coop" "8 3 rows4 native"; do
  set -- $cfg
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:sbx_render -s 3 -c 1 -o $OUT/prof_n$1_$4 \
      python tools/ncu_part.py clouds1080 $1 $2 $3 $4 > $OUT/ncu_n$1_$4.log 2>&1; tail -1 $OUT/ncu_n$1_$4.log
done
echo done
