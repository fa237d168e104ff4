#!/bin/bash
# round 2: the metric's scene on the larger frames north_star names (3840x2160, 7680x4320), one GPU: parity rows + rate
OUT=gpurun_out/r02t; mkdir -p $OUT
timeout 240 python -m pytest tests/test_gpu_parity.py -q -x -k "APP_CLOUDS_3840 or APP_CLOUDS_7680" 2>&1 | tail -3 | tee $OUT/pytest.txt
for wl in clouds2160 clouds4320; do
  timeout 120 python bench.py --workload $wl --steps 10 --warmup 3 --cpu-seconds 4 2> $OUT/$wl.err > $OUT/$wl.json
  python - $OUT/$wl.json <<'PY'
import json, sys
d = json.load(open(sys.argv[1]))
print(d["config"]["workload"], "value %.1f ms %.3f e2e %.1f hash %s cpu %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["frame"].get("frame_hash"), (d.get("cpu_baseline") or {}).get("value")))
PY
  tail -1 $OUT/$wl.err
done
