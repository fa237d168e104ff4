#!/bin/bash
# round 2 (8 GPUs): the metric's scene on 3840x2160 and 7680x4320 frames, row stripes over 8 GPUs
OUT=gpurun_out/r02u; mkdir -p $OUT
for wl in clouds4320 clouds2160; do
  timeout 100 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29546 bench.py --gpus 8 --workload $wl --steps 10 --warmup 3 --no-group --no-cpu 2> $OUT/$wl.err > $OUT/$wl.json
  python - $OUT/$wl.json <<'PY'
import json, sys
d = json.load(open(sys.argv[1]))
print(d["config"]["workload"], "n8 value %.1f ms %.4f e2e %.1f equal %s hash %s kernels %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["frame"].get("equals_single_gpu_render"),
      d["frame"].get("frame_hash"), " ".join("%.3f" % x for x in d["kernel_ms_per_rank"])))
PY
  tail -1 $OUT/$wl.err
done
