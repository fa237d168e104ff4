#!/bin/bash
# round 2, closing check (8 GPUs): the default bench.py line at N = 8, exactly as the driver launches it
OUT=gpurun_out/r02s; mkdir -p $OUT
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 8 --steps 20 --warmup 3 2> $OUT/n8.err > $OUT/n8.json
python - <<'PY'
import json
d = json.load(open("gpurun_out/r02s/n8.json"))
print("n8 value %.1f ms %.4f e2e %.1f equal %s group %s kernels %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["frame"].get("equals_single_gpu_render"),
      (d.get("single_process_group") or {}).get("value"), " ".join("%.3f" % x for x in d["kernel_ms_per_rank"])))
for k, v in d["extra_workloads"].items(): print("  ", k, {a: (round(b, 1) if isinstance(b, float) else b) for a, b in v.items() if a in ("value", "e2e", "equals_single_gpu_render", "error", "skipped")})
PY
tail -2 $OUT/n8.err
