#!/bin/bash
# round 2, visit q (2 GPUs): the final bench.py at N = 2 and N = 1, the multi-GPU tests, smoke
OUT=gpurun_out/r02q; mkdir -p $OUT
timeout 600 python -m pytest tests/test_multi_gpu.py -x -q -m gpu 2>&1 | tail -3 | tee $OUT/pytest_multi.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 3 2> $OUT/n2.err > $OUT/n2.json; python - <<'PY'
import json
d = json.load(open("gpurun_out/r02q/n2.json"))
print("n2 value %.1f ms %.4f e2e %.1f equal %s group %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["frame"].get("equals_single_gpu_render"), (d.get("single_process_group") or {}).get("value")))
for k, v in d["extra_workloads"].items(): print("  ", k, {a: (round(b, 1) if isinstance(b, float) else b) for a, b in v.items() if a in ("value", "e2e", "equals_single_gpu_render", "error", "skipped")})
PY
timeout 600 python bench.py 2> $OUT/n1.err > $OUT/n1.json; python - <<'PY'
import json
d = json.load(open("gpurun_out/r02q/n1.json"))
print("n1 value %.1f ms %.4f e2e %.1f pageable %.1f issue stale %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["pageable_value"], d["issue_roofline"]["stale"]))
for k, v in d["extra_workloads"].items(): print("  ", k, {a: (round(b, 1) if isinstance(b, float) else b) for a, b in v.items() if a in ("value", "e2e", "error", "skipped")})
PY
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | cut -c1-300
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
echo done
