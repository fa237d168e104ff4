#!/bin/bash
# round 2, visit h (8 GPUs): multi-GPU tests at world 2/4/8, strong-scaling bench lines, single-process group, host ceiling
OUT=gpurun_out/${1:-r02h}; mkdir -p $OUT
nvidia-smi -L | tee $OUT/gpus.txt
timeout 900 python -m pytest tests/test_multi_gpu.py "tests/test_abi_cpu.py::test_c_host_renders_one_frame_over_several_parts" -x -q -m gpu 2>&1 | tail -8 | tee $OUT/pytest.txt
run() { # name n extra-args
  local name=$1 n=$2; shift 2
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 300)) \
      bench.py --gpus $n --steps 20 --warmup 3 "$@" 2> $OUT/$name.err > $OUT/$name.json
  python - <<PY
import json
try:
    d = json.load(open("$OUT/$name.json"))
    g = d.get("single_process_group") or {}
    print("$name: value %.1f  ms/step %.4f  e2e %.1f  kernel ms/rank %s  equal %s  group %s/%s" % (d["value"], d["ms_per_step"], d["e2e"]["value"],
          " ".join("%.3f" % x for x in d["kernel_ms_per_rank"]), d["frame"].get("equals_single_gpu_render"), g.get("value"), g.get("e2e")))
    for k, v in (d.get("extra_workloads") or {}).items():
        print("   %s: value %.1f e2e %.1f ms %.4f equal %s" % (k, v["value"], v["e2e"], v["ms_per_step"], v.get("equals_single_gpu_render")))
except Exception as e:
    print("$name: FAILED", e); print(open("$OUT/$name.err").read()[-1500:])
PY
}
run n8 8
run n8_nccl 8 --completion nccl --no-extra --no-group
run n4 4
run n2 2 --no-extra
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu --no-extra 2> $OUT/n1.err > $OUT/n1.json; python -c "
import json; d=json.load(open('$OUT/n1.json')); print('n1: value %.1f ms/step %.4f e2e %.1f pageable %.1f' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['pageable_value']))"
for g in 1 2 4 8; do shaderbox_b200/sbx_cli render APP_CLOUDS 1920 1080 1.5 - --steps 128 --frames 10 --gpus $g; done 2>&1 | tee $OUT/cli.txt
echo done
