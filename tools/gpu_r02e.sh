#!/bin/bash
# round 2, visit e (2 GPUs): multi-GPU tests (IPC frame + completion flags, single-process group), bench at N=2
OUT=gpurun_out/r02e; mkdir -p $OUT
nvidia-smi -L | tee $OUT/gpus.txt
timeout 900 python -m pytest tests/test_multi_gpu.py tests/test_parts_gpu.py -x -q -m gpu 2>&1 | tail -15 | tee $OUT/pytest.txt
for sig in flags nccl; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 --signal $sig --no-extra 2> $OUT/bench_n2_$sig.err | tee $OUT/bench_n2_$sig.json | cut -c1-1200
tail -3 $OUT/bench_n2_$sig.err
done
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 3 2> $OUT/bench_n2_full.err | tee $OUT/bench_n2_full.json | cut -c1-600
tail -3 $OUT/bench_n2_full.err
shaderbox_b200/sbx_cli render APP_CLOUDS 1920 1080 1.5 - --steps 128 --frames 5 --gpus 2 | tee $OUT/cli_gpus2.json
shaderbox_b200/sbx_cli render APP_CLOUDS 1920 1080 1.5 - --steps 128 --frames 5 | tee $OUT/cli_gpus1.json
echo done
