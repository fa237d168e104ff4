#!/bin/bash
# round 2, visit p (1 GPU): the unchanged app_clouds.h with USE_NOISE_TEX as a plugin vs the hand-written TEX kernel
OUT=gpurun_out/r02p; mkdir -p $OUT
timeout 600 python -m pytest tests/test_noise_tex.py tests/test_sass_cpu.py -x -q -m "gpu or not gpu" 2>&1 | tail -4 | tee $OUT/pytest.txt
python - <<'PY' 2>&1 | tee $OUT/tex_variants.txt
import numpy as np, torch, shaderbox_b200 as sbx
from shaderbox_b200.abi import default_params
p = default_params(1920, 1080, 1.5, cld_march_steps=128)
frame = torch.empty((1080, 1920, 4), dtype=torch.float32, device="cuda"); flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
s = torch.cuda.current_stream()
for v in ("native", "plugin"):
    r = sbx.Renderer("APP_CLOUDS_TEX", variant=v)
    vol = r.bake_noise_volume(128); r.set_noise_volumes(vol, np.ascontiguousarray(vol[::-1]))
    ms = []
    for k in range(12):
        flush.zero_(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s); r.render_into(p, frame.data_ptr(), stream=s.cuda_stream); e1.record(s); torch.cuda.synchronize(); ms.append(e0.elapsed_time(e1))
    m = sorted(ms[2:])[5]; tm = r.timing()
    print("clouds_tex1080 %-7s %.4f ms %8.1f Mpix/s (regs %d, ctas/sm %d)" % (v, m, 1920 * 1080 / m * 1e-3, tm["regs_per_thread"], tm["blocks_per_sm"]))
    r.close()
PY
echo done
