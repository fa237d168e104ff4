"""Kernel time of a bench workload per kernel image variant (CUDA events, L2 flushed before each launch, median of 15):
    python tools/variant_time.py <workload> variant..."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import shaderbox_b200 as sbx
from bench import WORKLOADS
from shaderbox_b200.abi import default_params

wl = sys.argv[1]
app, w, h, t, ov = WORKLOADS[wl]
p = default_params(w, h, t, **ov)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
frame = torch.empty((h, w, 4), dtype=torch.float32, device="cuda")
s = torch.cuda.current_stream()
for v in sys.argv[2:]:
    r = sbx.Renderer(app, variant=None if v == "default" else v)
    for _ in range(3):
        r.render_into(p, frame.data_ptr(), stream=s.cuda_stream)
    ms = []
    for _ in range(15):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s); r.render_into(p, frame.data_ptr(), stream=s.cuda_stream); e1.record(s)
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    m = sorted(ms)[len(ms) // 2]
    tm = r.timing()
    print("%-16s %-12s %.4f ms  %9.1f Mpix/s  (regs %d, ctas/sm %d, grid %d)" % (wl, v, m, w * h / m * 1e-3, tm["regs_per_thread"], tm["blocks_per_sm"], tm["grid_blocks"]), flush=True)
    r.close()
