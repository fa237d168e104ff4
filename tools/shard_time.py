"""Time one rank's share of a frame on ONE GPU (what a rank of an N-GPU run executes):
    python tools/shard_time.py <workload> <n_parts> variant..."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import shaderbox_b200 as sbx
from bench import WORKLOADS
from shaderbox_b200.abi import default_params, shard_rows

wl, parts = sys.argv[1], int(sys.argv[2])
app, w, h, t, ov = WORKLOADS[wl]
p = default_params(w, h, t, **ov)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for v in sys.argv[3:]:
    r = sbx.Renderer(app, variant=v)
    for stripe in (4, 1):
        ms = []
        for part in (0, parts // 2):
            rows = len(shard_rows(stripe, parts, part, h))
            buf = torch.empty((rows, w, 4), dtype=torch.float32, device="cuda")
            s = torch.cuda.current_stream()
            for _ in range(3):
                r.render_into(p, buf.data_ptr(), shard=(stripe, parts, part), stream=s.cuda_stream)
            best = []
            for _ in range(10):
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(s); r.render_into(p, buf.data_ptr(), shard=(stripe, parts, part), stream=s.cuda_stream); e1.record(s)
                torch.cuda.synchronize(); best.append(e0.elapsed_time(e1))
            ms.append(sorted(best)[len(best) // 2])
        tm = r.timing()
        print("%-10s parts %d stripe %d: part0 %.4f ms  part%d %.4f ms   (grid %d, ctas/sm %d)" % (v, parts, stripe, ms[0], parts // 2, ms[1], tm["grid_blocks"], tm["blocks_per_sm"]))
    r.close()
