"""Render one part of a workload a few times (for `ncu -k regex:sbx_render -s 3 -c 1 ... python tools/ncu_part.py ...`):
    python tools/ncu_part.py <workload> <n_parts> <part> <rows4|tiles> [variant] [tail_waves_x100]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import shaderbox_b200 as sbx
from bench import WORKLOADS
from shaderbox_b200.abi import default_params

wl, parts, part, split = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), sys.argv[4]
variant = sys.argv[5] if len(sys.argv) > 5 and sys.argv[5] != "default" else None
tail = int(sys.argv[6]) if len(sys.argv) > 6 else 0
app, w, h, t, ov = WORKLOADS[wl]
p = default_params(w, h, t, **ov)
r = sbx.Renderer(app, variant=variant)
if variant is None:
    r.set_option("tail_waves_x100", tail)
    r.set_option("tail_max_waves_x100", 10 ** 6)
    r.set_option("coop_waves_x100", 0)
frame = torch.empty((h, w, 4), dtype=torch.float32, device="cuda")
kw = {} if parts == 1 else ({"tile_parts": parts, "tile_part": part} if split == "tiles" else {"shard": (int(split[4:]), parts, part)})
for _ in range(5):
    r.render_frame_part(p, frame.data_ptr(), stream=torch.cuda.current_stream().cuda_stream, **kw)
torch.cuda.synchronize()
print(r.timing())
r.close()
