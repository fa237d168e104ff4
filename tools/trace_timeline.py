"""Per-warp timeline of one launch (trace images, sbx_set_trace_buffer): where does a small launch lose its time?
    python tools/trace_timeline.py <workload> <n_parts> <part> <rows4|rows1|tiles> <variant_trace> [tail_waves_x100] [--out x.npz]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import shaderbox_b200 as sbx
from bench import WORKLOADS
from shaderbox_b200.abi import default_params

wl, parts, part, split, variant = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), sys.argv[4], sys.argv[5]
tail = int(sys.argv[6]) if len(sys.argv) > 6 and not sys.argv[6].startswith("--") else 0
out_path = sys.argv[sys.argv.index("--out") + 1] if "--out" in sys.argv else None
app, w, h, t, ov = WORKLOADS[wl]
p = default_params(w, h, t, **ov)
r = sbx.Renderer(app, variant=variant)
frame = torch.empty((h, w, 4), dtype=torch.float32, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
s = torch.cuda.current_stream()
kw = {"tile_parts": parts, "tile_part": part} if split == "tiles" else {"shard": (int(split[4:]), parts, part)}
if parts == 1:
    kw = {}
r.set_option("tail_waves_x100", tail)
r.set_option("tail_max_waves_x100", 10 ** 6)
for _ in range(3):
    r.render_frame_part(p, frame.data_ptr(), stream=s.cuda_stream, **kw)
torch.cuda.synchronize()
tm = r.timing()
nwarps = tm["grid_blocks"] * tm["block_threads"] // 32
rec = torch.zeros((nwarps, 4), dtype=torch.int64, device="cuda")
r.set_trace_buffer(rec.data_ptr())
flush.zero_()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(s)
r.render_frame_part(p, frame.data_ptr(), stream=s.cuda_stream, **kw)
e1.record(s)
torch.cuda.synchronize()
r.set_trace_buffer(0)
a = rec.cpu().numpy()
a = a[a[:, 1] > 0]
t0 = a[:, 0].min()
st, en, sm = (a[:, 0] - t0) * 1e-3, (a[:, 1] - t0) * 1e-3, a[:, 2]          # us
dur = en - st
span = en.max()
nsm = len(np.unique(sm))
slots = nsm * tm["blocks_per_sm"] * tm["block_threads"] // 32
print("%s %s part %d/%d %s tail %d: event %.1f us, trace span %.1f us, %d warps, %d SMs, tail_rows %d" % (
    wl, variant, part, parts, split, tail, e0.elapsed_time(e1) * 1e3, span, len(a), nsm, tm["tail_rows"]))
print("  warp duration us: mean %.1f p50 %.1f p90 %.1f p99 %.1f max %.1f;  sum(dur)/(span*slots) = %.3f" % (
    dur.mean(), np.percentile(dur, 50), np.percentile(dur, 90), np.percentile(dur, 99), dur.max(), dur.sum() / (span * slots)))
edges = np.linspace(0, span, 21)
occ = []
for lo, hi in zip(edges[:-1], edges[1:]):
    occ.append(float((np.clip(en, lo, hi) - np.clip(st, lo, hi)).sum() / ((hi - lo) * slots)))
print("  resident-warp occupancy per 5%% of the span: " + " ".join("%.2f" % x for x in occ))
order = np.argsort(en)[-5:]
print("  last 5 warps to finish: " + "; ".join("start %.0f dur %.0f (warp %d)" % (st[i], dur[i], i) for i in order))
print("  last warp started at %.1f us; warps longer than span/2: %d" % (st.max(), int((dur > span / 2).sum())))
per_sm_end = np.array([en[sm == k].max() for k in np.unique(sm)])
print("  per-SM last finish us: min %.1f mean %.1f max %.1f" % (per_sm_end.min(), per_sm_end.mean(), per_sm_end.max()))
if out_path:
    np.savez_compressed(out_path, rec=a, timing=np.array([e0.elapsed_time(e1)]))
r.close()
