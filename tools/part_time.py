"""Time every rank's part of a frame on ONE GPU (what each rank of an N-GPU run executes), for the ways the frame
can be cut and the kernel images that can march it:
    python tools/part_time.py <workload> <n_parts> [--variants a,b,..] [--splits rows4,rows1,tiles] [--tails 0,100] [--out f.json]
A variant "default" is sbx_load_app(app, NULL) with the hybrid tail option set from --tails.
Prints, per configuration: per-part kernel ms (CUDA events, L2 flushed before each), max, mean, ideal (whole frame / n)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import shaderbox_b200 as sbx
from bench import WORKLOADS
from shaderbox_b200.abi import default_params


def opt(name, default):
    return sys.argv[sys.argv.index(name) + 1] if name in sys.argv else default


wl, parts = sys.argv[1], int(sys.argv[2])
out_path = opt("--out", None)
variants = opt("--variants", "native,coop").split(",")
splits = opt("--splits", "rows4,rows1").split(",")
tails = [int(x) for x in opt("--tails", "0").split(",")]
app, w, h, t, ov = WORKLOADS[wl]
p = default_params(w, h, t, **ov)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
frame = torch.empty((h, w, 4), dtype=torch.float32, device="cuda")
s = torch.cuda.current_stream()


def timed(fn, reps=7, do_flush=True):
    for _ in range(2):
        fn()
    ms = []
    for _ in range(reps):
        if do_flush:
            flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s); fn(); e1.record(s)
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    return sorted(ms)[len(ms) // 2]


results = {}
r = sbx.Renderer(app, variant="native" if app in ("APP_CLOUDS", "APP_PLANET", "APP_RAYTRACER") else None)
full_ms = timed(lambda: r.render_frame_part(p, frame.data_ptr(), stream=s.cuda_stream))
tiny = default_params(8, 4, t)
rt = sbx.Renderer("APP_RAYTRACER")          # one warp of a cheap app: what a launch costs between two events
with_events = timed(lambda: rt.render_frame_part(tiny, frame.data_ptr(), stream=s.cuda_stream), reps=15)
rt.set_option("record_events", 0)
print("one-warp launch with the context's own timing events: %.1f us; without (below): " % (with_events * 1e3), end="")
fixed_flush = timed(lambda: rt.render_frame_part(tiny, frame.data_ptr(), stream=s.cuda_stream), reps=15)
fixed_warm = timed(lambda: rt.render_frame_part(tiny, frame.data_ptr(), stream=s.cuda_stream), reps=15, do_flush=False)
rt.close()
print("%s: whole frame %.4f ms -> ideal 1/%d = %.4f ms;  a one-warp RAYTRACER launch between events: %.1f us after an L2 flush, %.1f us warm" % (
    wl, full_ms, parts, full_ms / parts, fixed_flush * 1e3, fixed_warm * 1e3), flush=True)
results["full_ms"], results["fixed_us_flush"], results["fixed_us_warm"] = full_ms, fixed_flush * 1e3, fixed_warm * 1e3
r.close()

for variant in variants:
    for split in splits:
        for tail in (tails if variant == "default" else [0]):
            r = sbx.Renderer(app, variant=None if variant == "default" else variant)
            r.set_option("record_events", 0)
            if variant == "default":
                r.set_option("tail_waves_x100", tail)
                r.set_option("tail_max_waves_x100", 10 ** 6)
                r.set_option("coop_waves_x100", 0)
            ms = []
            for part in range(parts):
                kw = {"tile_parts": parts, "tile_part": part} if split == "tiles" else {"shard": (int(split[4:]), parts, part)}
                ms.append(timed(lambda: r.render_frame_part(p, frame.data_ptr(), stream=s.cuda_stream, **kw)))
            tm = r.timing()
            name = "%s/%s/tail%d" % (split, variant, tail)
            results[name] = {"ms": ms, "tail_rows": tm["tail_rows"], "grid": tm["grid_blocks"], "block": tm["block_threads"]}
            print("%-28s max %.4f mean %.4f (x%.3f of ideal; max/mean %.3f) tail_rows %d  parts: %s" % (
                name, max(ms), sum(ms) / len(ms), max(ms) / (full_ms / parts), max(ms) / (sum(ms) / len(ms)), tm["tail_rows"],
                " ".join("%.3f" % x for x in ms)), flush=True)
            r.close()
if out_path:
    json.dump(results, open(out_path, "w"), indent=1)
