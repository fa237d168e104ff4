"""First-light check on a B200: every app through the C ABI vs the _ref oracle, plus timings."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import shaderbox_b200 as sbx  # noqa: E402
from oracle import loader  # noqa: E402

out = {}
cases = [("APP_EGG", 256, 256, 0.0, {}), ("APP_EGG", 256, 256, 1.0, {}),
         ("APP_CLOUDS", 320, 180, 0.0, {}), ("APP_CLOUDS", 320, 180, 1.5, {"cld_march_steps": 128}),
         ("APP_ATMOSPHERE", 480, 270, 0.0, {}), ("APP_ATMOSPHERE", 480, 270, 1.0, {}),
         ("APP_PLANET", 320, 180, 0.0, {}), ("APP_PLANET", 320, 180, 2.0, {}),
         ("APP_RAYTRACER", 480, 270, 0.0, {}), ("APP_RAYTRACER", 480, 270, 1.0, {})]
rs = {}
for app, w, h, t, ov in cases:
    if app not in rs:
        rs[app] = sbx.Renderer(app, variant="plugin")
    r = rs[app]
    img = r.render(w, h, u_time=t, **ov)
    tm = r.timing()
    p = sbx.default_params(w, h, t, **ov)
    t0 = time.time()
    want = loader.ref_render(app, p)
    cpu_s = time.time() - t0
    both_nan = np.isnan(img) & np.isnan(want)
    d = np.abs(img - want)
    d[both_nan] = 0
    d = np.nan_to_num(d, nan=np.inf)
    key = "%s_%dx%d_t%g" % (app, w, h, t)
    out[key] = dict(max_abs=float(d.max()), n_gt_1e4=int((d > 1e-4).sum()), n_ne=int((d > 0).sum()),
                    nan_gpu=int(np.isnan(img).sum()), nan_ref=int(np.isnan(want).sum()),
                    kernel_ms=tm["kernel_ms"], regs=tm["regs_per_thread"], ctas_per_sm=tm["blocks_per_sm"], cpu_ref_s=cpu_s)
    print(key, out[key], flush=True)

# timings at the bench sizes (kernel only)
for app, w, h, t, ov in [("APP_CLOUDS", 1920, 1080, 1.5, {"cld_march_steps": 128}), ("APP_CLOUDS", 1920, 1080, 1.5, {}),
                         ("APP_ATMOSPHERE", 1920, 1080, 1.0, {}), ("APP_PLANET", 3840, 2160, 2.0, {}),
                         ("APP_RAYTRACER", 7680, 4320, 1.0, {}), ("APP_EGG", 1920, 1080, 1.0, {})]:
    r = rs[app]
    for use_hash in (1, 0):
        r.set_option("use_hash_table", use_hash)
        times = []
        for it in range(4):
            r.render(w, h, u_time=t, **ov)
            times.append(r.timing()["kernel_ms"])
        key = "time_%s_%dx%d_hash%d" % (app, w, h, use_hash)
        out[key] = dict(kernel_ms=times, mpix_s=w * h / min(times) * 1e-3)
        print(key, out[key], flush=True)
    r.set_option("use_hash_table", 1)

os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "first_light.json"), "w"), indent=1)
