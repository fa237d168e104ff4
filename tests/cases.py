"""Frame cases shared by tools/make_golden.py and the parity tests.

(app, width, height, u_time, uniform overrides).  Small enough that the CPU oracle renders each in
well under a second; the BASELINE.json configs at full size are covered by size-independent
properties in test_gpu_parity.py."""

FRAME_CASES = [
    ("APP_EGG", 64, 64, 0.0, {}),
    ("APP_EGG", 80, 45, 1.0, {}),
    ("APP_CLOUDS", 80, 45, 0.0, {}),
    ("APP_CLOUDS", 80, 45, 1.5, {"cld_march_steps": 128}),
    ("APP_CLOUDS", 48, 27, 10.0, {"cld_march_steps": 40, "illum_march_steps": 3, "cld_coverage": 0.6,
                                   "sun_dir": (0.2, 0.3, -0.9)}),
    ("APP_ATMOSPHERE", 80, 45, 0.0, {}),
    ("APP_ATMOSPHERE", 80, 45, 1.0, {}),
    ("APP_PLANET", 80, 45, 0.0, {}),
    ("APP_PLANET", 80, 45, 2.0, {}),
    ("APP_RAYTRACER", 80, 45, 0.0, {}),
    ("APP_RAYTRACER", 80, 45, 1.0, {}),
    ("APP_SDF_AO", 80, 45, 0.5, {}),
    ("APP_VINYL", 80, 45, 0.0, {}),
    ("APP_VINYL", 96, 54, 1.25, {}),
    # ragged sizes: not a multiple of the 8x4 warp tile, 1-pixel rows/columns
    ("APP_EGG", 37, 19, 0.5, {}),
    ("APP_RAYTRACER", 1, 7, 0.0, {}),
    ("APP_ATMOSPHERE", 13, 1, 1.0, {}),
]


def frame_key(case):
    app, w, h, t, ov = case
    extra = "".join("_%s=%s" % (k, ",".join(map(str, v)) if isinstance(v, tuple) else v) for k, v in sorted(ov.items()))
    return "%s_%dx%d_t%g%s" % (app, w, h, t, extra)
