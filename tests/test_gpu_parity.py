"""GPU (B200): the CUDA path through the C ABI against the oracle.  Bar: BIT-EXACT fp32 frames
(stricter than the 1e-4 per-channel tolerance north_star states; the tolerance is still asserted
first so a failure says how far off it is)."""
import ctypes as C

import numpy as np
import pytest

import opcases
import shaderbox_b200 as sbx
from cases import FRAME_CASES, frame_key
from oracle import loader
from shaderbox_b200 import abi
from util import bits_equal, diff_report

pytestmark = pytest.mark.gpu
TOL = 1e-4   # north_star: per-channel |d| <= 1e-4 vs the C++ CPU reference


@pytest.fixture(scope="module")
def renderers():
    cache = {}

    def get(app, variant=None):
        key = (app, variant)
        if key not in cache:
            cache[key] = sbx.Renderer(app, device=0, variant=variant)
        return cache[key]

    yield get
    for r in cache.values():
        r.close()


def check(got, want):
    with np.errstate(invalid="ignore"):
        d = np.abs(got - want)
    d[np.isnan(got) & np.isnan(want)] = 0
    d[np.isinf(got) & (got == want)] = 0          # equal infinities (inf - inf is nan)
    assert float(np.nan_to_num(d, nan=np.inf).max(initial=0)) <= TOL, diff_report(got, want)
    assert bits_equal(got, want), diff_report(got, want)


@pytest.mark.parametrize("case", FRAME_CASES, ids=frame_key)
def test_plugin_frames_match_reference_golden(case, golden_frames, renderers):
    """UNCHANGED reference app headers compiled against the device operator library."""
    app, w, h, t, ov = case
    got = renderers(app, "plugin").render(w, h, u_time=t, **ov)
    check(got, golden_frames[frame_key(case)])


@pytest.mark.parametrize("case", [c for c in FRAME_CASES if c[0] != "APP_SDF_AO"], ids=frame_key)
def test_default_variant_frames_match_reference_golden(case, golden_frames, renderers):
    """Whatever sbx_load_app(app, NULL) selects (the hand-written native kernel where one exists)."""
    app, w, h, t, ov = case
    got = renderers(app, None).render(w, h, u_time=t, **ov)
    check(got, golden_frames[frame_key(case)])


@pytest.mark.parametrize("app,w,h,t,ov", [
    ("APP_CLOUDS", 256, 144, 3.25, {"cld_march_steps": 128}),
    ("APP_CLOUDS", 199, 101, 77.0, {"wind_dir": (0.1, 0.0, 0.3), "sun_dir": (0.3, 0.2, -0.9), "cld_coverage": 0.48}),
    ("APP_PLANET", 256, 144, 5.5, {}),
    ("APP_EGG", 200, 120, 2.75, {}),
    ("APP_ATMOSPHERE", 320, 180, 0.4, {}),
    ("APP_RAYTRACER", 320, 180, 2.5, {}),
])
def test_frames_match_oracle_on_other_inputs(app, w, h, t, ov, renderers):
    want = loader.oracle_render(app, abi.default_params(w, h, t, **ov))
    for variant in (None, "plugin"):
        check(renderers(app, variant).render(w, h, u_time=t, **ov), want)


@pytest.mark.parametrize("op", sorted(opcases.CASES))
def test_operator_known_answers(op, golden_ops, renderers):
    """Every operator of the device library vs the reference's own header on the same inputs."""
    a = golden_ops[op + "/in"]
    want = golden_ops[op + "/out"]
    got = renderers("APP_EGG", "plugin").eval_op(op, a, want.shape[1])
    assert bits_equal(got, want), op + ": " + diff_report(got, want)


@pytest.mark.parametrize("op", ["sinf", "cosf", "expf", "powf", "hash", "noise_iq", "fbm4", "sd_bezier", "noise_w"])
def test_operator_large_seeded_run_vs_oracle_ref(op, renderers):
    if not loader.have_ref():
        pytest.skip("oracle/_ref not shipped")
    a, ow = opcases.inputs(op, 200_000, seed=99)
    got = renderers("APP_EGG", "plugin").eval_op(op, a, ow)
    want = loader.ref_eval_op(op, a, ow)
    assert bits_equal(got, want), op + ": " + diff_report(got, want)


@pytest.mark.parametrize("ov", [
    {}, {"cld_march_steps": 128}, {"cld_coverage": 0.3}, {"cld_coverage": 0.75}, {"cld_coverage": 0.999},
    {"cld_coverage": 0.0}, {"illum_march_steps": 1, "cld_march_steps": 17}, {"sigma_scattering": 0.9, "cld_thick": 300.0},
    {"sun_dir": (0.0, 1.0, 0.0), "sun_power": 3.0}, {"u_mouse": (200.0, 0.0, 0.0, 0.0)},
    # |sigma * dt| >= 80: the march instantiation that keeps expf's range test (sbx_beer_lambert<false>)
    {"sigma_scattering": 100.0}, {"sigma_scattering": 40.0, "cld_thick": 4000.0, "cld_march_steps": 16}, {"sigma_scattering": -90.0},
])
def test_native_clouds_kernel_matches_oracle_over_the_uniform_space(ov, renderers):
    """The hand-written CLOUDS kernel skips work it can prove irrelevant (lazy octaves); the proof
    must hold for every uniform setting, not only the defaults."""
    w, h, t = 192, 108, 2.5
    want = loader.oracle_render("APP_CLOUDS", abi.default_params(w, h, t, **ov))
    check(renderers("APP_CLOUDS", "native").render(w, h, u_time=t, **ov), want)
    check(renderers("APP_CLOUDS", "coop").render(w, h, u_time=t, **ov), want)
    check(renderers("APP_CLOUDS", "plugin").render(w, h, u_time=t, **ov), want)


@pytest.mark.parametrize("t", [40.0, 700.0, 5000.0, 123456.0])
def test_native_clouds_far_from_the_origin(t, renderers):
    """wind_dir * u_time * 1000 moves the lattice indices out of the memo table: the hand-written kernel
    must notice and redo those pixels on the generic (arithmetic-hash) path."""
    w, h = 96, 54
    want = loader.oracle_render("APP_CLOUDS", abi.default_params(w, h, t))
    check(renderers("APP_CLOUDS", "native").render(w, h, u_time=t), want)
    check(renderers("APP_CLOUDS", "coop").render(w, h, u_time=t), want)


@pytest.mark.parametrize("steps", [128, 100, 37, 3, 1])
def test_cooperative_march_equals_one_lane_per_pixel(steps, renderers):
    """The cooperative image puts 4 lanes on a pixel (march steps i = 4r + phase) and replays the accumulation in
    step order from shuffles; ragged widths leave lanes without a pixel, step counts leave phases without a step."""
    for w, h in ((203, 61), (64, 9), (7, 3)):
        a = renderers("APP_CLOUDS", "native").render(w, h, u_time=1.5, cld_march_steps=steps)
        b = renderers("APP_CLOUDS", "coop").render(w, h, u_time=1.5, cld_march_steps=steps)
        assert renderers("APP_CLOUDS", "coop").timing()["lanes_per_pixel"] == 4
        assert bits_equal(a, b), (w, h, steps, diff_report(a, b))
        sh = (4, 8, 5)
        assert bits_equal(renderers("APP_CLOUDS", "coop").render(w, h, u_time=1.5, cld_march_steps=steps, shard=sh),
                          a[abi.shard_rows(*sh, h)])


def test_default_variant_picks_the_cooperative_image_for_small_grids(renderers):
    r = renderers("APP_CLOUDS", None)
    big = r.render(1920, 1080, u_time=1.5)
    assert r.timing()["lanes_per_pixel"] == 1
    part = r.render(1920, 1080, u_time=1.5, shard=(4, 8, 3))          # one rank's share at 8 GPUs: ~2 waves of warps
    assert r.timing()["lanes_per_pixel"] == 4
    assert bits_equal(part, big[abi.shard_rows(4, 8, 3, 1080)])
    quarter = r.render(1920, 1080, u_time=1.5, shard=(4, 4, 1))      # at 4 GPUs: ~4.6 waves -> 2 lanes per pixel
    assert r.timing()["lanes_per_pixel"] == 2
    assert bits_equal(quarter, big[abi.shard_rows(4, 4, 1, 1080)])
    r.set_option("coop_waves_x100", 0)
    try:
        part1 = r.render(1920, 1080, u_time=1.5, shard=(4, 8, 3))
        assert r.timing()["lanes_per_pixel"] == 1
    finally:
        r.set_option("coop_waves_x100", 250)
    assert bits_equal(part1, part)


@pytest.mark.parametrize("variant", ["plugin", "native"])
def test_hash_memo_table_is_bit_identical_to_arithmetic(variant, renderers):
    """noise_iq.h memoises hash(n) for integer n; rendering with and without the table must agree."""
    r = renderers("APP_CLOUDS", variant)
    a = r.render(160, 90, u_time=1.5)
    r.set_option("use_hash_table", 0)
    try:
        b = r.render(160, 90, u_time=1.5)
    finally:
        r.set_option("use_hash_table", 1)
    assert bits_equal(a, b)
    # lattice indices beyond the table (far from the origin) take the arithmetic path
    r.set_option("hash_table_log2", 10)
    try:
        c = r.render(160, 90, u_time=1.5)
    finally:
        r.set_option("hash_table_log2", 18)
    assert bits_equal(a, c)


@pytest.mark.parametrize("app", ["APP_CLOUDS", "APP_PLANET", "APP_RAYTRACER"])
def test_shards_reassemble_to_the_full_frame(app, renderers):
    w, h, t = 150, 91, 1.0
    r = renderers(app, None)
    full = r.render(w, h, u_time=t)
    for stripe, parts in ((4, 2), (4, 8), (1, 3), (7, 4), (128, 2)):
        out = np.full_like(full, np.nan)
        for part in range(parts):
            rows = abi.shard_rows(stripe, parts, part, h)
            got = r.render(w, h, u_time=t, shard=(stripe, parts, part))
            assert got.shape[0] == len(rows)
            if rows:
                out[rows] = got
        assert bits_equal(out, full), (stripe, parts)


def test_device_render_and_unshard_kernel(renderers):
    import torch

    w, h, t, stripe, parts = 333, 77, 2.0, 4, 3
    r = renderers("APP_PLANET", None)
    p = abi.default_params(w, h, t)
    full = torch.empty((h, w, 4), dtype=torch.float32, device="cuda:0")
    r.render_into(p, full.data_ptr(), stream=torch.cuda.current_stream().cuda_stream)
    frame = torch.full((h, w, 4), float("nan"), dtype=torch.float32, device="cuda:0")
    for part in range(parts):
        n = len(abi.shard_rows(stripe, parts, part, h))
        buf = torch.empty((n, w, 4), dtype=torch.float32, device="cuda:0")
        s = torch.cuda.current_stream().cuda_stream
        r.render_into(p, buf.data_ptr(), shard=(stripe, parts, part), stream=s)
        r.unshard(w, h, (stripe, parts, part), buf.data_ptr(), frame.data_ptr(), stream=s)
    torch.cuda.synchronize()
    assert bits_equal(frame.cpu().numpy(), full.cpu().numpy())
    assert bits_equal(full.cpu().numpy(), r.render(w, h, u_time=t))


def test_host_frame_paths_agree(renderers):
    """sbx_render_host into pageable memory (render in HBM + copy) and into a pinned, mapped frame
    (kernel stores go straight to the host over PCIe) give the same bits."""
    import torch

    w, h, t = 640, 360, 1.5
    r = renderers("APP_CLOUDS", None)
    p = abi.default_params(w, h, t)
    pageable = r.render(w, h, u_time=t)
    assert r.timing()["zero_copy"] == 0
    pinned = torch.full((h, w, 4), float("nan"), dtype=torch.float32).pin_memory()
    r.render_host_ptr(p, pinned.data_ptr())
    assert r.timing()["zero_copy"] == 1
    assert bits_equal(pinned.numpy(), pageable)
    r.set_option("host_zero_copy", 0)
    try:
        pinned.fill_(float("nan"))
        r.render_host_ptr(p, pinned.data_ptr())
        assert r.timing()["zero_copy"] == 0
        assert bits_equal(pinned.numpy(), pageable)
    finally:
        r.set_option("host_zero_copy", 1)
    # sharded into a pinned part
    part = torch.empty((len(abi.shard_rows(4, 3, 1, h)), w, 4), dtype=torch.float32).pin_memory()
    r.render_host_ptr(p, part.data_ptr(), shard=(4, 3, 1))
    assert bits_equal(part.numpy(), pageable[abi.shard_rows(4, 3, 1, h)])


@pytest.mark.parametrize("app,w,h", [("APP_EGG", 96, 64), ("APP_VINYL", 64, 36), ("APP_CLOUDS", 120, 67), ("APP_PLANET", 80, 45)])
def test_time_sequence_in_one_launch_equals_single_frames(app, w, h, renderers):
    times = [0.0, 0.25, 1.0, 1.5, 7.75]
    r = renderers(app, None)
    seq = r.render_sequence(w, h, times)
    assert seq.shape == (len(times), h, w, 4)
    for k, t in enumerate(times):
        assert bits_equal(seq[k], r.render(w, h, u_time=t)), (app, t)
    part = r.render_sequence(w, h, times[1:3], shard=(4, 3, 1))
    for k, t in enumerate(times[1:3]):
        assert bits_equal(part[k], r.render(w, h, u_time=t, shard=(4, 3, 1)))


def unorm8(frame):
    """FLOAT -> R8G8B8A8_UNORM of the reference's swap chain (util/hlsltoy/src/hlsltoy.cpp:192): NaN -> 0, clamp,
    * 255 + 0.5 in fp32, truncate."""
    f = np.where(np.isnan(frame), np.float32(0), frame).astype(np.float32)
    f = np.minimum(np.maximum(f, np.float32(0)), np.float32(1))
    return (f * np.float32(255) + np.float32(0.5)).astype(np.float32).astype(np.int32).astype(np.uint8)


@pytest.mark.parametrize("case", [c for c in FRAME_CASES if c[1] > 1], ids=frame_key)
def test_rgba8_frames_match_quantised_reference_golden(case, golden_frames, renderers):
    app, w, h, t, ov = case
    want = unorm8(golden_frames[frame_key(case)])
    for variant in ((None, "plugin") if app != "APP_SDF_AO" else ("plugin",)):
        got = renderers(app, variant).render_rgba8(w, h, u_time=t, **ov)
        assert got.dtype == np.uint8 and (got == want).all(), (variant, int((got != want).sum()))


def test_rgba8_paths_shards_and_nan(renderers):
    import torch

    w, h, t = 333, 77, 2.0
    r = renderers("APP_PLANET", None)
    full = unorm8(r.render(w, h, u_time=t))
    assert (r.render_rgba8(w, h, u_time=t) == full).all()
    p = abi.default_params(w, h, t)
    # device frame, pinned zero-copy frame, pageable frame, shard
    dev = torch.zeros((h, w, 4), dtype=torch.uint8, device="cuda:0")
    r.render_rgba8_into(p, dev.data_ptr(), stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert (dev.cpu().numpy() == full).all()
    pinned = torch.zeros((h, w, 4), dtype=torch.uint8).pin_memory()
    r.render_rgba8_host_ptr(p, pinned.data_ptr())
    assert r.timing()["zero_copy"] == 1 and (pinned.numpy() == full).all()
    rows = abi.shard_rows(4, 3, 2, h)
    assert (r.render_rgba8(w, h, u_time=t, shard=(4, 3, 2)) == full[rows]).all()
    # the quantiser itself, incl. NaN / out-of-range / rounding ties, against the numpy statement
    vals = np.array([np.nan, -1.0, -0.0, 0.0, 0.5 / 255, 1.5 / 255, 0.499999 / 255, 1.0, 1.0001, np.inf, -np.inf,
                     0.7, 127.5 / 255, 254.5 / 255, 1e-30], np.float32)
    assert unorm8(vals).tolist() == [0, 0, 0, 0, 1, 2, 0, 255, 255, 255, 0, 179, 128, 255, 0]
    rng = np.random.default_rng(5)
    more = np.concatenate([vals, rng.uniform(-0.1, 1.1, 100000).astype(np.float32), (np.arange(0, 511) / np.float32(510)).astype(np.float32)])
    got = renderers("APP_EGG", "plugin").eval_op("unorm8", more, 1)[:, 0]
    assert (got.astype(np.uint8) == unorm8(more)).all()


def test_deterministic_run_to_run(renderers):
    r = renderers("APP_CLOUDS", None)
    a = r.render(320, 180, u_time=4.0)
    for _ in range(3):
        assert bits_equal(a, r.render(320, 180, u_time=4.0))


# ---- BASELINE.json configs at full size ----------------------------------------------------------
# (app, w, h, u_time, overrides, row step): every `step`-th row of the frame is compared; step 1 = the whole frame
FULL = [
    ("APP_EGG", 256, 256, 0.0, {}, 1),                                   # configs[0], the reference's own CPU-runnable case
    ("APP_EGG", 256, 256, 1.0, {}, 1),
    ("APP_CLOUDS", 1920, 1080, 1.5, {"cld_march_steps": 128}, 1),        # configs[1], the metric's configuration: EVERY pixel
    ("APP_ATMOSPHERE", 1920, 1080, 1.0, {}, 4),
    ("APP_PLANET", 3840, 2160, 2.0, {}, 8),
    ("APP_RAYTRACER", 7680, 4320, 1.0, {}, 8),
    ("APP_CLOUDS", 3840, 2160, 1.5, {"cld_march_steps": 128}, 8),        # the metric's scene on the larger frames north_star names
    ("APP_CLOUDS", 7680, 4320, 1.5, {"cld_march_steps": 128}, 8),
]


def _checker():
    """The reference's own headers (oracle/_ref) when they travelled to this box, else the plain-C restatement."""
    return (loader.ref_render, "oracle/_ref") if loader.have_ref() else (loader.oracle_render, "oracle/sbx_oracle.c")


@pytest.mark.parametrize("app,w,h,t,ov,step", FULL, ids=["%s_%dx%d_t%g" % c[:4] for c in FULL])
def test_full_size_frames_match_the_reference(app, w, h, t, ov, step, renderers):
    """At the bench sizes: the metric's frame (and the EGG correctness config) pixel for pixel, the other configs on
    every `step`-th row (>= 1/8 of the frame), against the reference's own headers; the whole frame must be finite
    with alpha 1, and the 8-way partitions (row stripes, tile checkerboard) must reproduce it bit for bit."""
    render, _ = _checker()
    r = renderers(app, None)
    full = r.render(w, h, u_time=t, **ov)
    assert np.isfinite(full).all() and (full[..., 3] == 1.0).all()
    sh = abi.Shard(1, step, step // 2)            # rows step//2, step//2 + step, ...
    rows = abi.shard_rows(1, step, step // 2, h)
    assert len(rows) * 8 >= h
    want = render(app, abi.default_params(w, h, t, **ov), shard=sh)
    check(full[rows], want)
    # the 8-GPU partitions of the frame reproduce it bit for bit
    for part in (0, 5):
        got = r.render(w, h, u_time=t, shard=(4, 8, part), **ov)
        assert bits_equal(got, full[abi.shard_rows(4, 8, part, h)])


def test_plugin_full_size_clouds_rows_match_the_reference(renderers):
    """The UNCHANGED app_clouds.h at the metric's size, every 8th row."""
    app, w, h, t, ov = "APP_CLOUDS", 1920, 1080, 1.5, {"cld_march_steps": 128}
    render, _ = _checker()
    got = renderers(app, "plugin").render(w, h, u_time=t, shard=(1, 8, 3), **ov)
    check(got, render(app, abi.default_params(w, h, t, **ov), shard=abi.Shard(1, 8, 3)))


def test_errors_are_loud(renderers):
    lib = sbx.lib()
    with pytest.raises(sbx.SbxError) as e:
        sbx.Renderer("APP_NOPE")
    assert e.value.status == abi.SBX_ERR_UNKNOWN_APP
    r = renderers("APP_EGG", None)
    p = abi.default_params(8, 8)
    bad = abi.Shard(4, 2, 5)
    buf = np.zeros((8, 8, 4), np.float32)
    assert lib.sbx_render_host(r._ctx, C.byref(p), C.byref(bad), buf.ctypes.data_as(C.c_void_p)) == abi.SBX_ERR_INVALID
    with pytest.raises(sbx.SbxError):
        r.eval_op("no_such_op", np.zeros((4, 1), np.float32), 1)


def test_smoke_entry():
    import __graft_entry__

    __graft_entry__.smoke()
