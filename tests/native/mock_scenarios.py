"""Scenarios run by tests/test_host_mock_cpu.py in a CHILD process whose LD_LIBRARY_PATH points at the mock driver
(tests/native/fake_cuda.c built as libcuda.so.1 in a temporary directory).  Test infrastructure only: the mock records
kernel launches and renders nothing; what is checked here is the host logic of libsbx.so -- which kernel image it picks,
the launch it plans, that the plan covers every pixel exactly once under the kernel's warp -> pixel mapping
(shaderbox_b200/include/sbx/sbx_launch.h), the N-GPU group and its flag protocol, and the error paths.

usage: python tests/native/mock_scenarios.py <scenario>      (prints one JSON object; exit code 0 = scenario ran)
"""
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import shaderbox_b200 as sbx  # noqa: E402
from shaderbox_b200 import abi  # noqa: E402


# ---- ctypes mirrors of the mock's launch record and of sbx_launch (sbx/sbx_launch.h) -------------------------------
class Region(C.Structure):
    _fields_ = [("warps", C.c_int), ("tiles_per_row", C.c_int), ("row0", C.c_int), ("rows", C.c_int), ("tile_rows", C.c_int),
                ("first_tile_row", C.c_int), ("magic", C.c_ulonglong)]


class Launch(C.Structure):
    _fields_ = [("p", abi.Params), ("stripe_rows", C.c_int), ("n_parts", C.c_int), ("part", C.c_int), ("local_rows", C.c_int),
                ("col_parts", C.c_int), ("col_part", C.c_int), ("reg", Region * 2), ("out", C.c_void_p), ("hash_tab", C.c_void_p),
                ("hash_bias", C.c_int), ("hash_len", C.c_int), ("hash_span", C.c_int), ("out_is_frame", C.c_int), ("out_rgba8", C.c_int),
                ("times", C.c_void_p), ("lut", C.c_void_p), ("done_counter", C.c_void_p), ("done_flag", C.c_void_p),
                ("done_value", C.c_uint), ("trace", C.c_void_p)]


class FakeLaunch(C.Structure):
    _fields_ = [("name", C.c_char * 64), ("grid", C.c_uint * 3), ("block", C.c_uint * 3), ("smem", C.c_uint), ("device", C.c_int),
                ("stream", C.c_void_p), ("n_params", C.c_int), ("param_size", C.c_uint * 4), ("param", (C.c_ubyte * 1024) * 4)]


def mock():
    m = C.CDLL("libcuda.so.1")
    assert hasattr(m, "fake_cuda_launch_count"), "the real driver was loaded, not the mock"
    m.fake_cuda_launch_count.restype = C.c_size_t
    m.fake_cuda_sizeof_launch.restype = C.c_size_t
    m.fake_cuda_get_launch.argtypes = [C.c_size_t, C.POINTER(FakeLaunch)]
    m.fake_cuda_live_allocs.restype = C.c_long
    m.fake_cuda_fail_alloc_at.argtypes = [C.c_long]
    assert m.fake_cuda_sizeof_launch() == C.sizeof(FakeLaunch)
    return m


def launches(m, name=b"sbx_render"):
    out = []
    for i in range(m.fake_cuda_launch_count()):
        rec = FakeLaunch()
        assert m.fake_cuda_get_launch(i, C.byref(rec))
        if name is None or rec.name == name:
            out.append(rec)
    return out


def launch_block(rec):
    assert rec.param_size[0] == C.sizeof(Launch), "kernel parameter size %d != host struct %d" % (rec.param_size[0], C.sizeof(Launch))
    return Launch.from_buffer_copy(bytes(rec.param[0])[:C.sizeof(Launch)])


# ---- the kernel's warp -> pixel mapping, restated from the contract in sbx_launch.h --------------------------------
def stores_of(rec, lanes_per_pixel, hybrid_lanes=0):
    """Flat pixel indices (into L.out, in pixels) that the launch `rec` stores, one entry per storing lane."""
    L = launch_block(rec)
    wpc = rec.block[0] // 32
    warp = np.arange(rec.grid[0] * wpc, dtype=np.int64)
    at_all = []
    for region in (0, 1):
        R = L.reg[region]
        if hybrid_lanes > 1:
            sel = (warp >= L.reg[0].warps) if region == 1 else (warp < L.reg[0].warps)
            w = warp[sel] - (L.reg[0].warps if region == 1 else 0)
            P = hybrid_lanes if region == 1 else 1
        else:
            if region == 1:
                assert R.warps == 0, "only hybrid images use the second region"
                continue
            w, P = warp, lanes_per_pixel
        if len(w) == 0 or R.tiles_per_row == 0:
            assert R.warps == 0
            continue
        tile_w, tile_h = (8, 4) if P == 1 else (32 // P, 1)
        trow = w // R.tiles_per_row
        if R.magic:   # the host promises the magic number divides exactly for EVERY warp index of the rounded-up grid
            fast = (w.astype(object) * int(R.magic)) >> 40
            assert (np.array(fast, dtype=np.int64) == trow).all(), "magic division is off for some warp"
        tile_x = w - trow * R.tiles_per_row
        live = w < R.warps
        assert R.warps == R.tile_rows * R.tiles_per_row and R.tile_rows == -(-R.rows // tile_h)
        assert 0 <= R.first_tile_row < max(R.tile_rows, 1)
        trow = np.where(live, (trow + R.first_tile_row) % max(R.tile_rows, 1), trow)
        lr0 = R.row0 + trow * tile_h
        if L.col_parts > 1:
            first = (L.col_part - (lr0 >> 2)) % L.col_parts
            tile_x = first + tile_x * L.col_parts
        lane = np.arange(32, dtype=np.int64)
        if P == 1:
            x = tile_x[:, None] * 8 + (lane & 7)[None, :]
            lr = lr0[:, None] + (lane // 8)[None, :]
            valid = live[:, None] & (x < L.p.width) & (lr < R.row0 + R.rows)
        else:
            x = tile_x[:, None] * tile_w + (lane // P)[None, :]
            lr = np.broadcast_to(lr0[:, None], x.shape)
            valid = live[:, None] & (x < L.p.width) & (lr < R.row0 + R.rows) & ((lane % P) == 0)[None, :]
        x, lr = x[valid], lr[valid]
        assert (lr >= 0).all() and (lr < L.local_rows).all()
        y = lr if L.n_parts == 1 else ((lr // L.stripe_rows) * L.n_parts + L.part) * L.stripe_rows + lr % L.stripe_rows
        assert (y < L.p.height).all()
        at_all.append((y if L.out_is_frame else lr) * L.p.width + x)
    return L, (np.concatenate(at_all) if at_all else np.zeros(0, np.int64))


def expected_pixels(L):
    """What the launch is supposed to render: its rows (all columns, or its share of the 8-pixel tile columns)."""
    w, h = L.p.width, L.p.height
    rows = np.array(abi.shard_rows(L.stripe_rows, L.n_parts, L.part, h), dtype=np.int64)
    if not L.out_is_frame:
        return (np.arange(len(rows))[:, None] * w + np.arange(w)[None, :]).ravel()
    if L.col_parts > 1:
        mask = abi.tile_part_mask(w, h, L.col_parts, L.col_part)
        sel = np.zeros((h, w), bool)
        sel[rows] = True
        return np.flatnonzero((mask & sel).ravel())
    return (rows[:, None] * w + np.arange(w)[None, :]).ravel()


def check_exactly_once(rec, timing):
    L, at = stores_of(rec, timing["lanes_per_pixel"], timing["tail_lanes_per_pixel"])
    want = expected_pixels(L)
    got = np.sort(at)
    assert len(got) == len(want) and (got == np.sort(want)).all(), "the launch does not store every pixel of its part exactly once"
    return L


# ---- scenarios -------------------------------------------------------------------------------------------------------
def scenario_plan():
    """Every shipped image, a spread of frame sizes and partitions: the planned launch covers its part exactly once."""
    m = mock()
    report = {"launches": 0, "images": {}}
    sizes = [(1920, 1080), (256, 144), (1, 1), (7, 5), (33, 9), (1000, 37), (3840, 2160)]
    parts = [None, (4, 8, 0), (4, 8, 5), (1, 3, 2), (8, 2, 1), (4, 8, 7)]
    for app in sbx.APPS:             # (APP_CLOUDS_TEX needs its textures: scenario_noise_tex)
        for variant in (None, "plugin", "coop", "coop2", "hybrid"):
            if variant in ("coop", "coop2", "hybrid") and app != "APP_CLOUDS":
                continue
            try:
                r = sbx.Renderer(app, device=0, variant=variant)
            except sbx.SbxError:
                continue
            key = "%s.%s" % (app, variant or "default")
            report["images"][key] = []
            for (w, h) in sizes:
                if w * h > 3_000_000 and app != "APP_CLOUDS":
                    continue
                for shard in parts:
                    p = abi.default_params(w, h, 1.0)
                    rows = len(abi.shard_rows(*(shard or (1, 1, 0)), h))
                    out = r.frame_alloc(max(rows, 1) * w * 16)
                    m.fake_cuda_reset_log()
                    r.render_into(p, out, shard=shard)
                    recs = launches(m)
                    if rows == 0:
                        assert len(recs) == 0
                    else:
                        assert len(recs) == 1
                        tm = r.timing()
                        L = check_exactly_once(recs[0], tm)
                        assert recs[0].grid[0] * (recs[0].block[0] // 32) >= L.reg[0].warps + L.reg[1].warps
                        assert recs[0].block[0] == tm["block_threads"] and recs[0].grid[0] == tm["grid_blocks"]
                        report["launches"] += 1
                        if (w, h) == (1920, 1080):
                            report["images"][key].append({"shard": shard, "lanes_per_pixel": tm["lanes_per_pixel"], "grid": recs[0].grid[0],
                                                          "regs": tm["regs_per_thread"], "blocks_per_sm": tm["blocks_per_sm"],
                                                          "first_tile_row": L.reg[0].first_tile_row})
                    r.frame_free(out)
            r.close()
    report["live_allocs_after_close"] = m.fake_cuda_live_allocs()
    return report


def scenario_frame_parts():
    """sbx_render_frame_part: row stripes x tile checkerboards of one frame are disjoint and their union is the frame."""
    m = mock()
    r = sbx.Renderer("APP_CLOUDS", device=0)
    report = {"cases": 0}
    for (w, h) in [(1920, 1080), (250, 130), (64, 4)]:
        for (stripe, n_rows, n_tiles) in [(4, 8, 1), (4, 1, 8), (4, 2, 4), (8, 3, 3)]:
            frame = r.frame_alloc(w * h * 16)
            flags = r.frame_alloc(4096)
            seen = np.zeros(w * h, np.int32)
            for rp in range(n_rows):
                for tp in range(n_tiles):
                    m.fake_cuda_reset_log()
                    r.render_frame_part(abi.default_params(w, h, 1.0), frame, shard=(stripe, n_rows, rp), tile_parts=n_tiles, tile_part=tp,
                                        done_flag=flags + 4 * tp, done_value=7)
                    recs = launches(m)
                    assert len(recs) == 1
                    L, at = stores_of(recs[0], r.timing()["lanes_per_pixel"], r.timing()["tail_lanes_per_pixel"])
                    assert L.out == frame and L.out_is_frame == 1 and L.done_flag == flags + 4 * tp and L.done_value == 7 and L.done_counter
                    np.add.at(seen, at, 1)
                    assert C.c_uint.from_address(flags + 4 * tp).value == 7     # the mock honours the kernel's completion flag
            assert (seen == 1).all(), "parts overlap or leave holes at %dx%d %r" % (w, h, (stripe, n_rows, n_tiles))
            report["cases"] += 1
            r.frame_free(frame)
            r.frame_free(flags)
    r.close()
    report["live_allocs_after_close"] = m.fake_cuda_live_allocs()
    return report


def scenario_image_choice():
    """Which CLOUDS image a launch gets: one lane per pixel for a full frame, 2 and 4 lanes as a GPU's share shrinks."""
    m = mock()
    r = sbx.Renderer("APP_CLOUDS", device=0)
    out = r.frame_alloc(1920 * 1080 * 16)
    picks = {}
    for n in (1, 2, 4, 8, 16):
        r.render_frame(abi.default_params(1920, 1080, 1.5), out, shard=(4, n, 0))
        picks[str(n)] = r.timing()["lanes_per_pixel"]
    r.frame_free(out)
    r.close()
    return {"lanes_per_pixel_by_parts": picks, "live_allocs_after_close": m.fake_cuda_live_allocs()}


def scenario_multi():
    """The single-process group on 8 mock GPUs: one launch per GPU per frame, the union is the frame, flags complete, no deadlock."""
    m = mock()
    n = sbx.lib().sbx_device_count()
    g = sbx.MultiRenderer("APP_CLOUDS", n_gpus=n)
    w, h = 1920, 1080
    p = abi.default_params(w, h, 1.5)
    report = {"gpus": n}
    for mode in ("device", "pinned_host", "pageable_host"):
        m.fake_cuda_reset_log()
        if mode == "device":
            g.render_device(p)
            g.sync()
        elif mode == "pinned_host":
            ptr = g.host_alloc(w * h * 16)
            g.render_host_ptr(p, ptr)
            g.host_free(ptr)
        else:
            buf = np.empty((h, w, 4), np.float32)
            g.render_host_ptr(p, buf.ctypes.data)
        recs = launches(m)
        assert len(recs) == n and sorted(rec.device for rec in recs) == list(range(n))
        seen = np.zeros(w * h, np.int32)
        outs = set()
        for rec in recs:
            L, at = stores_of(rec, g.timing(rec.device)["lanes_per_pixel"], 0)
            assert (L.n_parts, L.stripe_rows, L.out_is_frame) == (n, 4, 1) and L.part == rec.device
            outs.add(L.out)
            np.add.at(seen, at, 1)
        assert (seen == 1).all() and len(outs) == 1
        report[mode] = {"launches": len(recs)}
    # back-to-back frames: the worker hand-off must neither deadlock nor lose a frame
    m.fake_cuda_reset_log()
    ptr = g.host_alloc(64 * 64 * 16)
    small = abi.default_params(64, 64, 0.0)
    frames = 3000
    for _ in range(frames):
        g.render_host_ptr(small, ptr)
    assert m.fake_cuda_launch_count() == frames * n
    g.host_free(ptr)
    report["stress_frames"] = frames
    g.close()
    report["live_allocs_after_close"] = m.fake_cuda_live_allocs()
    return report


def scenario_sequence():
    """A u_time sequence is ONE launch (grid.y = frames) with its own device copy of the times; rgba8 output sets its flag."""
    m = mock()
    r = sbx.Renderer("APP_CLOUDS", device=0)
    w, h, times = 250, 130, [0.0, 0.25, 0.5, 0.75, 1.0]
    out = r.frame_alloc(w * h * 16 * len(times))
    p = abi.default_params(w, h, 0.0)
    seen_times = []
    for k in range(12):            # more launches than staging slots: the ring wraps
        m.fake_cuda_reset_log()
        r.render_sequence_into(p, [t + k for t in times], out)
        recs = launches(m)
        assert len(recs) == 1 and recs[0].grid[1] == len(times)
        L = check_exactly_once(recs[0], r.timing())
        assert L.times and not L.out_rgba8
        seen_times.append(L.times)
    m.fake_cuda_reset_log()
    r.render_rgba8_into(p, out)
    rec = launches(m)[0]
    assert launch_block(rec).out_rgba8 == 1 and rec.grid[1] == 1
    r.frame_free(out)
    r.close()
    return {"launches": 13, "live_allocs_after_close": m.fake_cuda_live_allocs()}


def scenario_errors():
    """Allocation failures at every point of create -> load -> render -> multi: an error status, no crash, no leak."""
    m = mock()
    report = {"failed_at": []}
    for k in range(1, 12):
        m.fake_cuda_fail_alloc_at(k)
        status = "ok"
        try:
            r = sbx.Renderer("APP_CLOUDS", device=0)
            try:
                r.render(64, 32, u_time=1.0)
                r.render_rgba8(64, 32, u_time=1.0)
                r.render_sequence(32, 16, [0.0, 0.5])
            except sbx.SbxError as e:
                status = "render: %d" % e.status
            r.close()
        except sbx.SbxError as e:
            status = "create/load: %d" % e.status
        report["failed_at"].append(status)
        assert m.fake_cuda_live_allocs() == 0, "leak after a failed allocation #%d (%s)" % (k, status)
    m.fake_cuda_fail_alloc_at(0)
    assert any(s != "ok" for s in report["failed_at"])
    return report


def scenario_noise_tex():
    """The USE_NOISE_TEX images take a second kernel parameter (the TMA descriptors) and refuse to launch without textures."""
    m = mock()
    report = {}
    for variant in ("plugin", "tma"):
        r = sbx.Renderer("APP_CLOUDS_TEX", device=0, variant=variant)
        refused = False
        try:
            r.render(64, 32, u_time=1.0)
        except sbx.SbxError as e:
            refused = e.status == -1
        assert refused, "a textured image must not launch without its textures"
        n = 16
        vol = np.random.default_rng(1).random((n, n, n, 4), dtype=np.float32)
        r.set_noise_volumes(vol, vol[::-1].copy())
        m.fake_cuda_reset_log()
        out = r.frame_alloc(64 * 32 * 16)
        r.render_into(abi.default_params(64, 32, 1.0), out)
        recs = launches(m)
        assert len(recs) == 1 and recs[0].n_params == 2
        check_exactly_once(recs[0], r.timing())
        report[variant] = {"param_sizes": [recs[0].param_size[0], recs[0].param_size[1]]}
        r.frame_free(out)
        r.close()
    report["live_allocs_after_close"] = m.fake_cuda_live_allocs()
    return report


def scenario_multi_launch_failure():
    """A launch that fails on one GPU of the group: the frame call reports which part failed, nothing hangs, and the group
    renders the next frame normally."""
    m = mock()
    m.fake_cuda_fail_launches.argtypes = [C.c_int, C.c_int]
    g = sbx.MultiRenderer("APP_CLOUDS", n_gpus=8)
    p = abi.default_params(256, 144, 1.0)
    ptr = g.host_alloc(256 * 144 * 16)
    report = {}
    for dev in (0, 5):
        m.fake_cuda_fail_launches(dev, 1)
        try:
            g.render_host_ptr(p, ptr)
            report[str(dev)] = "no error"
        except sbx.SbxError as e:
            report[str(dev)] = str(e)
        m.fake_cuda_reset_log()
        g.render_host_ptr(p, ptr)              # the next frame is complete again
        assert m.fake_cuda_launch_count() == 8
        g.render_device(p)
        g.sync()
    g.host_free(ptr)
    g.close()
    report["live_allocs_after_close"] = m.fake_cuda_live_allocs()
    return report


def scenario_wrong_device():
    """A GPU that is not sm_100 is refused at sbx_create (run with SBX_FAKE_CC_MAJOR=9)."""
    try:
        sbx.Renderer("APP_CLOUDS", device=0)
    except sbx.SbxError as e:
        return {"status": e.status, "message": str(e)}
    return {"status": 0}


def scenario_no_peer():
    """Without a peer path between the GPUs the group cannot be formed (run with SBX_FAKE_NO_PEER=1, SBX_FAKE_GPUS=2)."""
    m = mock()
    try:
        sbx.MultiRenderer("APP_CLOUDS", n_gpus=2)
    except sbx.SbxError as e:
        return {"status": e.status, "live_allocs": m.fake_cuda_live_allocs()}
    return {"status": 0}


def scenario_flags():
    """sbx_stream_write_flag / sbx_stream_wait_flags on pinned memory: GEQ comparison across wrap-around, bad arguments."""
    r = sbx.Renderer("APP_CLOUDS", device=0)
    flags = r.host_alloc(4096)
    C.memset(flags, 0, 4096)
    for i in range(8):
        r.stream_write_flag(flags + 4 * i, 41 + i)
    r.stream_wait_flags(flags, 8, 41)
    got = [C.c_uint.from_address(flags + 4 * i).value for i in range(8)]
    r.stream_write_flag(flags, 0xfffffffe)
    r.stream_wait_flags(flags, 1, 0xfffffff0)        # (int)(have - want) >= 0
    bad = []
    for args in ((0, 1, 1), (flags, 0, 1), (flags, -1, 1)):
        try:
            r.stream_wait_flags(*args)
            bad.append(0)
        except sbx.SbxError as e:
            bad.append(e.status)
    r.host_free(flags)
    r.close()
    return {"flags": got, "bad_argument_status": bad}


if __name__ == "__main__":
    fn = globals()["scenario_" + sys.argv[1]]
    print(json.dumps(fn()))
