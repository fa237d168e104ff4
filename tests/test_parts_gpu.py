"""GPU (one B200 is enough): the pieces of the N-GPU frame path behind the C ABI -- tile parts, the completion flag,
the hybrid image's tail, the single-process group (sbx_multi_*), pinned host frames, launches on several streams.
Everything is compared bit for bit with a plain one-launch render of the same frame."""
import ctypes as C

import numpy as np
import pytest
import torch

import shaderbox_b200 as sbx
from shaderbox_b200 import abi
from util import bits_equal

pytestmark = pytest.mark.gpu


def _full(app, w, h, t, variant=None, **ov):
    r = sbx.Renderer(app, device=0, variant=variant)
    try:
        return r.render(w, h, u_time=t, **ov)
    finally:
        r.close()


@pytest.mark.parametrize("app,w,h,t,ov", [
    ("APP_CLOUDS", 331, 203, 1.5, {"cld_march_steps": 48}),      # ragged: 331 = 41 tiles + 3 px, 203 = 50 tile rows + 3
    ("APP_PLANET", 256, 144, 2.0, {}),
    ("APP_EGG", 97, 61, 0.5, {}),
])
@pytest.mark.parametrize("parts", [2, 3, 8])
def test_tile_parts_partition_the_frame_exactly(app, w, h, t, ov, parts):
    """Each tile part renders exactly the pixels of abi.tile_part_mask (nothing else is touched), and together the
    parts are the frame -- bit for bit -- whatever image (one lane, hybrid tail) each launch picked."""
    want = _full(app, w, h, t, **ov)
    r = sbx.Renderer(app, device=0)
    p = abi.default_params(w, h, t, **ov)
    frame = torch.full((h, w, 4), float("nan"), dtype=torch.float32, device="cuda")
    covered = np.zeros((h, w), bool)
    s = torch.cuda.current_stream().cuda_stream
    for part in range(parts):
        frame.fill_(float("nan"))
        r.render_frame_part(p, frame.data_ptr(), tile_parts=parts, tile_part=part, stream=s)
        torch.cuda.synchronize()
        got = frame.cpu().numpy()
        mask = abi.tile_part_mask(w, h, parts, part)
        assert not np.isnan(got[mask]).any() and np.isnan(got[~mask]).all(), "part %d wrote the wrong pixels" % part
        assert bits_equal(got[mask], want[mask])
        assert not (covered & mask).any()
        covered |= mask
    assert covered.all()
    # rows and tiles combined: stripes of part 1 of 2, tiles of part 2 of 3
    frame.fill_(float("nan"))
    r.render_frame_part(p, frame.data_ptr(), shard=(4, 2, 1), tile_parts=3, tile_part=2, stream=s)
    torch.cuda.synchronize()
    got = frame.cpu().numpy()
    rows = abi.shard_rows(4, 2, 1, h)
    written = ~np.isnan(got[..., 0])
    assert written[rows].any() and not written[[y for y in range(h) if y not in set(rows)]].any()
    assert bits_equal(got[written], want[written])
    r.close()


def test_hybrid_tail_is_bit_identical_for_every_split():
    """The hybrid CLOUDS image marches the last rows of a launch with 4 lanes per pixel: any split gives the same frame."""
    w, h, t, ov = 640, 360, 1.5, {"cld_march_steps": 128}
    want = _full("APP_CLOUDS", w, h, t, variant="native", **ov)
    r = sbx.Renderer("APP_CLOUDS", device=0)
    seen = set()
    for tail in (0, 5, 30, 100, 100000):
        r.set_option("tail_waves_x100", tail)
        r.set_option("tail_max_waves_x100", 10 ** 6)
        got = r.render(w, h, u_time=t, **ov)
        tm = r.timing()
        seen.add(tm["tail_rows"])
        assert bits_equal(got, want), "tail_waves_x100=%d (tail rows %d)" % (tail, tm["tail_rows"])
        if tail == 0:
            assert tm["tail_rows"] == 0
        if tail == 100000:
            assert tm["tail_rows"] == h and tm["tail_lanes_per_pixel"] == 4
    assert len(seen) >= 3, seen
    # a shard of it, and an uneven far-from-default uniform block
    got = r.render(w, h, u_time=t, shard=(4, 8, 3), **ov)
    assert bits_equal(got, want[abi.shard_rows(4, 8, 3, h)])
    r.close()


def test_done_flag_is_published_after_the_pixels():
    w, h, t = 512, 288, 2.0
    r = sbx.Renderer("APP_CLOUDS", device=0)
    p = abi.default_params(w, h, t)
    want = r.render(w, h, u_time=t)
    frame = torch.zeros((h, w, 4), dtype=torch.float32, device="cuda")
    flags = torch.zeros(4, dtype=torch.int32, device="cuda")
    side = torch.cuda.Stream()
    main = torch.cuda.current_stream()
    torch.cuda.synchronize()
    for epoch in (1, 2, 3):
        frame.zero_()
        torch.cuda.synchronize()
        # the consumer is enqueued FIRST, on another stream: it can only run once the render kernels have published
        r.stream_wait_flags(flags.data_ptr(), 4, epoch, stream=side.cuda_stream)
        with torch.cuda.stream(side):
            copy = frame.clone()
        for part in range(4):
            r.render_frame_part(p, frame.data_ptr(), tile_parts=4, tile_part=part, done_flag=flags.data_ptr() + 4 * part,
                                done_value=epoch, stream=main.cuda_stream)
        torch.cuda.synchronize()
        assert flags.cpu().tolist() == [epoch] * 4
        assert bits_equal(copy.cpu().numpy(), want)
    r.close()


@pytest.mark.parametrize("devices", [[0], [0, 0], [0, 0, 0, 0, 0]])
def test_single_process_group_renders_the_same_frame(devices):
    """sbx_multi_*: one process, one part per listed device (a device listed twice shares the GPU)."""
    for app, w, h, t, ov in (("APP_CLOUDS", 480, 271, 1.5, {"cld_march_steps": 64}), ("APP_RAYTRACER", 333, 190, 1.0, {})):
        want = _full(app, w, h, t, **ov)
        m = sbx.MultiRenderer(app, devices=devices)
        p = abi.default_params(w, h, t, **ov)
        assert bits_equal(m.render(w, h, u_time=t, **ov), want)                     # pageable frame: device path + copy
        host = m.host_alloc(w * h * 16)                                            # pinned frame: every part stores into it
        arr = np.ctypeslib.as_array(C.cast(host, C.POINTER(C.c_float)), shape=(h, w, 4))
        for _ in range(3):
            arr[:] = np.nan
            m.render_host_ptr(p, host)
            assert bits_equal(np.array(arr), want)
        ptr = m.render_device(p)                                                   # frame stays on the first GPU
        m.sync()
        ms = m.kernel_ms()
        assert len(ms) == len(devices) and all(x > 0 for x in ms)
        r0 = sbx.Renderer(app, device=devices[0])
        assert bits_equal(r0.frame_read(ptr, h, w), want)
        r0.close()
        m.host_free(host)
        m.close()


def test_host_alloc_frames_take_the_zero_copy_path():
    w, h, t = 320, 180, 0.75
    r = sbx.Renderer("APP_ATMOSPHERE", device=0)
    want = r.render(w, h, u_time=t)
    assert r.timing()["zero_copy"] == 0                 # a numpy frame is pageable: rendered in HBM, then copied
    host = r.host_alloc(w * h * 16)
    arr = np.ctypeslib.as_array(C.cast(host, C.POINTER(C.c_float)), shape=(h, w, 4))
    arr[:] = np.nan
    r.render_host_ptr(abi.default_params(w, h, t), host)
    assert r.timing()["zero_copy"] == 1
    assert bits_equal(np.array(arr), want)
    r.host_free(host)
    r.close()


def test_launches_on_several_streams_share_one_context():
    """Per-context device state (math tables, hash memo) is complete before the call that builds it returns, and
    sequence launches carry their own times: two streams, no ordering between them, same frames."""
    w, h = 256, 144
    r = sbx.Renderer("APP_CLOUDS", device=0)            # fresh context: the first launch builds the tables
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    a = torch.empty((h, w, 4), dtype=torch.float32, device="cuda")
    b = torch.empty((3, h, w, 4), dtype=torch.float32, device="cuda")
    c = torch.empty((2, h, w, 4), dtype=torch.float32, device="cuda")
    p = abi.default_params(w, h, 1.0)
    r.render_into(p, a.data_ptr(), stream=s1.cuda_stream)                                   # builds the tables on s1
    r.render_sequence_into(p, [0.5, 1.0, 2.5], b.data_ptr(), stream=s2.cuda_stream)        # reads them on s2
    r.render_sequence_into(p, [7.0, 1.0], c.data_ptr(), stream=s1.cuda_stream)             # another times buffer in flight
    torch.cuda.synchronize()
    ref = sbx.Renderer("APP_CLOUDS", device=0)
    frames = {t: ref.render(w, h, u_time=t) for t in (0.5, 1.0, 2.5, 7.0)}
    ref.close()
    assert bits_equal(a.cpu().numpy(), frames[1.0])
    for k, t in enumerate((0.5, 1.0, 2.5)):
        assert bits_equal(b[k].cpu().numpy(), frames[t])
    for k, t in enumerate((7.0, 1.0)):
        assert bits_equal(c[k].cpu().numpy(), frames[t])
    r.close()


def test_bad_parts_are_rejected():
    r = sbx.Renderer("APP_EGG", device=0)
    p = abi.default_params(64, 64)
    buf = torch.zeros((64, 64, 4), dtype=torch.float32, device="cuda")
    for kw in ({"tile_parts": 4, "tile_part": 4}, {"tile_parts": 4, "tile_part": -1}, {"shard": (4, 2, 2)}, {"shard": (4, -2, 0)}):
        with pytest.raises(sbx.SbxError) as e:
            r.render_frame_part(p, buf.data_ptr(), **kw)
        assert e.value.status == abi.SBX_ERR_INVALID
    r.close()
