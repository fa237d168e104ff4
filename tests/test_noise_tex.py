"""The USE_NOISE_TEX cloud path (src/app_clouds.h:8-9,51-55,69-81) as the app "APP_CLOUDS_TEX".

PARITY UNPINNED against the reference: that branch exists for its HLSL hosts only (hardware Texture3D.SampleLevel) and
the reference ships neither textures nor images.  The sampler is DEFINED in oracle/sbx_oracle.c (D3D11 linear filtering,
WRAP, 8-bit sub-texel weights); these tests pin the CUDA path to that definition bit for bit, and the definition to the
properties a trilinear WRAP sampler must have."""
import numpy as np
import pytest

from oracle import loader
from shaderbox_b200 import abi
from util import bits_equal, diff_report


def _volumes(size, seed=3):
    """Two different size^3 RGBA32F volumes: the ddsvolgen bake (.r) and a rolled / transposed copy of it."""
    a = loader.oracle_bake_volume(size)
    b = np.ascontiguousarray(np.roll(a.transpose(2, 0, 1, 3), 5, axis=0))
    return a, b


def _sample(vol_r, pos):
    """numpy statement of the sampler rule for one point (independent of the C code)."""
    n = vol_r.shape[0]
    idx, wts = [], []
    for u in pos:
        u = np.float32(u)
        uw = np.float32(u - np.floor(u))
        t = np.float32(np.float32(uw * np.float32(n)) - np.float32(0.5))
        fl = np.floor(t)
        f = np.float32(t - fl)
        w = np.float32(np.floor(np.float32(f * np.float32(256.0)) + np.float32(0.5)) / np.float32(256.0))
        i0 = int(fl) % n
        idx.append((i0, (i0 + 1) % n))
        wts.append(w)
    lerp = lambda a, b, w: np.float32(np.float32(a * np.float32(np.float32(1.0) - w)) + np.float32(b * w))
    (x0, x1), (y0, y1), (z0, z1) = idx
    wx, wy, wz = wts
    c00 = lerp(vol_r[z0, y0, x0], vol_r[z0, y0, x1], wx); c10 = lerp(vol_r[z0, y1, x0], vol_r[z0, y1, x1], wx)
    c01 = lerp(vol_r[z1, y0, x0], vol_r[z1, y0, x1], wx); c11 = lerp(vol_r[z1, y1, x0], vol_r[z1, y1, x1], wx)
    return lerp(lerp(c00, c10, wy), lerp(c01, c11, wy), wz)


def test_sampler_rule_properties():
    """Texel centres return the texel; the sampler wraps with period 1; a constant volume samples to the constant."""
    rng = np.random.default_rng(1)
    n = 8
    vol = rng.uniform(0, 1, (n, n, n)).astype(np.float32)
    for (x, y, z) in ((0, 0, 0), (3, 5, 7), (7, 7, 7)):
        centre = ((x + 0.5) / n, (y + 0.5) / n, (z + 0.5) / n)
        assert _sample(vol, centre) == vol[z, y, x]
        assert _sample(vol, (centre[0] + 3.0, centre[1] - 2.0, centre[2] + 1.0)) == vol[z, y, x]
    assert _sample(np.full((n, n, n), 0.375, np.float32), (0.123, 0.456, 0.789)) == np.float32(0.375)
    # halfway between two texel centres along x: the two texels averaged (weight 128/256)
    assert _sample(vol, (2.0 / n, 1.5 / n, 4.5 / n)) == np.float32(np.float32(vol[4, 1, 1] * np.float32(0.5)) + np.float32(vol[4, 1, 2] * np.float32(0.5)))


def test_c_sampler_equals_the_numpy_statement_of_the_rule():
    rng = np.random.default_rng(7)
    n = 16
    a = np.zeros((n, n, n, 4), np.float32)
    b = np.zeros((n, n, n, 4), np.float32)
    a[..., 0] = rng.uniform(-0.2, 1.2, (n, n, n)).astype(np.float32)
    b[..., 0] = rng.uniform(0, 1, (n, n, n)).astype(np.float32)
    a[..., 1:] = 9.0                                        # .g .b .a are never read
    loader.oracle_set_noise_volumes(a, b)
    pts = np.concatenate([rng.uniform(-3, 3, (3000, 3)), rng.uniform(0, 1, (1000, 3)),
                          [[0, 0, 0], [1, 1, 1], [0.5 / n, 0.5 / n, 0.5 / n], [-1e-9, 1 - 1e-9, 0.99999]]]).astype(np.float32)
    for x, y, z in pts:
        for which, vol in ((0, a), (1, b)):
            assert loader.oracle_sample_noise(which, x, y, z) == _sample(vol[..., 0], (x, y, z)), (which, x, y, z)


def test_oracle_frame_follows_the_sampler_rule_and_has_clouds():
    """The C restatement renders clouds from the volumes (the frame differs from an empty-volume frame), deterministically."""
    a, b = _volumes(16)
    loader.oracle_set_noise_volumes(a, b)
    p = abi.default_params(64, 36, 1.5)
    img = loader.oracle_render("APP_CLOUDS_TEX", p)
    assert np.isfinite(img).all() and (img[..., 3] == 1.0).all()
    assert bits_equal(img, loader.oracle_render("APP_CLOUDS_TEX", p))
    loader.oracle_set_noise_volumes(np.zeros_like(a), np.zeros_like(b))
    empty = loader.oracle_render("APP_CLOUDS_TEX", p)
    assert (img != empty).any(), "the volumes produce no cloud at all: the test frame would prove nothing"
    assert bits_equal(empty[:9], img[:9])                      # rows under the horizon are sky either way


@pytest.mark.gpu
@pytest.mark.parametrize("size,w,h,t,ov", [
    (16, 96, 54, 1.5, {}),
    (32, 203, 117, 0.0, {"cld_march_steps": 64, "cld_coverage": 0.6}),
    (32, 16, 9, 40.0, {}),                                     # few, wide pixels: lanes do not share a texel box (global-load path)
    (128, 320, 180, 2.5, {"cld_march_steps": 128, "sun_dir": (0.3, 0.2, -0.9), "wind_dir": (0.3, 0.0, 0.2)}),
])
def test_noise_texture_clouds_match_the_oracle(size, w, h, t, ov):
    import shaderbox_b200 as sbx

    a, b = _volumes(size)
    loader.oracle_set_noise_volumes(a, b)
    want = loader.oracle_render("APP_CLOUDS_TEX", abi.default_params(w, h, t, **ov))
    # "tma": the hand-written kernel (TMA-staged texel boxes);  "plugin" (= the default): the UNCHANGED src/app_clouds.h compiled with
    # -DUSE_NOISE_TEX, its `Texture3D ... : register(tN)` declarations and SampleLevel calls resolved by include/sbx/hlsl_tex.h
    for variant in ("tma", "plugin", None):
        r = sbx.Renderer("APP_CLOUDS_TEX", device=0, variant=variant)
        with pytest.raises(sbx.SbxError):
            r.render(w, h, u_time=t, **ov)                     # no textures yet: refused, not rendered from garbage
        r.set_noise_volumes(a, b)
        got = r.render(w, h, u_time=t, **ov)
        assert bits_equal(got, want), str(variant) + ": " + diff_report(got, want)
        assert bits_equal(r.render(w, h, u_time=t, shard=(4, 3, 1), **ov), want[abi.shard_rows(4, 3, 1, h)])
        r.close()


@pytest.mark.gpu
def test_device_baked_volume_feeds_the_texture_path():
    """sbx_bake_noise_volume (the ddsvolgen volume, made on the GPU) as texture 0, end to end on the device side."""
    import shaderbox_b200 as sbx

    r = sbx.Renderer("APP_CLOUDS_TEX", device=0)
    a = r.bake_noise_volume(32)
    assert bits_equal(a, loader.oracle_bake_volume(32))
    b = np.ascontiguousarray(a[::-1])
    r.set_noise_volumes(a, b)
    loader.oracle_set_noise_volumes(a, b)
    w, h, t = 160, 90, 3.0
    assert bits_equal(r.render(w, h, u_time=t), loader.oracle_render("APP_CLOUDS_TEX", abi.default_params(w, h, t)))
    r.close()
