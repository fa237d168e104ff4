"""ctypes mirror of include/sbx.h (struct layouts and status codes only; no compute here).

Field names and defaults follow the reference uniform block, src/uniform_buffer.h:26-58.
"""
import ctypes as C

SBX_OK = 0
SBX_ERR_INVALID = -1
SBX_ERR_NO_DEVICE = -2
SBX_ERR_UNKNOWN_APP = -3
SBX_ERR_CUDA = -4
SBX_ERR_COMPILE = -5
SBX_ERR_NOMEM = -6
SBX_ERR_UNSUPPORTED = -7

APPS = ("APP_EGG", "APP_CLOUDS", "APP_ATMOSPHERE", "APP_PLANET", "APP_RAYTRACER", "APP_SDF_AO", "APP_VINYL")
# + "APP_CLOUDS_TEX": APP_CLOUDS with USE_NOISE_TEX (hand-written image only: the branch is HLSL in the reference)


class Params(C.Structure):
    """sbx_params: u_res / u_time / u_mouse + the aux uniform block."""

    _fields_ = [
        ("width", C.c_int),
        ("height", C.c_int),
        ("u_time", C.c_float),
        ("u_mouse", C.c_float * 4),
        ("wind_dir", C.c_float * 3),
        ("sun_dir", C.c_float * 3),
        ("sun_color", C.c_float * 3),
        ("sun_power", C.c_float),
        ("cld_march_steps", C.c_int),
        ("illum_march_steps", C.c_int),
        ("sigma_scattering", C.c_float),
        ("cld_coverage", C.c_float),
        ("cld_thick", C.c_float),
        ("atm_radius", C.c_float),
        ("atm_ground_y", C.c_float),
        ("fog_density", C.c_float),
        ("fog_falloff", C.c_float),
    ]


class Shard(C.Structure):
    """sbx_shard: interleaved row stripes, stripe s -> part s % n_parts."""

    _fields_ = [("stripe_rows", C.c_int), ("n_parts", C.c_int), ("part", C.c_int)]


class Timing(C.Structure):
    _fields_ = [
        ("kernel_ms", C.c_float),
        ("h2d_ms", C.c_float),
        ("d2h_ms", C.c_float),
        ("launches", C.c_int),
        ("grid_blocks", C.c_int),
        ("block_threads", C.c_int),
        ("regs_per_thread", C.c_int),
        ("blocks_per_sm", C.c_int),
        ("zero_copy", C.c_int),
        ("lanes_per_pixel", C.c_int),
        ("tail_rows", C.c_int),
        ("tail_lanes_per_pixel", C.c_int),
    ]


class FramePart(C.Structure):
    """sbx_frame_part: one GPU's part of a full frame (row stripes and/or a checkerboard of warp tiles) + completion flag."""

    _fields_ = [("rows", Shard), ("tile_parts", C.c_int), ("tile_part", C.c_int), ("done_flag", C.c_void_p), ("done_value", C.c_uint)]


def default_params(width, height, u_time=0.0, **overrides):
    """Reference defaults, src/uniform_buffer.h:41-58 (pure Python twin of sbx_default_params)."""
    p = Params()
    p.width, p.height, p.u_time = int(width), int(height), float(u_time)
    p.u_mouse[:] = [0.0, 0.0, 0.0, 0.0]
    p.wind_dir[:] = [0.0, 0.0, 0.2]
    p.sun_dir[:] = [0.0, 0.0, -1.0]
    p.sun_color[:] = [1.0, 0.7, 0.55]
    p.sun_power = 8.0
    p.cld_march_steps = 100
    p.illum_march_steps = 6
    p.sigma_scattering = 0.15
    p.cld_coverage = 0.535
    p.cld_thick = 125.0
    p.atm_radius = 5000.0
    p.atm_ground_y = 4750.0
    p.fog_density = 0.1
    p.fog_falloff = 0.5
    for k, v in overrides.items():
        cur = getattr(p, k)
        if hasattr(cur, "__len__"):
            cur[:] = list(v)
        else:
            setattr(p, k, v)
    return p


def tile_part_mask(width, height, tile_parts, tile_part):
    """Boolean [height, width] mask of the pixels a tile part renders; twin of the kernel's mapping (sbx_kernel.cuh):
    the 8-pixel-wide tile column tx of a tile whose first row is r belongs to part (tx + r // 4) % tile_parts --
    for 8x4 tiles and for the one-row tiles of a hybrid launch's tail alike, pixel (x, y) belongs to
    part (x // 8 + y // 4) % tile_parts."""
    import numpy as np

    return (np.arange(width)[None, :] // 8 + np.arange(height)[:, None] // 4) % tile_parts == tile_part


def shard_rows(stripe_rows, n_parts, part, height):
    """Rows (in frame order) that a shard renders; twin of sbx_shard_rows."""
    stripe_rows = max(1, int(stripe_rows))
    return [y for y in range(height) if (y // stripe_rows) % n_parts == part]
