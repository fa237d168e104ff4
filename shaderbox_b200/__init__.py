"""shaderbox_b200 -- B200-native host for shaderbox app shaders (per-pixel mainImage path).

Product code only: the C ABI (include/sbx.h, libsbx.so), its ctypes mirror and the sm_100a kernel
images.  The CPU checkers live in oracle/ and are never imported from here.
"""
from .abi import APPS, FramePart, Params, Shard, Timing, default_params, shard_rows, tile_part_mask  # noqa: F401
from .host import MultiRenderer, Renderer, SbxError, compile_app, dds_volume_header, lib, EXPORTS, LIB_PATH  # noqa: F401
