"""Host-side mirror of the reference's plugin interface over the C ABI (include/sbx.h).

A shaderbox "host" provides u_res / u_time / u_mouse (+ the aux uniform block) and calls
mainImage once per pixel (src/main.h:6-53, src/uniform_buffer.h:26-58).  `Renderer` is that host:
it selects an app by its reference define (src/Makefile:9, `-DAPP_PLANET`), takes the uniforms as
keyword arguments under their reference names, and returns the RGBA32F frame.

This module is plumbing only: ctypes calls into shaderbox_b200/libsbx.so.  There is no Python or
CPU implementation of the pixel path here -- if the library or a B200 is missing, calls raise.
"""
import ctypes as C
import os

import numpy as np

from .abi import FramePart, Params, Shard, Timing, default_params, shard_rows, SBX_OK

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libsbx.so")

_lib = None


class SbxError(RuntimeError):
    def __init__(self, status, where, detail=""):
        self.status = status
        super().__init__("%s failed: %d %s" % (where, status, detail))


def lib():
    """Load libsbx.so (built in-tree by __graft_entry__.build()); raises if it is missing."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise SbxError(-2, "load", "%s not built (run `python -c 'import __graft_entry__ as g; g.build()'`)" % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        P = C.POINTER
        L.sbx_version.restype = C.c_char_p
        L.sbx_strerror.restype = C.c_char_p
        L.sbx_strerror.argtypes = [C.c_int]
        L.sbx_last_error.restype = C.c_char_p
        L.sbx_last_error.argtypes = [C.c_void_p]
        L.sbx_default_params.argtypes = [P(Params), C.c_int, C.c_int]
        L.sbx_create.argtypes = [C.c_int, P(C.c_void_p)]
        L.sbx_destroy.argtypes = [C.c_void_p]
        L.sbx_destroy.restype = None
        L.sbx_load_app.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p]
        L.sbx_compile_app.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p, C.c_char_p]
        L.sbx_shard_rows.argtypes = [P(Shard), C.c_int]
        L.sbx_render_device.argtypes = [C.c_void_p, P(Params), P(Shard), C.c_void_p, C.c_void_p]
        L.sbx_render_host.argtypes = [C.c_void_p, P(Params), P(Shard), C.c_void_p]
        L.sbx_render_host_rgba8.argtypes = [C.c_void_p, P(Params), P(Shard), C.c_void_p]
        L.sbx_render_device_rgba8.argtypes = [C.c_void_p, P(Params), P(Shard), C.c_void_p, C.c_void_p]
        L.sbx_render_sequence_device.argtypes = [C.c_void_p, P(Params), P(Shard), C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.sbx_render_sequence_host.argtypes = [C.c_void_p, P(Params), P(Shard), C.c_void_p, C.c_int, C.c_void_p]
        L.sbx_render_frame.argtypes = [C.c_void_p, P(Params), P(Shard), C.c_void_p, C.c_void_p]
        L.sbx_render_frame_part.argtypes = [C.c_void_p, P(Params), P(FramePart), C.c_void_p, C.c_void_p]
        L.sbx_stream_wait_flags.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_uint, C.c_void_p]
        L.sbx_stream_write_flag.argtypes = [C.c_void_p, C.c_void_p, C.c_uint, C.c_void_p]
        L.sbx_host_alloc.argtypes = [C.c_void_p, C.c_size_t, P(C.c_void_p)]
        L.sbx_host_free.argtypes = [C.c_void_p, C.c_void_p]
        L.sbx_multi_create.argtypes = [P(C.c_int), C.c_int, P(C.c_void_p)]
        L.sbx_multi_destroy.argtypes = [C.c_void_p]
        L.sbx_multi_destroy.restype = None
        L.sbx_multi_gpus.argtypes = [C.c_void_p]
        L.sbx_multi_ctx.argtypes = [C.c_void_p, C.c_int]
        L.sbx_multi_ctx.restype = C.c_void_p
        L.sbx_multi_load_app.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p]
        L.sbx_multi_set_option.argtypes = [C.c_void_p, C.c_char_p, C.c_int]
        L.sbx_multi_render_device.argtypes = [C.c_void_p, P(Params), P(C.c_void_p)]
        L.sbx_multi_render_host.argtypes = [C.c_void_p, P(Params), C.c_void_p]
        L.sbx_multi_stream.argtypes = [C.c_void_p]
        L.sbx_multi_stream.restype = C.c_void_p
        L.sbx_multi_sync.argtypes = [C.c_void_p]
        L.sbx_multi_last_timing.argtypes = [C.c_void_p, P(C.c_float), C.c_int]
        L.sbx_multi_last_error.argtypes = [C.c_void_p]
        L.sbx_multi_last_error.restype = C.c_char_p
        L.sbx_set_trace_buffer.argtypes = [C.c_void_p, C.c_void_p]
        L.sbx_set_noise_volumes.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        L.sbx_frame_alloc.argtypes = [C.c_void_p, C.c_size_t, P(C.c_void_p)]
        L.sbx_frame_free.argtypes = [C.c_void_p, C.c_void_p]
        L.sbx_frame_export.argtypes = [C.c_void_p, C.c_void_p, C.c_char_p]
        L.sbx_frame_import.argtypes = [C.c_void_p, C.c_char_p, P(C.c_void_p)]
        L.sbx_frame_release.argtypes = [C.c_void_p, C.c_void_p]
        L.sbx_host_frame_register.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, P(C.c_void_p)]
        L.sbx_host_frame_unregister.argtypes = [C.c_void_p, C.c_void_p]
        L.sbx_bake_noise_volume_device.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        L.sbx_bake_noise_volume_host.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]
        L.sbx_dds_volume_header.argtypes = [C.c_int, C.c_void_p, C.c_int]
        L.sbx_frame_read.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
        L.sbx_unshard_device.argtypes = [C.c_void_p, C.c_int, C.c_int, P(Shard), C.c_void_p, C.c_void_p, C.c_void_p]
        L.sbx_set_option.argtypes = [C.c_void_p, C.c_char_p, C.c_int]
        L.sbx_last_timing.argtypes = [C.c_void_p, P(Timing)]
        L.sbx_eval_op.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int]
        _lib = L
    return _lib


EXPORTS = (
    "sbx_default_params", "sbx_create", "sbx_destroy", "sbx_load_app", "sbx_compile_app", "sbx_shard_rows",
    "sbx_render_device", "sbx_render_host", "sbx_unshard_device", "sbx_set_option", "sbx_last_timing",
    "sbx_last_error", "sbx_strerror", "sbx_version", "sbx_eval_op", "sbx_render_frame", "sbx_frame_alloc",
    "sbx_frame_free", "sbx_frame_export", "sbx_frame_import", "sbx_frame_release", "sbx_frame_read",
    "sbx_render_host_rgba8", "sbx_render_device_rgba8", "sbx_render_sequence_device", "sbx_render_sequence_host",
    "sbx_host_frame_register", "sbx_host_frame_unregister",
    "sbx_bake_noise_volume_device", "sbx_bake_noise_volume_host", "sbx_dds_volume_header",
    "sbx_render_frame_part", "sbx_stream_wait_flags", "sbx_host_alloc", "sbx_host_free",
    "sbx_multi_create", "sbx_multi_destroy", "sbx_multi_gpus", "sbx_multi_ctx", "sbx_multi_load_app", "sbx_multi_set_option",
    "sbx_multi_render_device", "sbx_multi_render_host", "sbx_multi_stream", "sbx_multi_sync", "sbx_multi_last_timing",
    "sbx_multi_last_error", "sbx_device_count", "sbx_set_trace_buffer", "sbx_stream_write_flag", "sbx_set_noise_volumes",
)


def dds_volume_header(size):
    """The 148-byte DDS + DX10 header ddsvolgen writes in front of a size^3 RGBA32F volume (needs no GPU)."""
    buf = C.create_string_buffer(148)
    n = lib().sbx_dds_volume_header(int(size), buf, 148)
    if n != 148:
        raise SbxError(n, "sbx_dds_volume_header")
    return buf.raw


def compile_app(header_path, app_name, image_out_path):
    """Compile an unchanged shaderbox app header to an sm_100a kernel image (no GPU needed)."""
    L = lib()
    rc = L.sbx_compile_app(None, os.fsencode(header_path), app_name.encode(), os.fsencode(image_out_path))
    if rc != SBX_OK:
        raise SbxError(rc, "sbx_compile_app", (L.sbx_last_error(None) or b"").decode(errors="replace"))
    return image_out_path


class Renderer:
    """One context on one GPU.  Usage:  r = Renderer("APP_CLOUDS"); img = r.render(1920, 1080, u_time=1.5)"""

    def __init__(self, app, device=0, variant=None):
        self._L = lib()
        self._ctx = C.c_void_p()
        rc = self._L.sbx_create(int(device), C.byref(self._ctx))
        if rc != SBX_OK:
            raise SbxError(rc, "sbx_create", (self._L.sbx_last_error(None) or b"").decode(errors="replace"))
        self.app = None
        self.device = int(device)
        self.load_app(app, variant)

    def _check(self, rc, where):
        if rc != SBX_OK:
            raise SbxError(rc, where, (self._L.sbx_last_error(self._ctx) or b"").decode(errors="replace"))

    def load_app(self, app, variant=None):
        self._check(self._L.sbx_load_app(self._ctx, app.encode(), variant.encode() if variant else None), "sbx_load_app")
        self.app = app

    def set_option(self, key, value):
        self._check(self._L.sbx_set_option(self._ctx, key.encode(), int(value)), "sbx_set_option")

    def close(self):
        if self._ctx:
            self._L.sbx_destroy(self._ctx)
            self._ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @staticmethod
    def params(width, height, u_time=0.0, **uniforms):
        return default_params(width, height, u_time, **uniforms)

    def render(self, width, height, u_time=0.0, shard=None, out=None, **uniforms):
        """Render to host memory; returns float32 [rows, width, 4] (row 0 = fragCoord.y 0.5)."""
        p = uniforms.pop("params", None) or default_params(width, height, u_time, **uniforms)
        sh = Shard(*shard) if shard is not None else Shard(1, 1, 0)
        rows = len(shard_rows(sh.stripe_rows, sh.n_parts, sh.part, p.height))
        if out is None:
            out = np.empty((rows, p.width, 4), dtype=np.float32)
        assert out.dtype == np.float32 and out.size == rows * p.width * 4 and out.flags["C_CONTIGUOUS"]
        self._check(self._L.sbx_render_host(self._ctx, C.byref(p), C.byref(sh), out.ctypes.data_as(C.c_void_p)),
                    "sbx_render_host")
        return out

    def render_rgba8(self, width, height, u_time=0.0, shard=None, out=None, **uniforms):
        """Render to host memory as R8G8B8A8_UNORM (the reference's 8-bit swap chain); returns uint8 [rows, width, 4]."""
        p = uniforms.pop("params", None) or default_params(width, height, u_time, **uniforms)
        sh = Shard(*shard) if shard is not None else Shard(1, 1, 0)
        rows = len(shard_rows(sh.stripe_rows, sh.n_parts, sh.part, p.height))
        if out is None:
            out = np.empty((rows, p.width, 4), dtype=np.uint8)
        assert out.dtype == np.uint8 and out.size == rows * p.width * 4 and out.flags["C_CONTIGUOUS"]
        self._check(self._L.sbx_render_host_rgba8(self._ctx, C.byref(p), C.byref(sh), out.ctypes.data_as(C.c_void_p)),
                    "sbx_render_host_rgba8")
        return out

    def render_rgba8_host_ptr(self, params, host_ptr, shard=None):
        sh = Shard(*shard) if shard is not None else Shard(1, 1, 0)
        self._check(self._L.sbx_render_host_rgba8(self._ctx, C.byref(params), C.byref(sh), C.c_void_p(host_ptr)),
                    "sbx_render_host_rgba8")

    def render_rgba8_into(self, params, dev_ptr, shard=None, stream=0):
        sh = Shard(*shard) if shard is not None else Shard(1, 1, 0)
        self._check(self._L.sbx_render_device_rgba8(self._ctx, C.byref(params), C.byref(sh), C.c_void_p(dev_ptr),
                                                    C.c_void_p(stream)), "sbx_render_device_rgba8")

    def render_sequence(self, width, height, times, shard=None, **uniforms):
        """Frames for u_time = times[k], all in one launch; returns float32 [len(times), rows, width, 4]."""
        p = uniforms.pop("params", None) or default_params(width, height, 0.0, **uniforms)
        sh = Shard(*shard) if shard is not None else Shard(1, 1, 0)
        rows = len(shard_rows(sh.stripe_rows, sh.n_parts, sh.part, p.height))
        t = np.ascontiguousarray(times, dtype=np.float32)
        out = np.empty((len(t), rows, p.width, 4), dtype=np.float32)
        self._check(self._L.sbx_render_sequence_host(self._ctx, C.byref(p), C.byref(sh), t.ctypes.data_as(C.c_void_p), len(t),
                                                     out.ctypes.data_as(C.c_void_p)), "sbx_render_sequence_host")
        return out

    def render_sequence_into(self, params, times, dev_ptr, shard=None, stream=0):
        sh = Shard(*shard) if shard is not None else Shard(1, 1, 0)
        t = np.ascontiguousarray(times, dtype=np.float32)
        self._check(self._L.sbx_render_sequence_device(self._ctx, C.byref(params), C.byref(sh), t.ctypes.data_as(C.c_void_p), len(t),
                                                       C.c_void_p(dev_ptr), C.c_void_p(stream)), "sbx_render_sequence_device")

    def render_into(self, params, dev_ptr, shard=None, stream=0):
        """Render into device memory at `dev_ptr` (e.g. a torch tensor's data_ptr()) on `stream`."""
        sh = Shard(*shard) if shard is not None else Shard(1, 1, 0)
        self._check(self._L.sbx_render_device(self._ctx, C.byref(params), C.byref(sh), C.c_void_p(dev_ptr),
                                              C.c_void_p(stream)), "sbx_render_device")

    def render_frame(self, params, dev_frame_ptr, shard=None, stream=0):
        """Render this shard's rows straight into a FULL frame at `dev_frame_ptr` (may be a peer GPU's)."""
        sh = Shard(*shard) if shard is not None else Shard(1, 1, 0)
        self._check(self._L.sbx_render_frame(self._ctx, C.byref(params), C.byref(sh), C.c_void_p(dev_frame_ptr),
                                             C.c_void_p(stream)), "sbx_render_frame")

    def render_frame_part(self, params, dev_frame_ptr, shard=None, tile_parts=1, tile_part=0, done_flag=0, done_value=0, stream=0):
        """This GPU's part of a full frame (row stripes and/or a checkerboard of warp tiles) straight into the frame at
        `dev_frame_ptr`; the launch's last thread block stores `done_value` at `done_flag` (device-visible address)."""
        part = FramePart(Shard(*shard) if shard is not None else Shard(1, 1, 0), int(tile_parts), int(tile_part),
                         C.c_void_p(done_flag or None), int(done_value))
        self._check(self._L.sbx_render_frame_part(self._ctx, C.byref(params), C.byref(part), C.c_void_p(dev_frame_ptr),
                                                  C.c_void_p(stream)), "sbx_render_frame_part")

    def stream_write_flag(self, dev_flag_ptr, value, stream=0):
        """After the work enqueued on `stream` so far: store `value` at the 32-bit flag (own, peer or mapped host memory)."""
        self._check(self._L.sbx_stream_write_flag(self._ctx, C.c_void_p(dev_flag_ptr), int(value), C.c_void_p(stream)),
                    "sbx_stream_write_flag")

    def stream_wait_flags(self, dev_flags_ptr, n, value, stream=0):
        """Stream-ordered wait (on the device) until each of the n 32-bit flags at `dev_flags_ptr` is >= value."""
        self._check(self._L.sbx_stream_wait_flags(self._ctx, C.c_void_p(dev_flags_ptr), int(n), int(value), C.c_void_p(stream)),
                    "sbx_stream_wait_flags")

    def set_noise_volumes(self, vol_a, vol_b):
        """The two size^3 RGBA32F noise textures of APP_CLOUDS_TEX (the USE_NOISE_TEX cloud path); only .r is sampled."""
        a = np.ascontiguousarray(vol_a, dtype=np.float32)
        b = np.ascontiguousarray(vol_b, dtype=np.float32)
        assert a.shape == b.shape and a.ndim == 4 and a.shape[0] == a.shape[1] == a.shape[2] and a.shape[3] == 4
        self._check(self._L.sbx_set_noise_volumes(self._ctx, a.ctypes.data_as(C.c_void_p), b.ctypes.data_as(C.c_void_p), int(a.shape[0])),
                    "sbx_set_noise_volumes")

    def set_trace_buffer(self, dev_ptr):
        self._check(self._L.sbx_set_trace_buffer(self._ctx, C.c_void_p(dev_ptr or None)), "sbx_set_trace_buffer")

    def host_alloc(self, nbytes):
        """A pinned + mapped host frame (sbx_host_alloc): sbx_render_host stores into it directly from the kernel."""
        out = C.c_void_p()
        self._check(self._L.sbx_host_alloc(self._ctx, nbytes, C.byref(out)), "sbx_host_alloc")
        return out.value

    def host_free(self, ptr):
        self._check(self._L.sbx_host_free(self._ctx, C.c_void_p(ptr)), "sbx_host_free")

    def frame_alloc(self, nbytes):
        out = C.c_void_p()
        self._check(self._L.sbx_frame_alloc(self._ctx, nbytes, C.byref(out)), "sbx_frame_alloc")
        return out.value

    def frame_free(self, ptr):
        self._check(self._L.sbx_frame_free(self._ctx, C.c_void_p(ptr)), "sbx_frame_free")

    def frame_export(self, ptr):
        buf = C.create_string_buffer(64)
        self._check(self._L.sbx_frame_export(self._ctx, C.c_void_p(ptr), buf), "sbx_frame_export")
        return buf.raw

    def frame_import(self, handle):
        out = C.c_void_p()
        self._check(self._L.sbx_frame_import(self._ctx, handle, C.byref(out)), "sbx_frame_import")
        return out.value

    def frame_release(self, ptr):
        self._check(self._L.sbx_frame_release(self._ctx, C.c_void_p(ptr)), "sbx_frame_release")

    def host_frame_register(self, host_ptr, nbytes):
        """Pin + map a (shared) host frame for this GPU; returns the device alias for render_frame."""
        out = C.c_void_p()
        self._check(self._L.sbx_host_frame_register(self._ctx, C.c_void_p(host_ptr), nbytes, C.byref(out)), "sbx_host_frame_register")
        return out.value

    def host_frame_unregister(self, host_ptr):
        self._check(self._L.sbx_host_frame_unregister(self._ctx, C.c_void_p(host_ptr)), "sbx_host_frame_unregister")

    def bake_noise_volume(self, size, z0=0, nz=None):
        """Slices [z0, z0+nz) of the ddsvolgen noise volume; returns float32 [nz, size, size, 4]."""
        nz = size - z0 if nz is None else nz
        out = np.empty((nz, size, size, 4), dtype=np.float32)
        self._check(self._L.sbx_bake_noise_volume_host(self._ctx, int(size), int(z0), int(nz), out.ctypes.data_as(C.c_void_p)),
                    "sbx_bake_noise_volume_host")
        return out

    def bake_noise_volume_into(self, size, dev_ptr, z0=0, nz=None, stream=0):
        nz = size - z0 if nz is None else nz
        self._check(self._L.sbx_bake_noise_volume_device(self._ctx, int(size), int(z0), int(nz), C.c_void_p(dev_ptr), C.c_void_p(stream)),
                    "sbx_bake_noise_volume_device")

    def frame_read(self, ptr, height, width, stream=0):
        out = np.empty((height, width, 4), dtype=np.float32)
        self._check(self._L.sbx_frame_read(self._ctx, C.c_void_p(ptr), out.ctypes.data_as(C.c_void_p), out.nbytes,
                                           C.c_void_p(stream)), "sbx_frame_read")
        return out

    def render_host_ptr(self, params, host_ptr, shard=None):
        """Render + copy to a host buffer given by address (e.g. a pinned torch tensor)."""
        sh = Shard(*shard) if shard is not None else Shard(1, 1, 0)
        self._check(self._L.sbx_render_host(self._ctx, C.byref(params), C.byref(sh), C.c_void_p(host_ptr)),
                    "sbx_render_host")

    def unshard(self, width, height, shard, dev_part, dev_frame, stream=0):
        sh = Shard(*shard)
        self._check(self._L.sbx_unshard_device(self._ctx, width, height, C.byref(sh), C.c_void_p(dev_part),
                                               C.c_void_p(dev_frame), C.c_void_p(stream)), "sbx_unshard_device")

    def timing(self):
        t = Timing()
        self._check(self._L.sbx_last_timing(self._ctx, C.byref(t)), "sbx_last_timing")
        return {k: getattr(t, k) for k, _ in Timing._fields_}

    def eval_op(self, op, inputs, out_width):
        """Evaluate one operator of the device library on the rows of `inputs` (test hook)."""
        a = np.ascontiguousarray(inputs, dtype=np.float32)
        if a.ndim == 1:
            a = a[:, None]
        out = np.zeros((a.shape[0], out_width), dtype=np.float32)
        self._check(self._L.sbx_eval_op(self._ctx, op.encode(), a.ctypes.data_as(C.c_void_p), a.shape[1],
                                        out.ctypes.data_as(C.c_void_p), out_width, a.shape[0]), "sbx_eval_op")
        return out


class MultiRenderer:
    """One frame over several GPUs from THIS process (sbx_multi_*): every GPU renders its 4-row stripes of
    the frame straight into the destination.  `devices` may repeat a device (the parts then share it)."""

    def __init__(self, app, devices=None, n_gpus=None, variant=None):
        self._L = lib()
        self._m = C.c_void_p()
        if devices is None:
            devices = list(range(int(n_gpus or 1)))
        self.devices = [int(d) for d in devices]
        arr = (C.c_int * len(self.devices))(*self.devices)
        rc = self._L.sbx_multi_create(arr, len(self.devices), C.byref(self._m))
        if rc != SBX_OK:
            raise SbxError(rc, "sbx_multi_create", (self._L.sbx_last_error(None) or b"").decode(errors="replace"))
        self._check(self._L.sbx_multi_load_app(self._m, app.encode(), variant.encode() if variant else None), "sbx_multi_load_app")
        self.app = app

    def _check(self, rc, where):
        if rc != SBX_OK:
            raise SbxError(rc, where, (self._L.sbx_multi_last_error(self._m) or b"").decode(errors="replace"))

    def set_option(self, key, value):
        self._check(self._L.sbx_multi_set_option(self._m, key.encode(), int(value)), "sbx_multi_set_option")

    def close(self):
        if self._m:
            self._L.sbx_multi_destroy(self._m)
            self._m = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def host_alloc(self, nbytes):
        out = C.c_void_p()
        self._check(self._L.sbx_host_alloc(self._L.sbx_multi_ctx(self._m, 0), nbytes, C.byref(out)), "sbx_host_alloc")
        return out.value

    def host_free(self, ptr):
        self._check(self._L.sbx_host_free(self._L.sbx_multi_ctx(self._m, 0), C.c_void_p(ptr)), "sbx_host_free")

    def render(self, width, height, u_time=0.0, out=None, **uniforms):
        """Render to host memory (pageable numpy frame: assembled on the first GPU, then copied)."""
        p = uniforms.pop("params", None) or default_params(width, height, u_time, **uniforms)
        if out is None:
            out = np.empty((p.height, p.width, 4), dtype=np.float32)
        self._check(self._L.sbx_multi_render_host(self._m, C.byref(p), out.ctypes.data_as(C.c_void_p)), "sbx_multi_render_host")
        return out

    def render_host_ptr(self, params, host_ptr):
        """Render into the host frame at `host_ptr`; zero-copy over every GPU's own PCIe link if it is pinned + mapped."""
        self._check(self._L.sbx_multi_render_host(self._m, C.byref(params), C.c_void_p(host_ptr)), "sbx_multi_render_host")

    def render_device(self, params):
        """Enqueue one frame into the group's frame on the first GPU; returns its device pointer (valid after sync(), or
        for work enqueued on stream())."""
        out = C.c_void_p()
        self._check(self._L.sbx_multi_render_device(self._m, C.byref(params), C.byref(out)), "sbx_multi_render_device")
        return out.value

    def stream(self):
        return self._L.sbx_multi_stream(self._m)

    def sync(self):
        self._check(self._L.sbx_multi_sync(self._m), "sbx_multi_sync")

    def timing(self, i):
        """sbx_last_timing of GPU i's context (launch shape, image, kernel time)."""
        t = Timing()
        self._check(self._L.sbx_last_timing(self._L.sbx_multi_ctx(self._m, int(i)), C.byref(t)), "sbx_last_timing")
        return {k: getattr(t, k) for k, _ in Timing._fields_}

    def kernel_ms(self):
        n = len(self.devices)
        arr = (C.c_float * n)()
        self._check(self._L.sbx_multi_last_timing(self._m, arr, n), "sbx_multi_last_timing")
        return [float(x) for x in arr]
