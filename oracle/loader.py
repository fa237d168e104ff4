"""Loader for the CPU checkers (TEST INFRASTRUCTURE).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module.  The product package (shaderbox_b200) never does.

  oracle/_ref/libsbx_ref.so  -- the reference's own headers compiled on oracle/ref/glsl_shim.h
                                (built by `make -C oracle ref` where /root/reference exists)
  oracle/liboracle.so        -- the plain-C restatement oracle/sbx_oracle.c
"""
import ctypes as C
import os

import numpy as np

from shaderbox_b200.abi import Params, Shard, shard_rows

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SO = os.path.join(HERE, "_ref", "libsbx_ref.so")
REF_FAST_SO = os.path.join(HERE, "_ref", "libsbx_ref_fast.so")   # speed baseline only (fast-math), never parity
ORACLE_SO = os.path.join(HERE, "liboracle.so")

_cache = {}


def _load(path):
    if path not in _cache:
        _cache[path] = C.CDLL(path)
    return _cache[path]


def have_ref():
    return os.path.exists(REF_SO)


def have_ref_fast():
    return os.path.exists(REF_FAST_SO)


def have_oracle():
    return os.path.exists(ORACLE_SO)


def _render(lib, prefix, app, params, shard=None, nthreads=0, want_counts=False):
    name = app.lower().replace("app_", "")
    fn = getattr(lib, "%s_render_%s" % (prefix, name))
    fn.restype = C.c_int
    fn.argtypes = [C.POINTER(Params), C.POINTER(Shard), C.POINTER(C.c_float), C.c_int,
                   C.POINTER(C.c_ulonglong)]
    if shard is None:
        shard = Shard(1, 1, 0)
    rows = len(shard_rows(shard.stripe_rows, shard.n_parts, shard.part, params.height))
    out = np.empty((rows, params.width, 4), dtype=np.float32)
    counts = (C.c_ulonglong * 7)()
    if nthreads <= 0:
        nthreads = os.cpu_count() or 1
    rc = fn(C.byref(params), C.byref(shard), out.ctypes.data_as(C.POINTER(C.c_float)),
            int(nthreads), counts)
    if rc != 0:
        raise RuntimeError("%s_render_%s failed: %d" % (prefix, name, rc))
    if want_counts:
        keys = ("sin", "cos", "exp", "pow", "sqrt", "other")
        return out, dict(zip(keys, [int(c) for c in counts[:6]]))
    return out


def ref_render(app, params, shard=None, nthreads=0, want_counts=False):
    """The reference's own arithmetic (verbatim headers + shim + glibc libm)."""
    return _render(_load(REF_SO), "sbxref", app, params, shard, nthreads, want_counts)


def ref_fast_render(app, params, shard=None, nthreads=0, want_counts=False):
    """The reference's headers with the reference's own (fast-math) build flags: TIMING ONLY."""
    return _render(_load(REF_FAST_SO), "sbxref", app, params, shard, nthreads, False)


def oracle_render(app, params, shard=None, nthreads=0, want_counts=False):
    """The plain-C restatement."""
    return _render(_load(ORACLE_SO), "sbxoracle", app, params, shard, nthreads, want_counts)


def ref_eval_op(op, inputs, out_width):
    """One operator of the reference's own headers on the rows of `inputs` (oracle/ref/ref_ops.cpp);
    same op names and layouts as sbx_eval_op."""
    lib = _load(REF_SO)
    a = np.ascontiguousarray(inputs, dtype=np.float32)
    if a.ndim == 1:
        a = a[:, None]
    out = np.zeros((a.shape[0], out_width), dtype=np.float32)
    fn = lib.sbxref_eval_op
    fn.restype = C.c_int
    fn.argtypes = [C.c_char_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int]
    rc = fn(op.encode(), a.ctypes.data_as(C.c_void_p), a.shape[1], out.ctypes.data_as(C.c_void_p), out_width,
            a.shape[0])
    if rc != 0:
        raise RuntimeError("sbxref_eval_op(%s) failed: %d" % (op, rc))
    return out


def ref_bake_volume(size, z0=0, nz=None):
    """Slices [z0, z0+nz) of the ddsvolgen noise volume from the reference's own noise_worley.h / fbm.h."""
    nz = size - z0 if nz is None else nz
    lib = _load(REF_SO)
    out = np.empty((nz, size, size, 4), dtype=np.float32)
    lib.sbxref_bake_volume.restype = C.c_int
    lib.sbxref_bake_volume.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p]
    rc = lib.sbxref_bake_volume(int(size), int(z0), int(nz), out.ctypes.data_as(C.c_void_p))
    if rc != 0:
        raise RuntimeError("sbxref_bake_volume failed: %d" % rc)
    return out


def oracle_bake_volume(size, z0=0, nz=None):
    """The same slices from the plain-C restatement (oracle/sbx_oracle.c)."""
    nz = size - z0 if nz is None else nz
    lib = _load(ORACLE_SO)
    out = np.empty((nz, size, size, 4), dtype=np.float32)
    lib.sbxoracle_bake_volume.restype = C.c_int
    lib.sbxoracle_bake_volume.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p]
    rc = lib.sbxoracle_bake_volume(int(size), int(z0), int(nz), out.ctypes.data_as(C.c_void_p))
    if rc != 0:
        raise RuntimeError("sbxoracle_bake_volume failed: %d" % rc)
    return out


def ref_dds_header(size):
    """The bytes ddsvolgen writes in front of the volume, built from the reference's vendored DDS.h."""
    lib = _load(REF_SO)
    buf = C.create_string_buffer(256)
    lib.sbxref_dds_header.restype = C.c_int
    lib.sbxref_dds_header.argtypes = [C.c_int, C.c_void_p, C.c_int]
    n = lib.sbxref_dds_header(int(size), buf, 256)
    if n <= 0:
        raise RuntimeError("sbxref_dds_header failed: %d" % n)
    return buf.raw[:n]


def oracle_set_noise_volumes(vol_a, vol_b):
    """The two size^3 RGBA32F noise textures of the USE_NOISE_TEX cloud path (kept alive by the caller)."""
    lib = _load(ORACLE_SO)
    a = np.ascontiguousarray(vol_a, dtype=np.float32)
    b = np.ascontiguousarray(vol_b, dtype=np.float32)
    assert a.shape == b.shape and a.ndim == 4 and a.shape[0] == a.shape[1] == a.shape[2] and a.shape[3] == 4
    lib.sbxoracle_set_noise_volumes.restype = C.c_int
    lib.sbxoracle_set_noise_volumes.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
    rc = lib.sbxoracle_set_noise_volumes(a.ctypes.data_as(C.c_void_p), b.ctypes.data_as(C.c_void_p), int(a.shape[0]))
    if rc != 0:
        raise RuntimeError("sbxoracle_set_noise_volumes failed: %d" % rc)
    _cache["noise_volumes"] = (a, b)      # the C side keeps the pointers
    return a, b


def oracle_sample_noise(which, x, y, z):
    """One sample of noise texture `which` through the C restatement's sampler."""
    lib = _load(ORACLE_SO)
    lib.sbxoracle_sample_noise.restype = C.c_float
    lib.sbxoracle_sample_noise.argtypes = [C.c_int, C.c_float, C.c_float, C.c_float]
    return np.float32(lib.sbxoracle_sample_noise(int(which), float(x), float(y), float(z)))
