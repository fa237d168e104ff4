"""CPU: the C-ABI library loads, exports what include/sbx.h declares, and fails loudly without a GPU."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

import shaderbox_b200 as sbx
from shaderbox_b200 import abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "sbx.h")


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(sbx_[a-z_0-9]+)\s*\(", text)))


def test_every_declared_symbol_is_exported():
    lib = sbx.lib()
    names = declared_symbols()
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), "libsbx.so does not export " + n
    assert set(names) == set(sbx.EXPORTS)


def test_no_oracle_or_torch_linked_into_product():
    out = subprocess.run(["ldd", sbx.LIB_PATH], capture_output=True, text=True).stdout
    assert "oracle" not in out and "torch" not in out and "libsbx_ref" not in out
    syms = subprocess.run(["nm", "-D", sbx.LIB_PATH], capture_output=True, text=True).stdout
    assert "sbxoracle" not in syms and "sbxref" not in syms


def test_default_params_match_reference_uniform_defaults():
    # src/uniform_buffer.h:41-58
    lib = sbx.lib()
    p = abi.Params()
    assert lib.sbx_default_params(C.byref(p), 1920, 1080) == 0
    q = abi.default_params(1920, 1080)
    assert bytes(p) == bytes(q)
    assert (p.width, p.height, p.cld_march_steps, p.illum_march_steps) == (1920, 1080, 100, 6)
    assert np.allclose(list(p.wind_dir), [0, 0, 0.2]) and np.allclose(list(p.sun_color), [1, 0.7, 0.55])
    assert np.isclose(p.sigma_scattering, 0.15) and np.isclose(p.cld_coverage, 0.535) and p.cld_thick == 125.0
    assert p.atm_radius == 5000.0 and p.atm_ground_y == 4750.0
    assert np.isclose(p.fog_density, 0.1) and p.fog_falloff == 0.5
    assert lib.sbx_default_params(None, 4, 4) == abi.SBX_ERR_INVALID
    assert lib.sbx_default_params(C.byref(p), 0, 4) == abi.SBX_ERR_INVALID


@pytest.mark.parametrize("height", [1, 3, 4, 7, 45, 1080, 2160])
def test_shard_rows_agree_with_python_twin(height):
    lib = sbx.lib()
    for stripe in (1, 2, 4, 7):
        for parts in (1, 2, 3, 8):
            total = 0
            for part in range(parts):
                sh = abi.Shard(stripe, parts, part)
                n = lib.sbx_shard_rows(C.byref(sh), height)
                assert n == len(abi.shard_rows(stripe, parts, part, height))
                total += n
            assert total == height
    bad = abi.Shard(4, 2, 2)
    assert lib.sbx_shard_rows(C.byref(bad), height) == abi.SBX_ERR_INVALID


def test_error_strings_and_version():
    lib = sbx.lib()
    assert lib.sbx_strerror(0) == b"ok"
    for code in range(-7, 0):
        assert lib.sbx_strerror(code) not in (b"", b"unknown status")
    assert b"sm_100a" in lib.sbx_version()


def test_invalid_arguments_are_rejected_without_a_context():
    lib = sbx.lib()
    p = abi.default_params(8, 8)
    buf = np.zeros((8, 8, 4), np.float32)
    assert lib.sbx_render_host(None, C.byref(p), None, buf.ctypes.data_as(C.c_void_p)) == abi.SBX_ERR_INVALID
    assert lib.sbx_render_device(None, C.byref(p), None, None, None) == abi.SBX_ERR_INVALID
    assert lib.sbx_load_app(None, b"APP_EGG", None) == abi.SBX_ERR_INVALID
    assert lib.sbx_create(0, None) == abi.SBX_ERR_INVALID


def test_new_entry_points_reject_bad_arguments_without_a_context():
    """rgba8 / sequence / shared host frame / volume bake: NULL context or buffers are SBX_ERR_INVALID, never a crash."""
    lib = sbx.lib()
    p = abi.default_params(8, 8)
    buf = np.zeros((8, 8, 4), np.float32)
    b8 = np.zeros((8, 8, 4), np.uint8)
    t = np.zeros(3, np.float32)
    vp = lambda a: a.ctypes.data_as(C.c_void_p)   # noqa: E731
    assert lib.sbx_render_host_rgba8(None, C.byref(p), None, vp(b8)) == abi.SBX_ERR_INVALID
    assert lib.sbx_render_device_rgba8(None, C.byref(p), None, None, None) == abi.SBX_ERR_INVALID
    assert lib.sbx_render_sequence_host(None, C.byref(p), None, vp(t), 3, vp(buf)) == abi.SBX_ERR_INVALID
    assert lib.sbx_render_sequence_device(None, C.byref(p), None, vp(t), 3, None, None) == abi.SBX_ERR_INVALID
    out = C.c_void_p()
    assert lib.sbx_host_frame_register(None, vp(buf), buf.nbytes, C.byref(out)) == abi.SBX_ERR_INVALID
    assert lib.sbx_host_frame_unregister(None, vp(buf)) == abi.SBX_ERR_INVALID
    assert lib.sbx_bake_noise_volume_host(None, 16, 0, 16, vp(buf)) == abi.SBX_ERR_INVALID
    assert lib.sbx_bake_noise_volume_device(None, 16, 0, 16, None, None) == abi.SBX_ERR_INVALID
    assert lib.sbx_set_option(None, b"coop_waves_x100", 1) == abi.SBX_ERR_INVALID


def test_compile_defines_reach_the_kernel_image(tmp_path, monkeypatch):
    """SBX_COMPILE_DEFINES: launch shape and out-of-line transcendentals are per-image build knobs (csrc/Makefile)."""
    hdr = tmp_path / "app_mini.h"
    hdr.write_text(MINI_APP)
    sizes = {}
    for tag, defs in (("inline", ""), ("outline", "SBX_MATH_OUTLINE=1;SBX_MIN_CTAS_PER_SM=8")):
        monkeypatch.setenv("SBX_COMPILE_DEFINES", defs)
        img = tmp_path / ("APP_MINI.%s.cubin" % tag)
        sbx.compile_app(str(hdr), "APP_MINI", str(img))
        res = subprocess.run(["cuobjdump", "-res-usage", str(img)], capture_output=True, text=True).stdout
        sizes[tag] = int(re.search(r"REG:(\d+)", res[res.index("sbx_render"):]).group(1))
    assert sizes["outline"] <= 64          # __launch_bounds__(128, 8) caps the registers at 65536 / (8 * 128)


def test_no_cpu_path_without_a_device():
    """On a machine without a B200 the product must refuse, not fall back."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("this box has a GPU")
    with pytest.raises(sbx.SbxError) as e:
        sbx.Renderer("APP_CLOUDS")
    assert e.value.status == abi.SBX_ERR_NO_DEVICE


MINI_APP = r"""
// a user-written app in the shaderbox plugin contract (src/main.h:3,35-50): camera, scene, render, FOV
#include "def.h"
#include "util.h"
#include "intersect.h"
#include "noise_iq.h"
#include "fbm.h"
DECL_FBM_FUNC(fbm, 3, noise_iq(p))
_mutable(float) seen = 0.;
void setup_camera(_inout(vec3) eye, _inout(vec3) look_at) { eye = vec3(0, 0, 3.); look_at = vec3(0, 0, 0); }
void setup_scene() { seen = 1.; }
vec3 render(_in(ray_t) ray, _in(vec3) point_cam)
{
    _constant(sphere_t) s = _begin(sphere_t) vec3(0, 0, 0), 1., 0 _end;
    hit_t hit = no_hit;
    intersect_sphere(ray, s, hit);
    if (hit.t >= max_dist) return vec3(.1, .2, .3) * seen;
    return vec3(fbm(hit.origin * 4. + u_time, 2., .5, .5));
}
#define FOV tan(radians(30.))
#include "main.h"
"""


def test_plugin_contract_compiles_without_a_gpu(tmp_path):
    """sbx_compile_app: an unchanged app header -> sm_100a cubin via NVRTC (no GPU needed)."""
    hdr = tmp_path / "app_mini.h"
    hdr.write_text(MINI_APP)
    img = tmp_path / "APP_MINI.plugin.cubin"
    sbx.compile_app(str(hdr), "APP_MINI", str(img))
    blob = img.read_bytes()
    assert blob[:4] == b"\x7fELF" and len(blob) > 4096
    dump = subprocess.run(["cuobjdump", "-elf", str(img)], capture_output=True, text=True).stdout
    assert "sm_100" in dump or "EF_CUDA_SM100" in dump
    assert "sbx_render" in dump


def test_compile_errors_are_reported_not_swallowed(tmp_path):
    hdr = tmp_path / "app_bad.h"
    hdr.write_text(MINI_APP.replace("return vec3(fbm(", "return vec3(no_such_function("))
    with pytest.raises(sbx.SbxError) as e:
        sbx.compile_app(str(hdr), "APP_BAD", str(tmp_path / "bad.cubin"))
    assert e.value.status == abi.SBX_ERR_COMPILE and "no_such_function" in str(e.value)


def test_prebuilt_kernel_images_are_sm100a_and_present():
    img_dir = os.path.join(ROOT, "shaderbox_b200", "images")
    for app in sbx.APPS:
        path = os.path.join(img_dir, app + ".plugin.cubin")
        assert os.path.exists(path), path + " missing: run __graft_entry__.build()"
    assert os.path.exists(os.path.join(img_dir, "sbx_util.cubin"))
    # the USE_NOISE_TEX branch of the unchanged app_clouds.h (HLSL texture declarations rewritten at compile time) and its
    # hand-written counterpart
    for name in ("APP_CLOUDS_TEX.plugin.cubin", "APP_CLOUDS_TEX.tma.cubin"):
        assert os.path.exists(os.path.join(img_dir, name)), name


def test_hlsl_register_bindings_are_rewritten_only_where_they_occur(tmp_path):
    """sbx_compile_app turns `Type name : register(tN);` into a bound member (include/sbx/hlsl_tex.h) and leaves every
    other ':' alone: a header with a ternary, a label-like comment and a binding compiles; the same header with an
    unknown resource type fails in the compiler, not in the rewrite."""
    ok = tmp_path / "app_tex_probe.h"
    ok.write_text('''#include "def.h"
#include "util.h"
Texture3D probe_tex : register(t2);
SamplerState probe_sampler : register(s0);   // register(t9) in a comment : stays
void setup_camera(_inout(vec3) eye, _inout(vec3) look_at) { eye = vec3(0, 0, 1); look_at = vec3(0, 0, 0); }
void setup_scene() {}
vec3 render(_in(ray_t) ray, _in(vec3) point_cam) {
    float s = probe_tex.SampleLevel(probe_sampler, ray.direction, 0).r;
    return point_cam.x > 0. ? vec3(s, s, s) : vec3(.5, .25, 0);
}
#define FOV 1.
#include "main.h"
''')
    os.environ["SBX_COMPILE_DEFINES"] = "USE_NOISE_TEX=1;SBX_USES_NOISE_TEX=1"
    try:
        sbx.compile_app(str(ok), "APP_TEX_PROBE", str(tmp_path / "probe.cubin"))
        assert os.path.getsize(tmp_path / "probe.cubin") > 1000
        bad = tmp_path / "app_tex_bad.h"
        bad.write_text(ok.read_text().replace("Texture3D probe_tex", "TextureCube probe_tex"))
        with pytest.raises(sbx.SbxError) as e:
            sbx.compile_app(str(bad), "APP_TEX_BAD", str(tmp_path / "bad.cubin"))
        assert e.value.status == abi.SBX_ERR_COMPILE and "TextureCube" in str(e.value)
    finally:
        del os.environ["SBX_COMPILE_DEFINES"]


def test_header_is_plain_c_and_links_from_c(tmp_path):
    """include/sbx.h compiles as strict C99 and a C program drives the device-free entry points through libsbx.so."""
    exe = tmp_path / "abi_c99"
    lib_dir = os.path.dirname(sbx.LIB_PATH)
    subprocess.run(["gcc", "-std=c99", "-pedantic", "-Wall", "-Wextra", "-Werror", os.path.join(ROOT, "tests", "native", "abi_c99.c"),
                    "-o", str(exe), "-L" + lib_dir, "-lsbx", "-Wl,-rpath," + lib_dir], check=True)
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert '"ok": true' in r.stdout


def test_example_host_of_the_integration_guide_builds(tmp_path):
    """examples/sbx_app.cpp is the C++ host INTEGRATION.md shows; it must compile and link against the shipped header + library
    (it needs a B200 to do anything: here it only has to refuse politely)."""
    exe = tmp_path / "sbx_app"
    lib_dir = os.path.dirname(sbx.LIB_PATH)
    subprocess.run(["g++", "-std=c++17", "-O1", "-Wall", "-Wextra", "-Werror", os.path.join(ROOT, "examples", "sbx_app.cpp"),
                    "-I" + os.path.join(ROOT, "include"), "-o", str(exe), "-L" + lib_dir, "-lsbx", "-Wl,-rpath," + lib_dir], check=True)
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 2 and "usage" in r.stderr
    import torch

    if not torch.cuda.is_available():
        r = subprocess.run([str(exe), "x.h", "APP_X", "8", "8", "1", str(tmp_path / "f%04d.ppm")], capture_output=True, text=True)
        assert r.returncode == 1 and "sbx_create" in r.stderr


@pytest.mark.parametrize("w,h", [(8, 4), (331, 203), (1920, 1080), (7, 3)])
def test_tile_parts_partition_every_frame(w, h):
    """abi.tile_part_mask (twin of the kernel's tile checkerboard): the parts are disjoint, cover the frame, and every
    part holds the same share of every tile row to within one tile."""
    for parts in (1, 2, 3, 8):
        total = np.zeros((h, w), int)
        for part in range(parts):
            m = abi.tile_part_mask(w, h, parts, part)
            total += m
            per_row = m.sum(axis=1)
            assert per_row.max() - per_row.min() <= 8 and abs(per_row.mean() - w / parts) <= 8
        assert (total == 1).all()


def _build_multi_c(tmp_path):
    exe = tmp_path / "multi_c"
    lib_dir = os.path.dirname(sbx.LIB_PATH)
    subprocess.run(["gcc", "-std=c99", "-pedantic", "-Wall", "-Wextra", "-Werror", os.path.join(ROOT, "tests", "native", "multi_c.c"),
                    "-o", str(exe), "-L" + lib_dir, "-lsbx", "-Wl,-rpath," + lib_dir], check=True)
    return exe


def test_c_host_of_the_multi_gpu_group_builds_and_refuses_without_a_device(tmp_path):
    """tests/native/multi_c.c (sbx_multi_* from plain C) compiles as strict C99; without a GPU it reports so and exits 3."""
    import torch

    exe = _build_multi_c(tmp_path)
    if not torch.cuda.is_available():
        r = subprocess.run([str(exe), "2"], capture_output=True, text=True)
        assert r.returncode == 3 and "no CUDA device" in r.stdout


@pytest.mark.gpu
@pytest.mark.parametrize("parts", [1, 2, 5, 8])
def test_c_host_renders_one_frame_over_several_parts(tmp_path, parts):
    """The C host: one process, `parts` parts (over the GPUs present, shared round-robin), frames byte-identical to 1 GPU."""
    exe = _build_multi_c(tmp_path)
    for app in ("APP_CLOUDS", "APP_PLANET"):
        r = subprocess.run([str(exe), str(parts), app], capture_output=True, text=True)
        assert r.returncode == 0 and '"ok": true' in r.stdout, r.stdout + r.stderr
