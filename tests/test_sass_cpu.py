"""CPU: what the built sm_100a kernel images contain (cuobjdump, no GPU needed) -- the SASS-level claims of DESIGN.md."""
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
IMAGES = os.path.join(ROOT, "shaderbox_b200", "images")
CUOBJDUMP = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"

pytestmark = pytest.mark.skipif(not os.path.exists(CUOBJDUMP), reason="cuobjdump not available")


def sass(image):
    out = subprocess.run([CUOBJDUMP, "-sass", os.path.join(IMAGES, image)], capture_output=True, text=True, check=True).stdout
    assert "sm_100a" in out
    return [re.sub(r"/\*.*?\*/", "", l).strip() for l in out.splitlines() if re.match(r"\s+/\*[0-9a-f]{4,5}\*/", l)]


def ops(lines):
    return [re.sub(r"^@!?U?P\w+\s+", "", l).split()[0].split(".")[0] for l in lines if l]


@pytest.mark.parametrize("image", ["APP_CLOUDS.native.cubin", "APP_CLOUDS.coop.cubin", "APP_CLOUDS.plugin.cubin", "APP_PLANET.native.cubin",
                                   "APP_RAYTRACER.native.cubin", "APP_EGG.plugin.cubin"])
def test_lut_block_is_staged_with_a_tma_bulk_copy(image):
    """cp.async.bulk global->shared completing on an mbarrier shows up as UBLKCP + SYNCS in SASS."""
    o = ops(sass(image))
    assert "UBLKCP" in o and "SYNCS" in o


def test_native_clouds_uses_packed_fp32_and_vector_loads_and_stays_in_registers():
    lines = sass("APP_CLOUDS.native.cubin")
    o = ops(lines)
    assert o.count("FFMA2") > 150                                       # the octave arithmetic, two lanes per instruction
    assert sum(1 for l in lines if l.startswith("LDG.E.128")) >= 16      # memo-table cells: two 16-byte loads per octave
    # no local-memory traffic inside the march loops (everything between the first and the last memo-table load);
    # ptxas may park a few per-pixel values (the sky colour) on the stack around the whole march
    loads = [i for i, l in enumerate(lines) if l.startswith("LDG.E.128")]
    first, last = loads[0], loads[15]                                    # view march (4 octaves) + light march (4 re-slices), 2 loads each
    assert not any(x in ("STL", "LDL") for x in o[first:last])
    assert sum(1 for x in o if x in ("STL", "LDL")) <= 24
    assert any(l.startswith("STG.E.EF.128") or l.startswith("STG.E.128") for l in lines)   # one float4 store per pixel
    assert len(lines) < 6000                                             # fits the instruction cache (PLANET lesson)


def test_cooperative_image_exchanges_steps_with_warp_shuffles():
    o = ops(sass("APP_CLOUDS.coop.cubin"))
    assert o.count("SHFL") >= 8 and "VOTE" in o                           # __shfl_sync of (T_i, A) per phase, __any/__all_sync


def test_native_raytracer_has_no_local_memory():
    o = ops(sass("APP_RAYTRACER.native.cubin"))
    assert not any(x in ("STL", "LDL") for x in o)
    plugin = ops(sass("APP_RAYTRACER.plugin.cubin"))
    assert plugin.count("STL") > 20                                      # the unchanged header keeps its scene tables in local memory


def test_native_planet_fits_the_instruction_cache():
    assert len(sass("APP_PLANET.native.cubin")) < 7000
    assert len(sass("APP_PLANET.plugin.cubin")) < 8000                    # 12 352 before fbm.h kept long octave loops rolled


def test_no_legacy_tensor_or_compat_paths():
    for image in os.listdir(IMAGES):
        o = set(ops(sass(image)))
        assert not (o & {"HMMA", "IMMA", "HGMMA"}), image                  # no dense contraction on this path: no tensor-core code


def test_noise_texture_image_stages_tiles_with_tma_tensor_loads():
    """APP_CLOUDS_TEX (USE_NOISE_TEX): the per-warp 4x4x4 texel boxes arrive by cp.async.bulk.tensor.3d (UTMALDG.3D) on the
    warp's mbarrier; the box the lanes share is found with warp reductions (CREDUX min/max)."""
    lines = sass("APP_CLOUDS_TEX.tma.cubin")
    o = ops(lines)
    assert sum(1 for l in lines if l.startswith("UTMALDG.3D")) >= 2       # one load per texture
    assert "SYNCS" in o and "CREDUX" in o and "LDS" in o


def test_every_image_prefetches_the_memo_table_into_l2():
    for image in ("APP_CLOUDS.native.cubin", "APP_CLOUDS.coop.cubin", "APP_PLANET.native.cubin"):
        assert "UBLKPF" in ops(sass(image)), image                        # cp.async.bulk.prefetch.L2
