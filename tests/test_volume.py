"""The 3-D noise texture bake (util/ddsvolgen/src/ddsvolgen.cpp): header bytes on the CPU, voxels on the GPU."""
import os

import numpy as np
import pytest

import shaderbox_b200 as sbx
from oracle import loader
from util import bits_equal, diff_report

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def golden_volume():
    return np.load(os.path.join(ROOT, "tests", "golden", "volume.npz"))


def test_dds_header_bytes_match_the_reference_header(golden_volume):
    for size in (128, 16):
        got = sbx.dds_volume_header(size)
        assert len(got) == 148 and got == golden_volume["hdr%d" % size].tobytes()
        if loader.have_ref():
            assert got == loader.ref_dds_header(size)
    assert got[:4] == b"DDS " and got[84:88] == b"DX10"
    lib = sbx.lib()
    import ctypes as C

    buf = C.create_string_buffer(148)
    assert lib.sbx_dds_volume_header(128, buf, 147) < 0 and lib.sbx_dds_volume_header(0, buf, 148) < 0


@pytest.mark.skipif(not loader.have_ref(), reason="oracle/_ref not built")
def test_ref_volume_reproduces_golden(golden_volume):
    assert bits_equal(loader.ref_bake_volume(16), golden_volume["v16"])
    assert bits_equal(loader.ref_bake_volume(128, 77, 1), golden_volume["v128_z77"])


def test_c_oracle_volume_reproduces_golden(golden_volume):
    assert bits_equal(loader.oracle_bake_volume(16), golden_volume["v16"])
    assert bits_equal(loader.oracle_bake_volume(128, 77, 1), golden_volume["v128_z77"])
    assert bits_equal(loader.oracle_bake_volume(128, 0, 1), golden_volume["v128_z0"])


@pytest.mark.gpu
def test_baked_volume_matches_reference_golden(golden_volume):
    r = sbx.Renderer("APP_CLOUDS", device=0)
    try:
        got = r.bake_noise_volume(16)
        assert bits_equal(got, golden_volume["v16"]), diff_report(got, golden_volume["v16"])
        for z in (0, 77):
            got = r.bake_noise_volume(128, z, 1)
            want = golden_volume["v128_z%d" % z]
            assert bits_equal(got, want), diff_report(got, want)
        # slabs tile the volume; empty and out-of-range requests
        a = r.bake_noise_volume(16, 3, 5)
        assert bits_equal(a, golden_volume["v16"][3:8])
        assert r.bake_noise_volume(16, 16, 0).shape[0] == 0
        with pytest.raises(sbx.SbxError):
            r.bake_noise_volume(16, 10, 7)
    finally:
        r.close()


@pytest.mark.gpu
def test_full_size_volume_properties(golden_volume):
    """The shipped size (128^3, 32 MiB): only R is written, values stay in the range of 1 - (F1 + .25) summed with gains
    1, .5, .25, .125, the bake is deterministic, and sampled slices agree with the reference (computed here if present)."""
    import torch

    r = sbx.Renderer("APP_CLOUDS", device=0)
    try:
        dev = torch.empty((128, 128, 128, 4), dtype=torch.float32, device="cuda:0")
        s = torch.cuda.current_stream().cuda_stream
        r.bake_noise_volume_into(128, dev.data_ptr(), stream=s)
        torch.cuda.synchronize()
        ms = r.timing()["kernel_ms"]
        v = dev.cpu().numpy()
        assert (v[..., 1:] == 0).all() and np.isfinite(v[..., 0]).all()
        assert v[..., 0].min() > -1.875 * 0.75 - 1e-3 and v[..., 0].max() < 1.875 * 0.75 + 1e-3
        assert bits_equal(v[0:1], golden_volume["v128_z0"]) and bits_equal(v[77:78], golden_volume["v128_z77"])
        if loader.have_ref():
            assert bits_equal(v[120:121], loader.ref_bake_volume(128, 120, 1))
        dev2 = torch.empty_like(dev)
        r.bake_noise_volume_into(128, dev2.data_ptr(), stream=s)
        torch.cuda.synchronize()
        assert torch.equal(dev, dev2)
        print("bake 128^3: %.3f ms, %.1f Mvoxel/s" % (ms, 128 ** 3 / ms * 1e-3))
    finally:
        r.close()
