"""Generate tests/golden/ from the reference ITSELF (oracle/_ref: the verbatim /root/reference/src
headers compiled on oracle/ref/glsl_shim.h).  Run here (needs /root/reference for `make -C oracle ref`);
the fixtures travel to the GPU box, the reference does not.

    python tools/make_golden.py

frames.npz : per case  <key>/rgba float32 [h, w, 4], plus the uniforms in the key
ops.npz    : per operator <op>/in, <op>/out (sbx_eval_op layouts, tests/opcases.py inputs)
volume.npz : the util/ddsvolgen noise volume (16^3 whole, two 128^3 slices) and its DDS header bytes
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import opcases  # noqa: E402
from cases import FRAME_CASES, frame_key  # noqa: E402
from oracle import loader  # noqa: E402
from shaderbox_b200.abi import default_params  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
N_OPS = 256


def main():
    assert loader.have_ref(), "build oracle/_ref first: make -C oracle ref"
    os.makedirs(OUT, exist_ok=True)
    frames = {}
    for case in FRAME_CASES:
        app, w, h, t, ov = case
        frames[frame_key(case)] = loader.ref_render(app, default_params(w, h, t, **ov))
    np.savez_compressed(os.path.join(OUT, "frames.npz"), **frames)
    ops = {}
    for op in opcases.CASES:
        a, ow = opcases.inputs(op, N_OPS)
        ops[op + "/in"] = a
        ops[op + "/out"] = loader.ref_eval_op(op, a, ow)
    np.savez_compressed(os.path.join(OUT, "ops.npz"), **ops)
    # the ddsvolgen noise volume: a whole 16^3 bake, two slices of the shipped 128^3 size, and the DDS header bytes
    vol = {"v16": loader.ref_bake_volume(16), "v128_z0": loader.ref_bake_volume(128, 0, 1), "v128_z77": loader.ref_bake_volume(128, 77, 1),
           "hdr128": np.frombuffer(loader.ref_dds_header(128), dtype=np.uint8), "hdr16": np.frombuffer(loader.ref_dds_header(16), dtype=np.uint8)}
    np.savez_compressed(os.path.join(OUT, "volume.npz"), **vol)
    for f in ("frames.npz", "ops.npz", "volume.npz"):
        print(f, os.path.getsize(os.path.join(OUT, f)), "bytes")


if __name__ == "__main__":
    main()
