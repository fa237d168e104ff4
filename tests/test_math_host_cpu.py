"""CPU: sbx_math.h (the transcendentals the kernels use) is bit-identical to the libm the reference's
C++ build links.  tests/native/math_vs_libm.c walks every `stride`-th fp32 bit pattern."""
import json
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def exe(tmp_path_factory):
    out = tmp_path_factory.mktemp("native") / "math_vs_libm"
    subprocess.run(["gcc", "-O2", "-std=c11", "-mfma", "-ffp-contract=off", "-fno-builtin",
                    os.path.join(ROOT, "tests", "native", "math_vs_libm.c"), "-o", str(out), "-lm", "-lpthread"],
                   check=True)
    return str(out)


@pytest.mark.parametrize("fn,stride", [("sinf", 257), ("cosf", 257), ("expf", 257), ("powf", 257), ("tanf", 257),
                                       ("acosf", 257), ("atanf", 257), ("atan2f", 257), ("sincosf", 257)])
def test_bit_identical_to_libm(exe, fn, stride):
    r = subprocess.run([exe, fn, str(stride), "8"], capture_output=True, text=True)
    rep = json.loads(r.stdout)
    assert rep["checked"] > 16_000_000
    assert rep["mismatch"] == 0, rep
