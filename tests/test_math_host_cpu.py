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


def test_constant_division_identity_of_the_atmosphere_kernel(tmp_path):
    """x / 7994 and x / 1200 as one multiply and two FMAs (app_atmosphere_native.h): bit-equal to the division on every
    257th fp32 bit pattern inside 2^-100 <= |x| <= 2^100 here (all 2^32 patterns were run with stride 1: 0 mismatches in
    range; outside it only +-inf, -0 and results in the subnormal range differ), and the reciprocals in the kernel are
    RN(1/d)."""
    import json
    import re
    import subprocess

    exe = tmp_path / "div_const"
    subprocess.run(["gcc", "-O2", "-mfma", "-ffp-contract=off", os.path.join(ROOT, "tests", "native", "div_const.c"), "-o", str(exe), "-lm"], check=True)
    r = subprocess.run([str(exe), "257", "7994", "1200"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout
    out = json.loads(r.stdout)
    assert out["ok"] and all(d["mismatch_in_range"] == 0 and d["checked"] > 16_000_000 for d in out["divisors"])
    src = open(os.path.join(ROOT, "shaderbox_b200", "csrc", "native", "app_atmosphere_native.h")).read()
    for d in out["divisors"]:
        assert re.search(r"%.1ff, %sf\)" % (d["d"], re.escape(d["r"])), src), "kernel reciprocal for %g is not %s" % (d["d"], d["r"])
