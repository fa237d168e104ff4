"""CPU: the launcher's host-side tools that need no GPU (sbx_cli flatten = util/inclxpnd/src/inclxpnd.cpp:8-41)."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "shaderbox_b200", "sbx_cli")
REF = os.path.join(ROOT, "oracle", "_ref", "inclxpnd_ref")

FILES = {
    "main.h": '#include "a.h"\nint main_body;\n  #include <b.h>   // indented, angle brackets\n#include "missing.h"\nlast line without newline',
    "a.h": "// a\n#include \"b.h\"\nfloat a;\n",
    "b.h": "float b;   \n\n#define X 1\n",
    "bad_short.h": "x\n#include \"\ny\n",
    "bad_quote.h": "x\n#include b.h\ny\n",
    "#include_not_first.h": "int x; #include \"b.h\"\n",
}
WANT_MAIN = "// a\nfloat b;   \n\n#define X 1\nfloat a;\nint main_body;\nfloat b;   \n\n#define X 1\n*** error: cannot include file: missing.h\nlast line without newline\n"


@pytest.fixture()
def tree(tmp_path):
    for name, text in FILES.items():
        (tmp_path / name).write_text(text)
    return tmp_path


def run(exe, cwd, name):
    r = subprocess.run([exe, name], cwd=cwd, capture_output=True)
    return r.returncode, r.stdout


def test_flatten_known_answer(tree):
    code, out = run_cli(tree, "main.h")
    assert code == 0 and out.decode() == WANT_MAIN
    assert run_cli(tree, "bad_short.h")[0] == 2          # inclxpnd.cpp:19
    assert run_cli(tree, "bad_quote.h")[0] == 3          # inclxpnd.cpp:21-25
    assert run_cli(tree, "nope.h")[0] == 1               # inclxpnd.cpp:53
    code, out = run_cli(tree, "#include_not_first.h")    # only a leading #include token expands
    assert code == 0 and out.decode() == FILES["#include_not_first.h"]


def run_cli(cwd, name):
    r = subprocess.run([CLI, "flatten", name], cwd=cwd, capture_output=True)
    return r.returncode, r.stdout


@pytest.mark.parametrize("name", sorted(FILES) + ["nope.h"])
def test_flatten_matches_the_reference_tool(tree, name):
    if not os.path.exists(REF):
        pytest.skip("oracle/_ref/inclxpnd_ref not built (needs /root/reference)")
    assert run_cli(tree, name) == run(REF, tree, name)


def test_flatten_reference_app_headers():
    src = "/root/reference/src"
    if not (os.path.exists(REF) and os.path.isdir(src)):
        pytest.skip("needs /root/reference and oracle/_ref/inclxpnd_ref")
    for app in ("app_clouds.h", "app_planet.h", "app_raytracer.h", "app_egg.h", "app_atmosphere.h"):
        assert run_cli(src, app) == run(REF, src, app), app
