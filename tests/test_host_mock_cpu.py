"""Host logic of libsbx.so against a MOCK CUDA driver (tests/native/fake_cuda.c), in child processes.

No GPU is involved and nothing is rendered: the mock parses the real sm_100a kernel images (symbols, initialised globals,
register counts, launch bounds, parameter sizes), keeps "device" memory on the host and records every kernel launch.
What these tests pin down is the part of the hot path that runs on the CPU: which image a launch gets, the launch plan
(regions, grid, issue order, magic-number division) covering every pixel of its part exactly once under the kernel's
warp -> pixel mapping (sbx/sbx_launch.h), the 8-GPU single-process group and its completion flags, error paths without
leaks, and -- with ThreadSanitizer -- the worker-thread hand-off of sbx_multi.cpp.  The mock is built into a temporary
directory and reaches the child through LD_LIBRARY_PATH only; the product never loads it.
"""
import json
import os
import shutil
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NATIVE = os.path.join(ROOT, "tests", "native")
CSRC = os.path.join(ROOT, "shaderbox_b200", "csrc")
INCLUDES = ["-I/usr/local/cuda/include", "-I" + os.path.join(ROOT, "include"), "-I" + os.path.join(ROOT, "shaderbox_b200", "include")]

pytestmark = pytest.mark.skipif(not os.path.exists("/usr/local/cuda/include/cuda.h") or shutil.which("gcc") is None,
                                reason="needs gcc and the CUDA headers to build the mock driver")


def build_mock(where, extra=()):
    out = os.path.join(str(where), "libcuda.so.1")
    subprocess.run(["gcc", "-O1", "-g", "-std=gnu11", "-Wall", "-Wextra", "-Werror", "-fPIC", "-shared", "-Wl,-soname,libcuda.so.1", *extra, *INCLUDES,
                    os.path.join(NATIVE, "fake_cuda.c"), "-o", out, "-lpthread"], check=True)
    return out


@pytest.fixture(scope="module")
def mock_dir(tmp_path_factory):
    d = tmp_path_factory.mktemp("mock_driver")
    build_mock(d)
    return str(d)


def scenario(mock_dir, name, timeout=300, **env):
    e = dict(os.environ, LD_LIBRARY_PATH=mock_dir, **{k: str(v) for k, v in env.items()})
    e.pop("CUDA_VISIBLE_DEVICES", None)
    r = subprocess.run([sys.executable, os.path.join(NATIVE, "mock_scenarios.py"), name], env=e, capture_output=True, text=True, timeout=timeout)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    return json.loads(r.stdout.strip().splitlines()[-1])


def test_every_image_plans_launches_that_cover_each_pixel_exactly_once(mock_dir):
    """7 apps x their images x 7 frame sizes (1x1 .. 3840x2160) x 6 row partitions."""
    rep = scenario(mock_dir, "plan")
    assert rep["launches"] >= 400 and rep["live_allocs_after_close"] == 0
    clouds = rep["images"]["APP_CLOUDS.default"]
    # the image properties come out of the real cubins: the native CLOUDS kernel is capped at 80 registers, 6 CTAs per SM
    assert clouds[0]["regs"] <= 80 and clouds[0]["blocks_per_sm"] == 6 and clouds[0]["grid"] == 16200
    # the whole 1080p frame starts at the horizon (sky rows last): 25 % of 270 tile rows
    assert clouds[0]["first_tile_row"] == 67
    # explicit images keep their lanes per pixel for every partition
    assert {x["lanes_per_pixel"] for x in rep["images"]["APP_CLOUDS.coop"]} == {4}
    assert {x["lanes_per_pixel"] for x in rep["images"]["APP_CLOUDS.coop2"]} == {2}
    assert {x["lanes_per_pixel"] for x in rep["images"]["APP_CLOUDS.plugin"]} == {1}


def test_row_stripes_times_tile_checkerboards_partition_a_frame(mock_dir):
    rep = scenario(mock_dir, "frame_parts")
    assert rep["cases"] == 12 and rep["live_allocs_after_close"] == 0


def test_a_time_sequence_is_one_launch(mock_dir):
    rep = scenario(mock_dir, "sequence")
    assert rep["launches"] == 13 and rep["live_allocs_after_close"] == 0


def test_image_choice_follows_the_size_of_a_gpus_share(mock_dir):
    """DESIGN.md multi-GPU: a GPU's share of CLOUDS 1080p is marched with 1 / 1 / 2 / 4 lanes per pixel at N = 1 / 2 / 4 / 8."""
    rep = scenario(mock_dir, "image_choice")
    assert rep["lanes_per_pixel_by_parts"] == {"1": 1, "2": 1, "4": 2, "8": 4, "16": 4}


def test_single_process_group_on_8_mock_gpus(mock_dir):
    rep = scenario(mock_dir, "multi", SBX_FAKE_GPUS=8)
    assert rep["gpus"] == 8 and rep["stress_frames"] == 3000 and rep["live_allocs_after_close"] == 0
    for mode in ("device", "pinned_host", "pageable_host"):
        assert rep[mode]["launches"] == 8


def test_allocation_failures_are_errors_without_leaks(mock_dir):
    rep = scenario(mock_dir, "errors")
    assert sum(s != "ok" for s in rep["failed_at"]) >= 4, rep


def test_textured_images_take_the_descriptor_parameter(mock_dir):
    rep = scenario(mock_dir, "noise_tex")
    for variant in ("plugin", "tma"):
        assert rep[variant]["param_sizes"] == [280, 320]          # sbx_launch, sbx_tex_params (two 128-byte CUtensorMap + 32)
    assert rep["live_allocs_after_close"] == 0


def test_flags_on_pinned_memory(mock_dir):
    rep = scenario(mock_dir, "flags")
    assert rep["flags"] == list(range(41, 49)) and rep["bad_argument_status"][0] == -1 and rep["bad_argument_status"][2] == -1


def test_a_gpu_that_is_not_sm100_is_refused(mock_dir):
    rep = scenario(mock_dir, "wrong_device", SBX_FAKE_CC_MAJOR=9)
    assert rep["status"] == -2 and "sm_100a only" in rep["message"]


def test_group_needs_a_peer_path(mock_dir):
    rep = scenario(mock_dir, "no_peer", SBX_FAKE_GPUS=2, SBX_FAKE_NO_PEER=1)
    assert rep["status"] == -7 and rep["live_allocs"] == 0


def _sanitized_tree(tmp_path, flags, programs):
    """libsbx's sources, the mock driver and the given C hosts built with the sanitizer `flags` into tmp_path."""
    probe = tmp_path / "probe.c"
    probe.write_text("int main(void){return 0;}\n")
    if subprocess.run(["gcc", *flags, str(probe), "-o", str(tmp_path / "probe")], capture_output=True).returncode != 0:
        pytest.skip("this gcc has no runtime for " + " ".join(flags))
    d = str(tmp_path)
    build_mock(d, extra=flags)
    os.symlink(os.path.join(ROOT, "shaderbox_b200", "images"), os.path.join(d, "images"))      # the library finds its images beside itself
    os.symlink(os.path.join(ROOT, "shaderbox_b200", "include"), os.path.join(d, "include"))
    subprocess.run(["g++", *flags, "-O1", "-g", "-std=c++17", "-fPIC", "-shared", "-Wno-tsan", *INCLUDES,
                    *[os.path.join(CSRC, f) for f in ("sbx_host.cpp", "sbx_multi.cpp", "sbx_rtc.cpp")], "-o", os.path.join(d, "libsbx.so"), "-ldl", "-lpthread"],
                   check=True)
    for prog in programs:
        subprocess.run(["gcc", *flags, "-O1", "-g", "-std=gnu99", "-Wall", "-Wextra", "-Werror", os.path.join(NATIVE, prog + ".c"), "-o", os.path.join(d, prog),
                        "-L" + d, "-lsbx", "-Wl,-rpath," + d], check=True)
    return d


def test_worker_hand_off_under_thread_sanitizer(tmp_path):
    """libsbx's sources + the mock + tests/native/multi_stress.c, all built with -fsanitize=thread: 3000 frames over 8 mock
    GPUs (pinned, device and pageable destinations interleaved, workers put to sleep now and then) without a report."""
    d = _sanitized_tree(tmp_path, ["-fsanitize=thread"], ["multi_stress"])
    env = dict(os.environ, LD_LIBRARY_PATH=d, SBX_FAKE_GPUS="8", TSAN_OPTIONS="exitcode=66 halt_on_error=0")
    r = subprocess.run([os.path.join(d, "multi_stress"), "8", "3000"], env=env, capture_output=True, text=True, timeout=600, cwd=d)
    assert "ThreadSanitizer" not in r.stderr, r.stderr[-6000:]
    assert r.returncode == 0, r.stdout + r.stderr[-2000:]
    assert json.loads(r.stdout.strip().splitlines()[-1]) == {"ok": True, "gpus": 8, "frames": 3000}


def test_every_entry_point_under_address_and_ub_sanitizers(tmp_path):
    """tests/native/abi_tour.c calls the whole single-GPU ABI (all 7 apps x native / plugin images, every output flavour,
    parts + flags, host frames, IPC, options, noise volumes, the textured cloud images), multi_stress.c the group.  With
    the mock, device memory is host memory of exactly the requested size: a copy with a wrong size or offset, a leak or
    an undefined shift in the host library is a sanitizer report."""
    d = _sanitized_tree(tmp_path, ["-fsanitize=address,undefined", "-fno-omit-frame-pointer"], ["abi_tour", "multi_stress"])
    env = dict(os.environ, LD_LIBRARY_PATH=d, SBX_FAKE_GPUS="8", ASAN_OPTIONS="detect_leaks=1", UBSAN_OPTIONS="print_stacktrace=1")
    for cmd in (["abi_tour"], ["multi_stress", "8", "600"]):
        r = subprocess.run([os.path.join(d, cmd[0]), *cmd[1:]], env=env, capture_output=True, text=True, timeout=600, cwd=d)
        assert "Sanitizer" not in r.stderr and "runtime error" not in r.stderr, r.stderr[-6000:]
        assert r.returncode == 0, r.stdout + r.stderr[-2000:]
        assert json.loads(r.stdout.strip().splitlines()[-1])["ok"] is True


def test_cli_paths_under_the_mock(mock_dir, tmp_path):
    """sbx_cli (the C++ launcher of INTEGRATION.md): every sub-path runs to completion and writes files of the right size."""
    cli = os.path.join(ROOT, "shaderbox_b200", "sbx_cli")
    env = dict(os.environ, LD_LIBRARY_PATH=mock_dir, SBX_FAKE_GPUS="8")

    def run(*args):
        return subprocess.run([cli, *args], env=env, capture_output=True, text=True, timeout=120)

    out = str(tmp_path / "f.rgba32f")
    r = run("render", "APP_CLOUDS", "640", "360", "1.5", out, "--steps", "128", "--gpus", "8")
    assert r.returncode == 0, r.stderr
    line = json.loads(r.stdout)
    assert line["gpus"] == 8 and os.path.getsize(out) == 640 * 360 * 16
    r = run("render", "APP_EGG", "64", "36", "0", "-", "--ppm", str(tmp_path / "f.ppm"), "--rgba8", str(tmp_path / "f.rgba8"))
    assert r.returncode == 0 and os.path.getsize(tmp_path / "f.rgba8") == 64 * 36 * 4
    assert open(tmp_path / "f.ppm", "rb").read(15).startswith(b"P6\n64 36\n255\n")
    r = run("render", "APP_PLANET", "64", "36", "0", out, "--frames", "3", "--variant", "plugin")
    assert r.returncode == 0 and json.loads(r.stdout)["regs"] > 0
    r = run("bake", "16", str(tmp_path / "v.dds"))
    assert r.returncode == 0 and os.path.getsize(tmp_path / "v.dds") == 148 + 16 ** 3 * 16
    assert run("render", "APP_CLOUDS", "0", "36", "0", out).returncode == 2
    assert run("render", "APP_NOPE", "8", "8", "0", out).returncode == 1
    assert run("render", "APP_CLOUDS", "8", "8", "0", out, "--device", "9").returncode == 1


def test_a_failed_launch_on_one_gpu_of_the_group_is_reported_and_survivable(mock_dir):
    rep = scenario(mock_dir, "multi_launch_failure", SBX_FAKE_GPUS=8)
    assert "GPU part 0" in rep["0"] and "GPU part 5" in rep["5"] and rep["live_allocs_after_close"] == 0
