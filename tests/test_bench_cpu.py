"""CPU: the parts of bench.py that need no GPU -- the workload table against BASELINE.json, the frame digest, the
staleness bookkeeping of the ncu-derived instruction count, and the reference arm's contract under a multi-rank launch."""
import json
import os
import re
import subprocess
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def test_workloads_cover_the_baseline_configs():
    cfg = json.load(open(os.path.join(ROOT, "BASELINE.json")))["configs"]
    table = {(app, w, h) for app, w, h, _, _ in bench.WORKLOADS.values()}
    for line in cfg:
        m = re.match(r"(APP_[A-Z_]+) (\d+)x(\d+)", line)
        assert m and (m.group(1), int(m.group(2)), int(m.group(3))) in table, line
    assert bench.WORKLOADS["clouds1080"][4] == {"cld_march_steps": 128}           # "128 march steps"
    assert set(bench.EXTRA_WORKLOADS) >= {"atmosphere1080", "planet2160", "raytracer4320"}


def test_frame_digest_is_deterministic_and_sees_every_pixel():
    g = torch.Generator().manual_seed(3)
    frame = torch.rand((270, 480, 4), generator=g)
    s0, h0 = bench.frame_digest(frame)
    assert (s0, h0) == bench.frame_digest(frame.clone())
    moved = frame.clone()
    moved[100, 200, 1], moved[100, 201, 1] = frame[100, 201, 1], frame[100, 200, 1]    # swap two pixels off the sample grid
    s1, h1 = bench.frame_digest(moved)
    assert h1 != h0                                                                    # the hash is position-weighted
    one_ulp = frame.clone()
    one_ulp.view(torch.int32)[269, 479, 3] += 1
    assert bench.frame_digest(one_ulp)[1] != h0


def test_instruction_count_carries_the_hash_of_its_image():
    t = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
    assert re.fullmatch(r"[0-9a-f]{64}", t["clouds1080_cubin_sha256"])
    assert t["clouds1080_warp_inst"] > 10 ** 9 and t["clouds1080"] < 33177600           # DRAM traffic below the algorithmic bytes
    sha = bench.cubin_sha256("APP_CLOUDS", "native")
    assert sha is None or re.fullmatch(r"[0-9a-f]{64}", sha)


def test_reference_arm_prints_one_line_on_rank_0_only():
    """`bench.py --impl reference` under a multi-rank launch: rank 0 measures, the other ranks exit 0 without work."""
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, env=env, timeout=120)
    assert r.returncode == 0 and r.stdout.strip() == ""
    env = dict(os.environ, RANK="0", WORLD_SIZE="1", LOCAL_RANK="0")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0", "--workload", "egg256"],
                       capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0, r.stderr[-500:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["value"] > 0 and line["gpu_launches"] == 0
    assert line["cpu_baseline"]["kind"] in ("reference", "port") and line["cpu_baseline"]["cores"] >= 1
    assert line["ms_per_step"] > 0 and line["e2e"]["h2d_bytes_per_step"] == 0
