#!/usr/bin/env python
"""bench.py -- Mpixels/s of the shaderbox per-pixel path (BASELINE.json: APP_CLOUDS at 1920x1080).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one frame: every pixel's mainImage (src/main.h:6-53 -> src/app_clouds.h:204-218) for the
uniforms (u_res, u_time, aux block).  There is no input data besides those uniforms; the frame is the
output (RGBA32F, 16 B/pixel).

  value      frames stay in HBM: CUDA events around each step's launches on the launching stream
  e2e        the reference-facing call (sbx_render_host through the C ABI): uniforms from the host, the frame
             in host memory from sbx_host_alloc (pinned + mapped: the kernel stores into it over PCIe), wall
             time around the synchronous calls; e2e.pageable_value = the same call with a malloc'd frame
  roofline   mandated HBM figure (16 B/pixel written / kernel time / measured copy peak) -- and the
             FP32-issue figure that actually bounds this path (DESIGN.md "Roofline")
  cpu_baseline  oracle/_ref (the reference's own headers compiled for the host, AVX2+FMA fast-math
             build, all cores) on a bounded sample of rows of the same frame

  extra_workloads   BASELINE.json configs 3-5 (ATMOSPHERE 1080p, PLANET 4K, RAYTRACER 8K), same measurement
  frame      digest of the assembled frame (the same sample and 64-bit hash at every N; at N > 1 the
             frame is also compared bit for bit with rank 0's GPU rendering it alone, inside the run)

N > 1 (one process per GPU, torchrun): the frame is cut into 4-row stripes dealt round-robin to the
ranks (strong scaling; --split tiles = a checkerboard of warp tiles instead); every rank's render kernel stores its pixels straight into rank 0's frame
over NVLink and publishes a completion flag behind them; rank 0's stream waits on the flags -- all
inside the timed region (see shaderbox_b200/multi.py).  single_process_group = the same frame through
sbx_multi_* (one process driving all GPUs: what a C++ host calls).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # BASELINE.json configs[1]: the configuration the metric is quoted on
    "clouds1080": ("APP_CLOUDS", 1920, 1080, 1.5, {"cld_march_steps": 128}),
    "clouds1080_default100": ("APP_CLOUDS", 1920, 1080, 1.5, {}),
    # the same scene on the larger frames north_star names (1920x1080 -> 7680x4320)
    "clouds2160": ("APP_CLOUDS", 3840, 2160, 1.5, {"cld_march_steps": 128}),
    "clouds4320": ("APP_CLOUDS", 7680, 4320, 1.5, {"cld_march_steps": 128}),
    "atmosphere1080": ("APP_ATMOSPHERE", 1920, 1080, 1.0, {}),
    "planet2160": ("APP_PLANET", 3840, 2160, 2.0, {}),
    "raytracer4320": ("APP_RAYTRACER", 7680, 4320, 1.0, {}),
    "egg256": ("APP_EGG", 256, 256, 1.0, {}),
    "vinyl1080": ("APP_VINYL", 1920, 1080, 1.25, {}),
    "sdf_ao1080": ("APP_SDF_AO", 1920, 1080, 0.5, {}),
    # the USE_NOISE_TEX cloud path (two 128^3 noise textures, software D3D11 sampler, TMA-staged texel boxes)
    "clouds_tex1080": ("APP_CLOUDS_TEX", 1920, 1080, 1.5, {"cld_march_steps": 128}),
}
METRIC = "Mpixels/sec at 1920x1080 (APP_CLOUDS)"


def metric_name(workload):
    """BASELINE.json's metric for the default workload; the same quantity named by its own frame for the others."""
    app, w, h = WORKLOADS[workload][:3]
    return METRIC if workload == "clouds1080" else "Mpixels/sec at %dx%d (%s)" % (w, h, app)
# FP32 lane-instruction cost table for the algorithmic-work estimate (DESIGN.md "Roofline"):
# SASS instruction counts of the sbx_math.h routines and of the noise_iq body around its 8 hashes.
COST = {"sin": 40, "cos": 40, "exp": 26, "pow": 60, "sqrt": 8, "other": 60}


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return float(d["hbm_gbs"]), float(d.get("sm_max_mhz", 1965.0)), "measured (MEASURED_PEAKS.json)"
    return 6650.0, 1965.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region (B200_PROFILING.md): NVML polled
    every 5 ms from a thread (the nvidia-smi -lms recipe needs longer to start than a short run lasts)."""

    def __init__(self, device):
        self.device = device
        self.samples = []          # (t, sm_mhz, power_w, reasons_bitmask)
        self.stop_flag = False
        self.thread = None
        self.max_mhz = None
        self.marks = [None, None]

    def start(self):
        try:
            import pynvml

            pynvml.nvmlInit()
            # honour CUDA_VISIBLE_DEVICES-free boxes: index == CUDA ordinal on the gpurun boxes
            self.h = pynvml.nvmlDeviceGetHandleByIndex(self.device)
            self.nv = pynvml
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception as e:   # noqa: BLE001
            self.err = repr(e)
            self.nv = None
            return
        self.thread = threading.Thread(target=self._poll, daemon=True)
        self.thread.start()

    def _poll(self):
        nv = self.nv
        while not self.stop_flag:
            try:
                mhz = float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                pw = nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0
                try:
                    rs = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                except Exception:   # noqa: BLE001
                    rs = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                self.samples.append((time.perf_counter(), mhz, pw, rs))
            except Exception:   # noqa: BLE001
                pass
            time.sleep(0.005)

    def mark(self, which):
        self.marks[which] = time.perf_counter()

    def stop(self):
        self.stop_flag = True
        if self.thread:
            self.thread.join(timeout=2)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["no NVML samples: " + getattr(self, "err", "")]}
        t0, t1 = self.marks
        timed = [x for x in self.samples if t0 is not None and t1 is not None and t0 <= x[0] <= t1] or self.samples
        mhz = sorted(x[1] for x in timed)
        bits = 0
        for x in timed:
            bits |= x[3]
        names = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
                 0x80: "hw_power_brake_slowdown"}
        return {"sm_mhz": mhz[len(mhz) // 2], "sm_max_mhz": self.max_mhz, "power_w_max": max(x[2] for x in timed),
                "samples": len(timed), "reasons": sorted(v for k, v in names.items() if bits & k)}


# ---------------------------------------------------------------------------------------------------
# CPU arm: the reference's own implementation of the path on the host cores
# ---------------------------------------------------------------------------------------------------
def cpu_reference_rate(workload, budget_s=15.0, max_parts=None):
    """Mpix/s of oracle/_ref on all host cores over a bounded sample of rows of the workload frame."""
    from oracle import loader
    from shaderbox_b200.abi import Shard, default_params

    app, w, h, t, ov = WORKLOADS[workload]
    p = default_params(w, h, t, **ov)
    cores = os.cpu_count() or 1
    if loader.have_ref_fast():
        render, kind, build = loader.ref_fast_render, "reference", "oracle/_ref fast build (-O3 -ffast-math -mavx2 -mfma, src/Makefile:12-13 minus -march=native)"
    elif loader.have_ref():
        render, kind, build = loader.ref_render, "reference", "oracle/_ref strict build (-O2 -ffp-contract=off)"
    else:
        render, kind, build = loader.oracle_render, "port", "oracle/sbx_oracle.c (-O2 strict)"
    # calibrate on rows spread over the frame (every (h/8)-th row), then size the sample to the budget
    parts = max(1, h // 8)
    t0 = time.perf_counter()
    render(app, p, shard=Shard(1, parts, parts // 2), nthreads=cores)
    dt = time.perf_counter() - t0
    rows_cal = len(range(parts // 2, h, parts))
    per_row = dt / max(1, rows_cal)
    want_rows = int(max(rows_cal, min(h, budget_s / max(per_row, 1e-9))))
    parts2 = max(1, h // want_rows)
    sh = Shard(1, parts2, parts2 // 2)
    rows = len(range(parts2 // 2, h, parts2))
    t0 = time.perf_counter()
    render(app, p, shard=sh, nthreads=cores)
    dt = time.perf_counter() - t0
    return {"value": rows * w / dt * 1e-6, "unit": "Mpixels/s", "cores": cores, "kind": kind,
            "sample": "%d of %d rows (every %d-th) of the %dx%d frame, %.1f s, %s" % (rows, h, parts2, w, h, dt, build),
            "seconds": dt, "rows": rows, "parts": parts2}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    workload = args.workload
    app, w, h, t, ov = WORKLOADS[workload]
    # each step is one bounded sample; size it so steps+warmup finish within a few minutes
    total_budget = 120.0
    per_step = max(2.0, total_budget / max(1, args.steps + args.warmup))
    first = cpu_reference_rate(workload, budget_s=per_step)
    from oracle import loader
    from shaderbox_b200.abi import Shard, default_params

    render = loader.ref_fast_render if loader.have_ref_fast() else (loader.ref_render if loader.have_ref() else loader.oracle_render)
    p = default_params(w, h, t, **ov)
    sh = Shard(1, first["parts"], first["parts"] // 2)
    cores = first["cores"]
    for _ in range(max(0, args.warmup - 1)):
        render(app, p, shard=sh, nthreads=cores)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        render(app, p, shard=sh, nthreads=cores)
    dt = time.perf_counter() - t0
    value = first["rows"] * w * args.steps / dt * 1e-6
    line = {
        "impl": "reference", "metric": metric_name(args.workload), "value": value, "unit": "Mpixels/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "ms_per_full_frame_extrapolated": dt / args.steps * 1e3 * (h / first["rows"]),
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "%s %dx%d u_time=%g %s" % (app, w, h, t, json.dumps(ov)), "host": "CPU, %d threads" % cores,
                   "step": "one step = the bounded row sample named in cpu_baseline.sample (%d of %d rows); ms_per_step is measured, "
                           "ms_per_full_frame_extrapolated scales it to the frame" % (first["rows"], h)},
        "cpu_baseline": {"value": value, "unit": "Mpixels/s", "cores": cores, "kind": first["kind"], "sample": first["sample"]},
        "e2e": {"value": value, "unit": "Mpixels/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), file=args.json_out, flush=True)
    return 0


# ---------------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------------
EXTRA_WORKLOADS = ("atmosphere1080", "planet2160", "raytracer4320", "clouds_tex1080")   # BASELINE.json configs[2..4] + SURVEY f3


def frame_digest(frame_t):
    """(sample checksum, 64-bit position-weighted hash of the frame's bit patterns), computed on the GPU.  The sample
    is the one every line of every N prints (rows ::97, columns ::89, RGB), so lines can be compared across runs."""
    import torch

    sample = float(frame_t[::97, ::89, :3].double().sum())
    bits = frame_t.reshape(-1).view(torch.int32).to(torch.int64)
    acc = 0
    chunk = 1 << 26
    for o in range(0, bits.numel(), chunk):
        b = bits[o:o + chunk]
        wgt = (torch.arange(o, o + b.numel(), device=b.device, dtype=torch.int64) % 65521) + 1
        acc = (acc + int((b * wgt).sum())) & 0xFFFFFFFFFFFFFFFF
    return sample, "%016x" % acc


class Job:
    """One workload on this rank: the renderer, the frame it renders into, and the timed loops."""

    def __init__(self, workload, env, variant=None, stripe=4, split="rows", signal="flags"):
        import torch

        import shaderbox_b200 as sbx
        from shaderbox_b200 import multi
        from shaderbox_b200.abi import default_params

        self.env, self.workload, self.split, self.signal, self.stripe = env, workload, split, signal, stripe
        self.app, self.w, self.h, self.t, self.ov = WORKLOADS[workload]
        self.p = default_params(self.w, self.h, self.t, **self.ov)
        self.r = sbx.Renderer(self.app, device=env.local_rank, variant=variant)
        if self.app == "APP_CLOUDS_TEX":
            # texture 0: the ddsvolgen volume baked on this GPU; texture 1: the same volume mirrored in z (synthetic: the
            # reference ships no textures)
            import numpy as np

            vol = self.r.bake_noise_volume(128)
            self.r.set_noise_volumes(vol, np.ascontiguousarray(vol[::-1]))
        self.shared = multi.SharedFrame(self.r, self.w, self.h)          # world 1: a plain frame on this GPU
        self.stream = torch.cuda.current_stream(env.dev)
        self.px = self.w * self.h

    def launch(self):
        self.shared.launch(self.p, self.split, self.stripe, self.signal, stream=self.stream.cuda_stream)

    def step(self):
        self.launch()
        self.shared.complete(self.signal, stream=self.stream.cuda_stream)

    def time_device(self, steps, warmup):
        """`steps` frames, each: [L2 flush, barrier] untimed, then event | this rank's part | event | completion | event.
        Returns (total step ms: max over ranks, [per-rank mean kernel ms])."""
        import torch
        import torch.distributed as dist

        env = self.env
        self.r.set_option("record_events", 0)          # this loop brackets the launches with its own events
        for _ in range(max(3, warmup)):
            self.step()
            env.flush.zero_()
        env.barrier()
        ev = [tuple(torch.cuda.Event(enable_timing=True) for _ in range(3)) for _ in range(steps)]
        env.barrier()
        for k in range(steps):
            if env.world > 1:
                dist.barrier()                         # all ranks start the step together ...
            env.flush.zero_()                          # ... with the previous frame (and the memo table) evicted from L2; untimed
            ev[k][0].record(self.stream)
            self.launch()
            ev[k][1].record(self.stream)
            self.shared.complete(self.signal, stream=self.stream.cuda_stream)
            ev[k][2].record(self.stream)
        env.barrier()
        self.r.set_option("record_events", 1)
        kernel = sum(a.elapsed_time(b) for a, b, _ in ev)
        total = sum(a.elapsed_time(c) for a, _, c in ev)
        tot = torch.tensor([total], dtype=torch.float64, device=env.dev)
        per_rank = torch.zeros(env.world, dtype=torch.float64, device=env.dev)
        per_rank[env.rank] = kernel / steps
        if env.world > 1:
            dist.all_reduce(tot, op=dist.ReduceOp.MAX)
            dist.all_reduce(per_rank)
        return float(tot[0]), [float(x) for x in per_rank]

    def digest(self, check_against_single_gpu):
        """Digest of the assembled frame (rank 0); at N > 1 also compares it, bit for bit, with a render of the whole
        frame by rank 0's GPU alone."""
        import torch

        env = self.env
        self.step()
        torch.cuda.synchronize(env.dev)
        env.barrier()
        out = None
        if env.rank == 0:
            frame = self.shared.tensor()
            sample, h64 = frame_digest(frame)
            out = {"checksum": sample, "frame_hash": h64}
            if check_against_single_gpu:
                alone = torch.empty_like(frame)
                self.r.render_frame_part(self.p, alone.data_ptr(), stream=self.stream.cuda_stream)
                torch.cuda.synchronize(env.dev)
                same = bool(torch.equal(alone.view(torch.int32), frame.view(torch.int32)))
                out["equals_single_gpu_render"] = same
                if not same:
                    raise SystemExit("bench.py: the %d-GPU frame of %s differs from the 1-GPU frame" % (env.world, self.workload))
        env.barrier()
        return out

    def time_e2e(self, steps, pageable=False):
        """The reference-facing call with HOST frames, wall clock around synchronous calls, copies inside.
        N = 1: sbx_render_host into a frame from sbx_host_alloc (or a pageable numpy frame).
        N > 1: every rank's kernel stores its part into one host frame shared by the processes."""
        import ctypes as C

        import numpy as np
        import torch
        import torch.distributed as dist

        from shaderbox_b200 import multi

        env = self.env
        if env.world == 1:
            nbytes = self.px * 16
            if pageable:
                arr = np.empty((self.h, self.w, 4), np.float32)
                ptr, free = arr.ctypes.data, None
            else:
                ptr = self.r.host_alloc(nbytes)
                free = ptr
                arr = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_float)), shape=(self.h, self.w, 4))

            def call():
                self.r.render_host_ptr(self.p, ptr)
            closer = (lambda: self.r.host_free(free)) if free else (lambda: None)
        else:
            shost = multi.SharedHostFrame(self.r, self.w, self.h)
            arr = shost.array

            def call():
                shost.render(self.p, self.stripe, split=self.split)
            closer = shost.close
        for _ in range(2):
            call()
        env.barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            call()
        env.barrier()
        dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=env.dev)
        if env.world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        checksum = float(np.asarray(arr[::97, ::89, :3], dtype=np.float64).sum()) if env.rank == 0 else 0.0   # the host frame is read
        zero_copy = bool(self.r.timing()["zero_copy"]) if env.world == 1 else True
        closer()
        return float(dt[0]), checksum, zero_copy

    def close(self):
        self.shared.close()
        self.r.close()


class Env:
    def __init__(self):
        import torch
        import torch.distributed as dist

        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device -- shaderbox_b200 has no CPU path (use --impl reference for the CPU arm)")
        torch.cuda.set_device(self.local_rank)
        self.dev = torch.device("cuda", self.local_rank)
        self.host_group = None
        if self.world > 1:
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")   # stdout carries exactly one JSON line
            dist.init_process_group("nccl", device_id=self.dev)
            self.host_group = dist.new_group(backend="gloo")          # host-side waits that keep the GPUs idle
        self.flush = torch.empty(256 << 20, dtype=torch.uint8, device=self.dev)   # > 126 MB L2

    def barrier(self):
        import torch
        import torch.distributed as dist

        if self.world > 1:
            dist.barrier()
        torch.cuda.synchronize(self.dev)


def cubin_sha256(app, variant):
    import hashlib

    path = os.path.join(ROOT, "shaderbox_b200", "images", "%s.%s.cubin" % (app, variant))
    return hashlib.sha256(open(path, "rb").read()).hexdigest() if os.path.exists(path) else None


def extra_workload(env, name, steps, args):
    """One BASELINE.json config besides the metric's, same measurement, compact result."""
    job = Job(name, env, stripe=args.stripe_rows, split=args.split, signal=args.signal)
    total_ms, per_rank = job.time_device(steps, 3)
    dig = job.digest(env.world > 1)
    e2e_s, e2e_sum, _ = job.time_e2e(steps)
    tm = job.r.timing()
    out = None
    if env.rank == 0:
        hbm_peak, _, _ = peaks()
        kernel_ms = max(per_rank)
        bytes_rank0 = 16.0 * job.px / env.world
        out = {"workload": "%s %dx%d u_time=%g" % (job.app, job.w, job.h, job.t), "value": job.px * steps / (total_ms * 1e-3) * 1e-6,
               "unit": "Mpixels/s", "ms_per_step": total_ms / steps, "kernel_ms_per_rank": per_rank, "steps": steps,
               "e2e": job.px * steps / e2e_s * 1e-6, "e2e_checksum": e2e_sum,
               "hbm_frac": bytes_rank0 / (kernel_ms * 1e-3) * 1e-9 / hbm_peak, "regs": tm["regs_per_thread"], "ctas_per_sm": tm["blocks_per_sm"]}
        out.update(dig)
    job.close()
    return out


def run_ours(args):
    import torch
    import torch.distributed as dist

    env = Env()
    world, rank = env.world, env.rank
    if world == 1 and (args.format == "rgba8" or args.frames > 1):
        return run_ours_formats(args, env)
    if args.format == "rgba8" or args.frames > 1:
        raise SystemExit("bench.py: --format rgba8 / --frames are 1-GPU measurements")

    job = Job(args.workload, env, variant=args.variant, stripe=args.stripe_rows, split=args.split, signal=args.signal)
    sampler = ClockSampler(env.local_rank)   # runs through warm-up + the timed region (nvidia-smi needs ~0.3 s to start)
    if rank == 0:
        sampler.start()
    sampler.mark(0)
    wall0 = time.perf_counter()
    total_ms, per_rank_kernel = job.time_device(args.steps, args.warmup)
    wall = time.perf_counter() - wall0
    sampler.mark(1)
    clocks = sampler.stop() if rank == 0 else None
    tm = job.r.timing()
    dig = job.digest(world > 1)

    if args.no_zero_copy:
        job.r.set_option("host_zero_copy", 0)
    e2e_s, e2e_sum, zero_copy = job.time_e2e(args.steps)
    e2e_page = job.time_e2e(args.steps, pageable=True) if world == 1 else None

    # The side measurements must never cost the headline line: each runs under a watchdog; after the first failure (on
    # any rank: a collective may be left half-done) the rest are skipped and the process exits without teardown.
    extras = {}
    broken = None
    if not args.no_extra and args.workload == "clouds1080":
        for name in EXTRA_WORKLOADS:
            if broken:
                extras[name] = {"skipped": "after " + broken}
                continue
            try:
                with Watchdog(args.extra_timeout):
                    extras[name] = extra_workload(env, name, args.extra_steps, args)
            except BaseException as e:   # noqa: BLE001
                broken = "%s: %r" % (name, e)
                extras[name] = {"error": repr(e)}

    group_line = None
    if world > 1 and not args.no_group and not broken:
        try:
            with Watchdog(args.extra_timeout):
                group_line = single_process_group(env, job, args)
        except BaseException as e:   # noqa: BLE001
            broken = "single_process_group: %r" % (e,)
            group_line = {"error": repr(e)}

    if rank == 0:
        hbm_peak, sm_max_mhz, peak_src = peaks()
        px, w, h = job.px, job.w, job.h
        value = px * args.steps / (total_ms * 1e-3) * 1e-6
        avg_kernel_ms = max(per_rank_kernel)              # the dominant kernel's launch on the slowest rank
        alg_bytes = 16.0 * px / world                     # one rank's launch: 16 B/pixel (RGBA32F) written, 0 read
        achieved = alg_bytes / (avg_kernel_ms * 1e-3) * 1e-9
        traffic_db = {}
        tpath = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tpath):
            traffic_db = json.load(open(tpath))
        traffic = traffic_db.get(args.workload) if world == 1 else None
        split_text = {"tiles": "a checkerboard of 8x4 warp tiles dealt to the %d ranks (every rank renders 1/%d of every row)" % (world, world),
                      "rows": "%d-row stripes round-robin over %d ranks" % (args.stripe_rows, world)}[args.split]
        signal_text = {"flags": "each rank's stream stores a completion flag in rank 0's memory behind its launch (cuStreamWriteValue32); rank 0's stream "
                                "waits on the %d flags (cuStreamWaitValue32)" % world,
                       "kernel": "each launch's last thread block stores a completion flag behind its pixels in rank 0's memory; rank 0's stream waits "
                                 "on the %d flags (cuStreamWaitValue32)" % world,
                       "nccl": "a 1-element NCCL all-reduce on the launching stream as the completion barrier"}[args.signal]
        line = {
            "metric": metric_name(args.workload), "value": value, "unit": "Mpixels/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(3, args.warmup), "ms_per_step": total_ms / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "%s %dx%d u_time=%g %s" % (job.app, w, h, job.t, json.dumps(job.ov)), "variant": tm_variant(job.r, args),
                       "format": "RGBA32F", "frames_per_step": 1,
                       "l2": "flushed before every step (256 MiB memset on the same stream, after the per-step barrier, outside the timed events); the frame is write-only",
                       "sharding": "none" if world == 1 else (
                           "%s; every rank's render kernel stores its pixels into rank 0's frame over NVLink (CUDA-IPC peer mapping); %s; "
                           "all inside the step" % (split_text, signal_text)),
                       "grid": tm["grid_blocks"], "block": tm["block_threads"], "regs": tm["regs_per_thread"], "ctas_per_sm": tm["blocks_per_sm"],
                       "tail_rows": tm["tail_rows"], "tail_lanes_per_pixel": tm["tail_lanes_per_pixel"]},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                         "traffic": traffic, "peak_source": peak_src, "kernel": "sbx_render", "kernel_ms": avg_kernel_ms,
                         "algorithmic_bytes_per_launch": alg_bytes,
                         "note": "write-only path: 16 B/pixel out, 0 in; the binding roof is instruction issue (see issue_roofline)"},
            "kernel_ms_per_rank": per_rank_kernel,
            "e2e": {"value": px * args.steps / e2e_s * 1e-6, "unit": "Mpixels/s", "h2d_bytes_per_step": sbx_params_bytes(),
                    "d2h_bytes_per_step": int(16 * px), "d2h_bytes_per_step_this_rank": int(16 * px / world), "checksum": e2e_sum,
                    "api": ("sbx_render_host (C ABI) into a frame from sbx_host_alloc (pinned + mapped: what INTEGRATION.md tells a host to use)" if world == 1 else
                            "sbx_render_frame_part on every rank into one shared host frame (sbx_host_frame_register); each rank's stream publishes a "
                            "completion flag in the frame's control page behind its launch, every rank's host polls the %d flags" % world),
                    "d2h": "kernel stores straight into the pinned+mapped host frame (zero-copy over PCIe)" if zero_copy
                           else "frame assembled in HBM, then one async copy to host memory",
                    "loop": "back-to-back synchronous calls, wall clock; NO L2 flush and no per-step barrier (unlike `value`), so the two are not like for like"},
            "frame": dig,
            "gpu_launches": args.steps,
            "clocks": clocks,
            "wall_ms_per_step_incl_flush": wall / args.steps * 1e3,
            "kernel_only": {"ms": avg_kernel_ms, "mpix_s": px / (avg_kernel_ms * 1e-3) * 1e-6 if world == 1 else None},
        }
        if e2e_page is not None:
            line["e2e"]["pageable_value"] = px * args.steps / e2e_page[0] * 1e-6
            line["e2e"]["pageable_note"] = "same call with a malloc'd (pageable) frame: rendered in HBM, then cuMemcpyDtoHAsync"
        if extras:
            line["extra_workloads"] = extras
        if group_line:
            line["single_process_group"] = group_line
        if world == 1 and not args.no_cpu:
            cb = cpu_reference_rate(args.workload, budget_s=args.cpu_seconds)
            line["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
            line["reference_work"] = reference_work(args.workload, avg_kernel_ms, sm_max_mhz, clocks)
        winst = traffic_db.get(args.workload + "_warp_inst")
        if world == 1 and winst and args.variant in (None, "native"):
            # the roof that binds: issue slots.  Warp instructions per launch are a property of (kernel image, workload),
            # counted once by ncu (profiles/, smsp__inst_executed.sum); the time is this run's.  The count is only valid
            # for the image it was taken on: its SHA-256 is stored beside it.
            image = "hybrid" if tm["tail_rows"] > 0 else "native"
            sha_now, sha_then = cubin_sha256(job.app, image), traffic_db.get(args.workload + "_cubin_sha256")
            clk = (clocks or {}).get("sm_mhz") or sm_max_mhz
            slots = 148 * 4 * clk * 1e6 * avg_kernel_ms * 1e-3
            line["issue_roofline"] = {"bound": "issue", "warp_inst_per_launch": winst, "achieved": winst / (avg_kernel_ms * 1e-3) * 1e-9,
                                      "peak": 148 * 4 * clk * 1e-3, "unit": "G warp-inst/s", "frac": winst / slots,
                                      "stale": sha_now != sha_then, "image": image, "image_sha256": sha_now,
                                      "source": "ncu smsp__inst_executed.sum (profiles/traffic.json) / live kernel time; peak = 148 SMs x 4 schedulers x SM clock"}
        print(json.dumps(line), file=args.json_out, flush=True)
    if broken:
        sys.stderr.write("bench.py: a side measurement failed (%s); exiting without teardown\n" % broken)
        sys.stderr.flush()
        os._exit(0)
    job.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


class Watchdog:
    """SIGALRM after `seconds`: turns a hang in a side measurement (a peer that died inside a collective) into an exception."""

    def __init__(self, seconds):
        self.seconds = int(seconds)

    def __enter__(self):
        import signal

        def fire(signum, frame):
            raise TimeoutError("no progress within %d s" % self.seconds)
        self.old = signal.signal(signal.SIGALRM, fire)
        signal.alarm(self.seconds)
        return self

    def __exit__(self, *exc):
        import signal

        signal.alarm(0)
        signal.signal(signal.SIGALRM, self.old)
        return False


def single_process_group(env, job, args):
    """The same frame through sbx_multi_* (include/sbx.h): ONE process (rank 0's) drives every GPU of the run -- the call
    a C++ host makes.  The other ranks wait on the host (gloo), their GPUs idle but for rank 0's work."""
    import ctypes as C

    import numpy as np
    import torch
    import torch.distributed as dist

    import shaderbox_b200 as sbx

    out = None
    torch.cuda.synchronize(env.dev)
    dist.barrier(group=env.host_group)
    if env.rank == 0:
        try:
            m = sbx.MultiRenderer(job.app, n_gpus=env.world, variant=args.variant)
            p, px = job.p, job.px
            for _ in range(3):
                m.render_device(p)
            m.sync()
            t0 = time.perf_counter()
            for _ in range(args.steps):
                m.render_device(p)
            m.sync()
            dev_s = time.perf_counter() - t0
            kms = m.kernel_ms()
            host = m.host_alloc(px * 16)
            arr = np.ctypeslib.as_array(C.cast(host, C.POINTER(C.c_float)), shape=(job.h, job.w, 4))
            for _ in range(2):
                m.render_host_ptr(p, host)
            t0 = time.perf_counter()
            for _ in range(args.steps):
                m.render_host_ptr(p, host)
            host_s = time.perf_counter() - t0
            out = {"api": "sbx_multi_render_device / sbx_multi_render_host (one process, %d GPUs)" % env.world,
                   "value": px * args.steps / dev_s * 1e-6, "e2e": px * args.steps / host_s * 1e-6, "unit": "Mpixels/s",
                   "timing": "wall clock around back-to-back frames (pipelined on the device path), no L2 flush",
                   "kernel_ms_per_gpu": kms, "checksum": float(np.asarray(arr[::97, ::89, :3], dtype=np.float64).sum())}
            m.host_free(host)
            m.close()
        except Exception as e:   # noqa: BLE001 -- an extra, must not cost the headline line
            out = {"error": repr(e)}
    dist.barrier(group=env.host_group)
    return out


def run_ours_formats(args, env):
    """1-GPU side measurements: the 8-bit swap-chain format and time sequences in one launch."""
    import ctypes as _C

    import numpy as _np
    import torch

    import shaderbox_b200 as sbx
    from shaderbox_b200.abi import default_params

    app, w, h, t, ov = WORKLOADS[args.workload]
    p = default_params(w, h, t, **ov)
    r = sbx.Renderer(app, device=env.local_rank, variant=args.variant)
    rgba8 = args.format == "rgba8"
    nf = max(1, args.frames)
    seq_times = [t + k / 60.0 for k in range(nf)]             # a 60 Hz animation starting at the workload's u_time
    stream = torch.cuda.current_stream(env.dev)
    part = torch.empty((nf * h, w, 4), dtype=torch.uint8 if rgba8 else torch.float32, device=env.dev)
    px_bytes = 4 if rgba8 else 16

    def step():
        if nf > 1:
            r.render_sequence_into(p, seq_times, part.data_ptr(), stream=stream.cuda_stream)
        else:
            (r.render_rgba8_into if rgba8 else r.render_into)(p, part.data_ptr(), stream=stream.cuda_stream)

    for _ in range(max(3, args.warmup)):
        step()
        env.flush.zero_()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    for k in range(args.steps):
        env.flush.zero_()
        ev[k][0].record(stream)
        step()
        ev[k][1].record(stream)
    torch.cuda.synchronize()
    total_ms = sum(a.elapsed_time(b) for a, b in ev)
    host = torch.empty((nf * h, w, 4), dtype=torch.uint8 if rgba8 else torch.float32).pin_memory()
    _times = _np.asarray(seq_times, dtype=_np.float32)

    def e2e_step():
        if nf > 1:
            r._check(r._L.sbx_render_sequence_host(r._ctx, _C.byref(p), None, _times.ctypes.data_as(_C.c_void_p), nf,
                                                   _C.c_void_p(host.data_ptr())), "sbx_render_sequence_host")
        else:
            (r.render_rgba8_host_ptr if rgba8 else r.render_host_ptr)(p, host.data_ptr())
    for _ in range(2):
        e2e_step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    e2e_s = time.perf_counter() - t0
    px = w * h * nf
    hbm_peak, _, peak_src = peaks()
    kernel_ms = total_ms / args.steps
    line = {"metric": metric_name(args.workload), "value": px * args.steps / (total_ms * 1e-3) * 1e-6, "unit": "Mpixels/s", "n_gpus": 1, "steps": args.steps,
            "warmup": max(3, args.warmup), "ms_per_step": kernel_ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": "%s %dx%d u_time=%g %s" % (app, w, h, t, json.dumps(ov)), "format": "RGBA8_UNORM" if rgba8 else "RGBA32F",
                       "frames_per_step": nf, "l2": "flushed between steps"},
            "roofline": {"bound": "hbm", "achieved": px_bytes * px / (kernel_ms * 1e-3) * 1e-9, "peak": hbm_peak, "unit": "GB/s",
                         "frac": px_bytes * px / (kernel_ms * 1e-3) * 1e-9 / hbm_peak, "traffic": None, "peak_source": peak_src},
            "e2e": {"value": px * args.steps / e2e_s * 1e-6, "unit": "Mpixels/s", "h2d_bytes_per_step": sbx_params_bytes(),
                    "d2h_bytes_per_step": int(px_bytes * px), "checksum": float(host[::97, ::89, :3].double().sum())},
            "gpu_launches": args.steps}
    print(json.dumps(line), file=args.json_out, flush=True)
    r.close()
    return 0


def tm_variant(r, args):
    return args.variant or "default (native / hybrid if present, else plugin)"


def sbx_params_bytes():
    import ctypes

    from shaderbox_b200.abi import Params

    return ctypes.sizeof(Params)


def reference_work(workload, kernel_ms, sm_max_mhz, clocks):
    """How much of the REFERENCE's arithmetic a frame stands for: oracle call counters x a cost table (SASS instruction
    counts of sbx_math.h), as a lane-instruction rate.  Not a roofline: hash(n) calls the memo table serves still
    count as the reference's sin, so the rate exceeds the machine's issue peak -- it measures work avoided."""
    from oracle import loader
    from shaderbox_b200.abi import Shard, default_params

    app, w, h, t, ov = WORKLOADS[workload]
    if not loader.have_oracle():
        return None
    parts = max(1, h // 16)
    rows = len(range(parts // 2, h, parts))
    _, c = loader.oracle_render(app, default_params(w, h, t, **ov), shard=Shard(1, parts, parts // 2), want_counts=True)
    per_px = {k: v / float(rows * w) for k, v in c.items()}
    ops_px = sum(per_px[k] * COST[k] for k in COST)
    noise_px = per_px["sin"] / 8.0 * 74.0 if app in ("APP_CLOUDS", "APP_PLANET") else 0.0
    alg = (ops_px + noise_px) * w * h
    clk = (clocks or {}).get("sm_mhz") or sm_max_mhz
    peak = 148 * 128 * clk * 1e6
    return {"calls_per_pixel": per_px, "reference_lane_instr_per_frame": alg, "reference_lane_instr_per_s": alg / (kernel_ms * 1e-3),
            "machine_lane_instr_per_s": peak, "work_avoided_factor": alg / (kernel_ms * 1e-3) / peak, "clock_mhz_used": clk,
            "note": "call counts from the oracle on %d sampled rows; cost table in bench.py COST; NOT a fraction of peak" % rows}


def _stdout_for_json_only():
    """Libraries underneath (NCCL's version banner) write to fd 1; the contract is ONE JSON line on stdout.
    Point fd 1 at stderr for everything else and keep a private handle on the real stdout."""
    real = os.fdopen(os.dup(1), "w")
    sys.stdout.flush()
    os.dup2(2, 1)
    return real


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="clouds1080", choices=sorted(WORKLOADS))
    ap.add_argument("--variant", default=None, help="native | plugin (default: native if present)")
    ap.add_argument("--stripe-rows", type=int, default=4)
    ap.add_argument("--split", default="rows", choices=["tiles", "rows"],
                    help="N>1: how the frame is cut: a checkerboard of warp tiles (balanced by construction) or interleaved row stripes")
    ap.add_argument("--completion", dest="signal", default="flags", choices=["flags", "kernel", "nccl"],
                    help="N>1: completion = flags in rank 0's memory written by each rank's stream (cuStreamWriteValue32) or by its kernel's last "
                         "thread block, waited for with cuStreamWaitValue32 -- or a 1-element NCCL all-reduce")
    ap.add_argument("--no-extra", action="store_true", help="skip the extra_workloads (BASELINE configs 3-5)")
    ap.add_argument("--extra-steps", type=int, default=20)
    ap.add_argument("--extra-timeout", type=int, default=90, help="watchdog (s) per side measurement")
    ap.add_argument("--no-group", action="store_true", help="N>1: skip the single-process sbx_multi_* measurement")
    ap.add_argument("--format", default="f32", choices=["f32", "rgba8"],
                    help="frame format: raw float4 (BASELINE.json) or the 8-bit swap-chain format of the reference's presenting hosts")
    ap.add_argument("--frames", type=int, default=1, help="frames per step: a u_time sequence rendered by ONE launch (sbx_render_sequence_*)")
    ap.add_argument("--cpu-seconds", type=float, default=15.0)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-zero-copy", action="store_true", help="e2e: render in HBM and copy instead of storing into the host frame")
    args = ap.parse_args()
    args.json_out = _stdout_for_json_only()
    if args.impl == "reference":
        return run_reference_arm(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
