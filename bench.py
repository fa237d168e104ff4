#!/usr/bin/env python
"""bench.py -- Mpixels/s of the shaderbox per-pixel path (BASELINE.json: APP_CLOUDS at 1920x1080).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one frame: every pixel's mainImage (src/main.h:6-53 -> src/app_clouds.h:204-218) for the
uniforms (u_res, u_time, aux block).  There is no input data besides those uniforms; the frame is the
output (RGBA32F, 16 B/pixel).

  value      frames stay in HBM: CUDA events around each step's launches on the launching stream
  e2e        the reference-facing call (sbx_render_host through the C ABI): uniforms from the host,
             frame copied back to pinned host memory, wall time around the synchronous calls
  roofline   mandated HBM figure (16 B/pixel written / kernel time / measured copy peak) -- and the
             FP32-issue figure that actually bounds this path (DESIGN.md "Roofline")
  cpu_baseline  oracle/_ref (the reference's own headers compiled for the host, AVX2+FMA fast-math
             build, all cores) on a bounded sample of rows of the same frame

N > 1: the frame is cut into 4-row stripes dealt round-robin to the ranks (strong scaling); every
step ends with the frame assembled on rank 0 (see shaderbox_b200/multi.py), inside the timed region.
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
    "atmosphere1080": ("APP_ATMOSPHERE", 1920, 1080, 1.0, {}),
    "planet2160": ("APP_PLANET", 3840, 2160, 2.0, {}),
    "raytracer4320": ("APP_RAYTRACER", 7680, 4320, 1.0, {}),
    "egg256": ("APP_EGG", 256, 256, 1.0, {}),
    "vinyl1080": ("APP_VINYL", 1920, 1080, 1.25, {}),
    "sdf_ao1080": ("APP_SDF_AO", 1920, 1080, 0.5, {}),
}
METRIC = "Mpixels/sec at 1920x1080 (APP_CLOUDS)"
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
        "impl": "reference", "metric": METRIC, "value": value, "unit": "Mpixels/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3 * (h / first["rows"]),
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "%s %dx%d u_time=%g %s" % (app, w, h, t, json.dumps(ov)), "host": "CPU, %d threads" % cores,
                   "ms_per_step_note": "extrapolated to the full frame from the row sample"},
        "cpu_baseline": {"value": value, "unit": "Mpixels/s", "cores": cores, "kind": first["kind"], "sample": first["sample"]},
        "e2e": {"value": value, "unit": "Mpixels/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), file=args.json_out, flush=True)
    return 0


# ---------------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist

    import shaderbox_b200 as sbx
    from shaderbox_b200 import multi
    from shaderbox_b200.abi import default_params

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- shaderbox_b200 has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")   # stdout carries exactly one JSON line
        dist.init_process_group("nccl", device_id=dev)

    app, w, h, t, ov = WORKLOADS[args.workload]
    p = default_params(w, h, t, **ov)
    r = sbx.Renderer(app, device=local_rank, variant=args.variant)
    stripe = args.stripe_rows
    shard = multi.shard_of(rank, world, stripe)
    rows = multi.part_rows(h, shard)
    stream = torch.cuda.current_stream(dev)
    rgba8 = args.format == "rgba8"
    if rgba8 and world > 1:
        raise SystemExit("bench.py: --format rgba8 is a 1-GPU measurement")
    nf = max(1, args.frames)                                  # frames per step (time sequence in one launch)
    if nf > 1 and (world > 1 or rgba8):
        raise SystemExit("bench.py: --frames is a 1-GPU float-frame measurement")
    seq_times = [t + k / 60.0 for k in range(nf)]             # a 60 Hz animation starting at the workload's u_time
    part = torch.empty((nf * rows, w, 4), dtype=torch.uint8 if rgba8 else torch.float32, device=dev)
    px_bytes = 4 if rgba8 else 16
    frame = torch.empty((h, w, 4), dtype=torch.float32, device=dev) if (world > 1 and rank == 0) else None
    fused = world > 1 and args.gather == "p2p"
    shared = multi.SharedFrame(r, w, h) if fused else None
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > 126 MB L2
    launches_per_step = 1 if (world == 1 or fused) else 1 + (world if rank == 0 else 0)

    def step():
        """kernel (+ gather + unshard at N > 1), all on `stream`"""
        if nf > 1:
            r.render_sequence_into(p, seq_times, part.data_ptr(), shard=shard, stream=stream.cuda_stream)
        elif world == 1:
            (r.render_rgba8_into if rgba8 else r.render_into)(p, part.data_ptr(), shard=shard, stream=stream.cuda_stream)
        elif fused:
            shared.render(p, stripe)
        else:
            multi.render_distributed(r, p, stripe, frame_out=frame, part_out=part)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    sampler = ClockSampler(local_rank)   # runs through warm-up + the timed region (nvidia-smi needs ~0.3 s to start)
    if rank == 0:
        sampler.start()
    for _ in range(max(3, args.warmup)):
        step()
        flush.zero_()
    barrier()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
          for _ in range(args.steps)]
    barrier()
    sampler.mark(0)
    wall0 = time.perf_counter()
    for k in range(args.steps):
        flush.zero_()                          # evict the previous frame from L2 (untimed)
        if world > 1:
            dist.barrier()                     # all ranks start the step together
        ev[k][0].record(stream)
        if world == 1:
            step()
            ev[k][1].record(stream)
        elif fused:
            r.render_frame(p, shared.ptr, shard=shard, stream=stream.cuda_stream)
            ev[k][1].record(stream)
            dist.all_reduce(shared._done)      # stream-ordered "every rank's stripes have landed in rank 0's frame"
        else:
            r.render_into(p, part.data_ptr(), shard=shard, stream=stream.cuda_stream)
            ev[k][1].record(stream)
            parts = multi.gather_parts(part, w, h, stripe, 0)
            if parts is not None:
                for q, prt in enumerate(parts):
                    if prt.shape[0]:
                        r.unshard(w, h, multi.shard_of(q, world, stripe), prt.data_ptr(), frame.data_ptr(), stream=stream.cuda_stream)
        ev[k][2].record(stream)
    barrier()
    wall = time.perf_counter() - wall0
    sampler.mark(1)
    kernel_ms = [a.elapsed_time(b) for a, b, _ in ev]
    step_ms = [a.elapsed_time(c) for a, _, c in ev]
    clocks = sampler.stop() if rank == 0 else None
    tm = r.timing()

    # max over ranks of the timed total
    tot = torch.tensor([sum(step_ms), sum(kernel_ms)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tot, op=dist.ReduceOp.MAX)
    total_ms, total_kernel_ms = float(tot[0]), float(tot[1])

    # ---- e2e: the reference-facing host call, host buffers, copies inside the timed region --------
    if args.no_zero_copy:
        r.set_option("host_zero_copy", 0)
    if world == 1:
        host = torch.empty((nf * rows, w, 4), dtype=torch.uint8 if rgba8 else torch.float32).pin_memory()
        import ctypes as _C
        import numpy as _np
        _times = _np.asarray(seq_times, dtype=_np.float32)

        def e2e_step():
            if nf > 1:                                                # sbx_render_sequence_host: synchronous
                r._check(r._L.sbx_render_sequence_host(r._ctx, _C.byref(p), _C.byref(sbx.Shard(*shard)), _times.ctypes.data_as(_C.c_void_p),
                                                       nf, _C.c_void_p(host.data_ptr())), "sbx_render_sequence_host")
            else:
                (r.render_rgba8_host_ptr if rgba8 else r.render_host_ptr)(p, host.data_ptr(), shard=shard)   # sbx_render_host[_rgba8]
        d2h_bytes = int(px_bytes * rows * w * nf)
    else:
        # N ranks: every rank's kernel stores its stripes into ONE host frame shared by the processes (POSIX shared memory,
        # pinned + mapped per GPU), each over its own PCIe link; complete after each rank's stream sync + a host barrier
        shost = multi.SharedHostFrame(r, w, h)
        host = torch.from_numpy(shost.array) if rank == 0 else None

        def e2e_step():
            shost.render(p, stripe)
        d2h_bytes = int(16 * rows * w)
    for _ in range(2):
        e2e_step()
    barrier()
    e0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    barrier()
    e_local = time.perf_counter() - e0
    e_t = torch.tensor([e_local], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e_t, op=dist.ReduceOp.MAX)
    e2e_s = float(e_t[0])
    checksum = float(host[::97, ::89, :3].double().sum()) if host is not None else 0.0   # the host result is actually read
    zero_copy = bool(r.timing()["zero_copy"]) if world == 1 else True   # N > 1: SharedHostFrame is mapped host memory by construction

    if rank == 0:
        hbm_peak, sm_max_mhz, peak_src = peaks()
        px = w * h * nf
        value = px * args.steps / (total_ms * 1e-3) * 1e-6
        avg_kernel_ms = total_kernel_ms / args.steps
        alg_bytes = float(px_bytes) * rows * w * nf       # this rank's launch: 16 (RGBA32F) or 4 (RGBA8) B/pixel written, 0 read
        achieved = alg_bytes / (avg_kernel_ms * 1e-3) * 1e-9
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tpath):
            traffic = json.load(open(tpath)).get(args.workload if world == 1 else "", None)
        line = {
            "metric": METRIC, "value": value, "unit": "Mpixels/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(3, args.warmup), "ms_per_step": total_ms / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "%s %dx%d u_time=%g %s" % (app, w, h, t, json.dumps(ov)), "variant": tm_variant(r, args), "format": "RGBA8_UNORM" if rgba8 else "RGBA32F", "frames_per_step": nf,
                       "l2": "flushed between steps (256 MiB memset on the same stream, outside the timed events); the frame is write-only",
                       "sharding": "none" if world == 1 else (
                           "%d-row stripes round-robin over %d ranks; every rank's render kernel stores its rows into rank 0's frame over NVLink "
                           "(CUDA-IPC peer mapping), then a 1-element all-reduce as the completion barrier, all inside the step" % (stripe, world)
                           if fused else
                           "%d-row stripes round-robin over %d ranks, NCCL gather to rank 0 + unshard kernels inside the step" % (stripe, world)),
                       "grid": tm["grid_blocks"], "block": tm["block_threads"], "regs": tm["regs_per_thread"], "ctas_per_sm": tm["blocks_per_sm"]},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                         "traffic": traffic, "peak_source": peak_src, "kernel": "sbx_render", "kernel_ms": avg_kernel_ms,
                         "algorithmic_bytes_per_launch": alg_bytes,
                         "note": "write-only path: 16 B/pixel out, 0 in; the binding roof is instruction issue (see issue_roofline)"},
            "e2e": {"value": px * args.steps / e2e_s * 1e-6, "unit": "Mpixels/s", "h2d_bytes_per_step": sbx_params_bytes(),
                    "d2h_bytes_per_step": d2h_bytes if world == 1 else int(16 * w * h), "d2h_bytes_per_step_this_rank": d2h_bytes, "checksum": checksum,
                    "api": ("sbx_render_host (C ABI) via shaderbox_b200.Renderer.render_host_ptr, pinned host frame" if world == 1 else
                            "sbx_render_frame on every rank into one shared host frame (sbx_host_frame_register); stream sync, then a host barrier on a shared control page"),
                    "d2h": "kernel stores straight into the pinned+mapped host frame (zero-copy over PCIe)" if zero_copy
                           else "frame assembled in HBM, then one async copy to pinned host memory"},
            "gpu_launches": launches_per_step * args.steps,
            "clocks": clocks,
            "wall_ms_per_step_incl_flush": wall / args.steps * 1e3,
            "kernel_only": {"ms": avg_kernel_ms, "mpix_s": px / (avg_kernel_ms * 1e-3) * 1e-6 if world == 1 else None},
        }
        if world == 1 and not args.no_cpu:
            cb = cpu_reference_rate(args.workload, budget_s=args.cpu_seconds)
            line["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
            line["fp32_issue"] = fp32_issue(args.workload, avg_kernel_ms, sm_max_mhz, clocks)
        if world == 1 and nf == 1 and not rgba8 and os.path.exists(tpath):
            # the roof that binds: issue slots.  Warp instructions per launch are a property of (kernel image, workload),
            # counted once by ncu (profiles/, smsp__inst_executed.sum); the time is this run's.
            winst = json.load(open(tpath)).get(args.workload + "_warp_inst")
            if winst and args.variant in (None, "native"):
                clk = (clocks or {}).get("sm_mhz") or sm_max_mhz
                slots = 148 * 4 * clk * 1e6 * avg_kernel_ms * 1e-3
                line["issue_roofline"] = {"bound": "issue", "warp_inst_per_launch": winst, "achieved": winst / (avg_kernel_ms * 1e-3) * 1e-9,
                                          "peak": 148 * 4 * clk * 1e-3, "unit": "G warp-inst/s", "frac": winst / slots,
                                          "source": "ncu smsp__inst_executed.sum (profiles/traffic.json) / live kernel time; peak = 148 SMs x 4 schedulers x SM clock"}
        print(json.dumps(line), file=args.json_out, flush=True)
    if world > 1:
        shost.close()
    if shared is not None:
        shared.close()
    r.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def tm_variant(r, args):
    return args.variant or "default (native if present, else plugin)"


def sbx_params_bytes():
    import ctypes

    from shaderbox_b200.abi import Params

    return ctypes.sizeof(Params)


def fp32_issue(workload, kernel_ms, sm_max_mhz, clocks):
    """Algorithmic FP32 lane-instructions (oracle call counters x COST) / kernel time / issue peak."""
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
    return {"calls_per_pixel": per_px, "alg_lane_instr_per_frame": alg, "achieved_lane_instr_per_s": alg / (kernel_ms * 1e-3),
            "peak_lane_instr_per_s": peak, "frac": alg / (kernel_ms * 1e-3) / peak, "clock_mhz_used": clk,
            "note": "call counts from the oracle on %d sampled rows; cost table in bench.py COST (SASS counts of sbx_math.h); "
                    "hash(n) calls served by the memo table still count as the reference's sin" % rows}


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
    ap.add_argument("--gather", default="p2p", choices=["p2p", "nccl"],
                    help="N>1: p2p = render kernels store into rank 0's frame over NVLink; nccl = compacted parts + ncclGather + unshard")
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
