"""GPU, >= 2 B200s: the N-rank frame (both gather paths) is bit-identical to the 1-GPU frame."""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
APP, W, H, T, OV = "APP_CLOUDS", 480, 271, 1.5, {"cld_march_steps": 64}


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out_dir):
    import torch.distributed as dist

    import shaderbox_b200 as sbx
    from shaderbox_b200 import multi

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        r = sbx.Renderer(APP, device=rank)
        p = sbx.default_params(W, H, T, **OV)
        # path 1: compacted parts + one NCCL gather + unshard kernels
        frame = multi.render_distributed(r, p, stripe_rows=4)
        torch.cuda.synchronize()
        if rank == 0:
            np.save(os.path.join(out_dir, "nccl.npy"), frame.cpu().numpy())
        # path 2: every rank's kernel stores into rank 0's frame over NVLink (CUDA IPC mapping)
        shared = multi.SharedFrame(r, W, H)
        for name, kw in (("p2p4", {"split": "rows", "stripe_rows": 4, "signal": "nccl"}),
                         ("p2p1", {"split": "rows", "stripe_rows": 1, "signal": "kernel"}),
                         ("tiles", {"split": "tiles", "signal": "flags"}),
                         ("tiles2", {"split": "tiles", "signal": "flags"})):      # a second frame: the flags count frames
            if rank == 0:
                shared.tensor().fill_(float("nan"))
            torch.cuda.synchronize()
            dist.barrier()
            shared.render(p, **kw)
            if rank == 0:
                # NO host synchronisation between the render and the read: the copy is stream-ordered behind the
                # completion flags (or the all-reduce), which is exactly what is being tested
                np.save(os.path.join(out_dir, name + ".npy"), shared.tensor().cpu().numpy())
            torch.cuda.synchronize()
            dist.barrier()
        shared.close()
        # path 3: every rank's kernel stores into ONE shared host frame over its own PCIe link
        shost = multi.SharedHostFrame(r, W, H)
        if rank == 0:
            shost.array[:] = np.nan
        dist.barrier()
        for name, split in (("host", "tiles"), ("host_rows", "rows")):
            shost.render(p, stripe_rows=4, split=split)
            if rank == 0:
                np.save(os.path.join(out_dir, name + ".npy"), np.array(shost.array))
                shost.array[:] = np.nan
            dist.barrier()
        shost.close()
        r.close()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4, 8])
def test_multi_gpu_frame_equals_single_gpu_frame(tmp_path, world):
    if torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs" % world)
    import torch.multiprocessing as mp

    import shaderbox_b200 as sbx
    from util import bits_equal

    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    r = sbx.Renderer(APP, device=0)
    single = r.render(W, H, u_time=T, **OV)
    r.close()
    for name in ("nccl.npy", "p2p4.npy", "p2p1.npy", "tiles.npy", "tiles2.npy", "host.npy", "host_rows.npy"):
        assert bits_equal(np.load(str(tmp_path / name)), single), name


@pytest.mark.parametrize("world", [2, 4, 8])
def test_single_process_group_over_real_gpus(world):
    """sbx_multi_* with one part per physical GPU (peer stores over NVLink, per-GPU PCIe stores into a pinned frame)."""
    if torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs" % world)
    import ctypes as C

    import shaderbox_b200 as sbx
    from util import bits_equal

    r = sbx.Renderer(APP, device=0)
    single = r.render(W, H, u_time=T, **OV)
    r.close()
    m = sbx.MultiRenderer(APP, n_gpus=world)
    p = sbx.default_params(W, H, T, **OV)
    assert bits_equal(m.render(W, H, u_time=T, **OV), single)
    host = m.host_alloc(W * H * 16)
    arr = np.ctypeslib.as_array(C.cast(host, C.POINTER(C.c_float)), shape=(H, W, 4))
    for _ in range(3):
        arr[:] = np.nan
        m.render_host_ptr(p, host)
        assert bits_equal(np.array(arr), single)
    m.host_free(host)
    m.close()
