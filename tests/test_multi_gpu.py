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
        for stripe in (4, 1):
            if rank == 0:
                shared.tensor().fill_(float("nan"))
            torch.cuda.synchronize()
            dist.barrier()
            shared.render(p, stripe_rows=stripe)
            torch.cuda.synchronize()
            if rank == 0:
                np.save(os.path.join(out_dir, "p2p%d.npy" % stripe), shared.tensor().cpu().numpy())
            dist.barrier()
        shared.close()
        # path 3: every rank's kernel stores into ONE shared host frame over its own PCIe link
        shost = multi.SharedHostFrame(r, W, H)
        if rank == 0:
            shost.array[:] = np.nan
        dist.barrier()
        shost.render(p, stripe_rows=4)
        if rank == 0:
            np.save(os.path.join(out_dir, "host.npy"), np.array(shost.array))
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
    for name in ("nccl.npy", "p2p4.npy", "p2p1.npy", "host.npy"):
        assert bits_equal(np.load(str(tmp_path / name)), single), name
