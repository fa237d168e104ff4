"""CPU, world_size 2 over gloo: the N>1 host logic (stripe sharding, the single gather, reassembly).

The parts are produced by the CPU oracle here (the checker standing in for the renderer in a test);
on the GPU box the same functions move CUDA tensors over NCCL (test_gpu_parity.py, bench.py)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import loader
from shaderbox_b200 import multi
from shaderbox_b200.abi import Shard, default_params
from util import bits_equal

APP, W, H, T = "APP_PLANET", 40, 23, 2.0   # 23 rows: ragged against every stripe size used below


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, stripe, result_path):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        p = default_params(W, H, T)
        sh = multi.shard_of(rank, world, stripe)
        part = torch.from_numpy(loader.oracle_render(APP, p, shard=Shard(*sh), nthreads=1))
        assert part.shape[0] == multi.part_rows(H, sh)
        parts = multi.gather_parts(part, W, H, stripe, dst=0)
        if rank == 0:
            frame = multi.assemble_rows(parts, W, H, stripe)
            np.save(result_path, frame.numpy())
        else:
            assert parts is None
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,stripe", [(2, 4), (2, 1), (3, 5)])
def test_gathered_frame_is_bit_identical_to_single_rank(tmp_path, world, stripe):
    result = str(tmp_path / "frame.npy")
    mp.spawn(_worker, args=(world, _free_port(), stripe, result), nprocs=world, join=True)
    full = loader.oracle_render(APP, default_params(W, H, T))
    assert bits_equal(np.load(result), full)


def test_single_rank_is_identity():
    p = default_params(W, H, T)
    full = torch.from_numpy(loader.oracle_render(APP, p))
    assert multi.part_rows(H, multi.shard_of(0, 1)) == H
    out = multi.assemble_rows([full], W, H, 4)
    assert bits_equal(out.numpy(), full.numpy())


class _OracleHostRenderer:
    """Stands in for shaderbox_b200.Renderer in the shared-host-frame protocol test: "registering" the frame returns its
    own address and render_frame writes this rank's rows (from the CPU oracle) at their frame positions -- what the
    render kernel's zero-copy stores do on a GPU box."""

    device = 0

    def __init__(self, app):
        self.app = app

    def host_frame_register(self, host_ptr, nbytes):
        self.ptr = host_ptr
        return host_ptr

    def host_frame_unregister(self, host_ptr):
        assert host_ptr == self.ptr

    def render_frame_part(self, params, alias, shard=None, tile_parts=1, tile_part=0, done_flag=0, done_value=0, stream=0):
        import ctypes

        from shaderbox_b200.abi import shard_rows, tile_part_mask

        frame = np.ctypeslib.as_array(ctypes.cast(alias, ctypes.POINTER(ctypes.c_float)), shape=(params.height, params.width, 4))
        if tile_parts > 1:      # a checkerboard of warp tiles: this part's pixels of every row
            full = loader.oracle_render(self.app, params, nthreads=1)
            mask = tile_part_mask(params.width, params.height, tile_parts, tile_part)
            frame[mask] = full[mask]
        else:
            rows = shard_rows(shard[0], shard[1], shard[2], params.height)
            part = loader.oracle_render(self.app, params, shard=Shard(*shard), nthreads=1)
            if rows:
                frame[rows] = part
        if done_flag:           # what the launch's last thread block does after its pixel stores
            ctypes.c_uint.from_address(done_flag).value = done_value

    def stream_write_flag(self, flag, value, stream=0):   # what the stream does behind the launch
        import ctypes

        ctypes.c_uint.from_address(flag).value = value


def _host_frame_worker(rank, world, port, stripe, split, result_path):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        shost = multi.SharedHostFrame(_OracleHostRenderer(APP), W, H)
        for frame_no, t in enumerate((T, T + 1.0)):             # two frames: the barrier epoch advances
            if rank == 0:
                shost.array[:] = np.nan
            dist.barrier()
            shost.render(default_params(W, H, t), stripe, split=split)
            if rank == 0:
                np.save(result_path % frame_no, np.array(shost.array))
            dist.barrier()
        shost.close()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,stripe,split", [(2, 4, "rows"), (3, 1, "rows"), (2, 4, "tiles"), (3, 4, "tiles")])
def test_shared_host_frame_protocol(tmp_path, world, stripe, split):
    """One anonymous shared-memory frame mapped by every rank, each rank writing its own part (row stripes or a
    checkerboard of tiles), completion by the per-rank flags on the control page: on return from render() rank 0 reads
    the whole frame."""
    result = str(tmp_path / "host%d.npy")
    mp.spawn(_host_frame_worker, args=(world, _free_port(), stripe, split, result), nprocs=world, join=True)
    for frame_no, t in enumerate((T, T + 1.0)):
        assert bits_equal(np.load(result % frame_no), loader.oracle_render(APP, default_params(W, H, t)))
