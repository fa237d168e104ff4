"""Row-stripe sharding of one frame over the ranks of a torch.distributed job (one process per GPU).

Pixels are independent (src/main.h:6-53 reads nothing but uniforms), so the frame is cut into stripes
of `stripe_rows` rows dealt round-robin to the ranks (include/sbx.h sbx_shard); every rank renders
its stripes compacted into one contiguous buffer, and ONE collective brings the parts to rank 0
(NCCL over NVLink when the buffers are CUDA tensors, gloo on CPU tensors for the host-logic tests).
Rank 0 then scatters every part to its frame rows (sbx_unshard_device on the GPU).

This module is plumbing: it never computes a pixel.
"""
import torch
import torch.distributed as dist

from .abi import shard_rows

DEFAULT_STRIPE_ROWS = 4   # one warp-tile row (SBX_TILE_H); keeps CLOUDS max/mean load <= 1.18


def shard_of(rank, world_size, stripe_rows=DEFAULT_STRIPE_ROWS):
    return (int(stripe_rows), int(world_size), int(rank))


def part_rows(height, shard):
    return len(shard_rows(shard[0], shard[1], shard[2], height))


def gather_parts(local_part, width, height, stripe_rows=DEFAULT_STRIPE_ROWS, dst=0, group=None):
    """Gather every rank's compacted part ([rows_r, width, 4] float32) to `dst` with one collective.

    Returns on dst a list of per-rank tensors (views of one receive buffer), elsewhere None.
    Parts may differ in row count by up to stripe_rows, so they are gathered padded to the largest."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    counts = [part_rows(height, shard_of(r, world, stripe_rows)) for r in range(world)]
    assert local_part.shape[0] == counts[rank] and local_part.shape[1] == width and local_part.shape[2] == 4
    if world == 1:
        return [local_part]
    max_rows = max(counts)
    send = local_part
    if counts[rank] != max_rows:
        send = torch.zeros((max_rows, width, 4), dtype=local_part.dtype, device=local_part.device)
        send[: counts[rank]] = local_part
    send = send.contiguous()
    if rank == dst:
        recv = torch.empty((world, max_rows, width, 4), dtype=local_part.dtype, device=local_part.device)
        dist.gather(send, list(recv.unbind(0)), dst=dst, group=group)
        return [recv[r, : counts[r]] for r in range(world)]
    dist.gather(send, None, dst=dst, group=group)
    return None


def assemble_rows(parts, width, height, stripe_rows=DEFAULT_STRIPE_ROWS, out=None):
    """Index-copy form of the unshard step (any device); the GPU path uses Renderer.unshard instead."""
    world = len(parts)
    if out is None:
        out = torch.empty((height, width, 4), dtype=parts[0].dtype, device=parts[0].device)
    for r, part in enumerate(parts):
        rows = shard_rows(stripe_rows, world, r, height)
        if rows:
            out[torch.as_tensor(rows, device=out.device)] = part
    return out


def render_distributed(renderer, params, stripe_rows=DEFAULT_STRIPE_ROWS, group=None, frame_out=None, part_out=None):
    """Every rank renders its stripes on its own GPU; rank 0 returns the full frame (CUDA tensor), others None.

    renderer: shaderbox_b200.Renderer bound to this rank's device."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    w, h = params.width, params.height
    shard = shard_of(rank, world, stripe_rows)
    rows = part_rows(h, shard)
    dev = torch.device("cuda", renderer.device)
    if part_out is None:
        part_out = torch.empty((rows, w, 4), dtype=torch.float32, device=dev)
    stream = torch.cuda.current_stream(dev).cuda_stream
    renderer.render_into(params, part_out.data_ptr(), shard=shard, stream=stream)
    if world == 1:
        return part_out   # one part, stripes in frame order already
    parts = gather_parts(part_out, w, h, stripe_rows, 0, group)
    if parts is None:
        return None
    if frame_out is None:
        frame_out = torch.empty((h, w, 4), dtype=torch.float32, device=dev)
    for r, part in enumerate(parts):
        if part.shape[0]:
            renderer.unshard(w, h, shard_of(r, world, stripe_rows), part.data_ptr(), frame_out.data_ptr(), stream=stream)
    return frame_out
