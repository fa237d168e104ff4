"""Row-stripe sharding of one frame over the ranks of a torch.distributed job (one process per GPU).

Pixels are independent (src/main.h:6-53 reads nothing but uniforms), so the frame is cut into stripes
of `stripe_rows` rows dealt round-robin to the ranks (include/sbx.h sbx_shard); every rank renders
its stripes compacted into one contiguous buffer, and ONE collective brings the parts to rank 0
(NCCL over NVLink when the buffers are CUDA tensors, gloo on CPU tensors for the host-logic tests).
Rank 0 then scatters every part to its frame rows (sbx_unshard_device on the GPU).

This module is plumbing: it never computes a pixel.
"""
import time

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


class _CudaArray:
    """Minimal __cuda_array_interface__ holder so a frame pointer owned by libsbx can be viewed as a torch tensor."""

    def __init__(self, ptr, shape):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": "<f4", "data": (int(ptr), False), "version": 2}


class SharedFrame:
    """The full frame on rank `dst`, mapped into every rank's address space with CUDA IPC
    (sbx_frame_alloc / sbx_frame_export / sbx_frame_import), so each rank's render kernel stores its
    stripes straight into it over NVLink: the gather is fused into the render kernel."""

    def __init__(self, renderer, width, height, dst=0, group=None):
        self.renderer, self.width, self.height, self.dst, self.group = renderer, int(width), int(height), dst, group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.nbytes = self.width * self.height * 16
        self.owner = self.rank == dst
        handle = [None]
        if self.owner:
            self.ptr = renderer.frame_alloc(self.nbytes)
            handle[0] = renderer.frame_export(self.ptr) if self.world > 1 else None
        if self.world > 1:
            dist.broadcast_object_list(handle, src=dst, group=group)
            if not self.owner:
                self.ptr = renderer.frame_import(handle[0])
            dist.barrier(group)
        dev = torch.device("cuda", renderer.device)
        self._done = torch.zeros(1, dtype=torch.float32, device=dev)

    def tensor(self):
        """[height, width, 4] float32 view of the frame (owner rank only)."""
        assert self.owner
        return torch.as_tensor(_CudaArray(self.ptr, (self.height, self.width, 4)), device=torch.device("cuda", self.renderer.device))

    def render(self, params, stripe_rows=DEFAULT_STRIPE_ROWS):
        """One frame: every rank launches its stripes into the shared frame, then a one-element all-reduce
        on the same stream orders "all stripes have landed" before anything rank `dst` enqueues next."""
        assert params.width == self.width and params.height == self.height
        dev = torch.device("cuda", self.renderer.device)
        stream = torch.cuda.current_stream(dev).cuda_stream
        self.renderer.render_frame(params, self.ptr, shard=shard_of(self.rank, self.world, stripe_rows), stream=stream)
        if self.world > 1:
            dist.all_reduce(self._done, group=self.group)

    def close(self):
        if self.world > 1:
            torch.cuda.synchronize()
            dist.barrier(self.group)
        if self.owner:
            self.renderer.frame_free(self.ptr)
        else:
            self.renderer.frame_release(self.ptr)
        self.ptr = None


class SharedHostFrame:
    """The full frame in HOST memory shared by the per-GPU processes of one box (a POSIX shared-memory file mapped by
    every rank, pinned and device-mapped by sbx_host_frame_register): each rank's render kernel stores its stripes
    straight into it over its own PCIe link.  This is the end-to-end path of an N-GPU host: no gather to one GPU, no
    device->host copy of the assembled frame.

    Completion is a HOST-side barrier, in this order: every rank synchronises its own stream (its stores have then
    landed in host memory), publishes the frame number in its slot of a control page behind the frame, and waits
    until every slot shows it.  (A GPU-side collective after the kernel would not do: nothing orders one GPU's
    PCIe writes to host memory before its NVLink traffic to another GPU.)"""

    def __init__(self, renderer, width, height, dst=0, group=None):
        import mmap
        import os

        import numpy as np

        self.renderer, self.width, self.height, self.dst, self.group = renderer, int(width), int(height), dst, group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.nbytes = self.width * self.height * 16
        page = mmap.PAGESIZE
        frame_bytes = (self.nbytes + page - 1) // page * page
        size = frame_bytes + page                               # + one control page: int64 slot per rank
        name = [None]
        if self.rank == dst:
            base = "/dev/shm"
            try:
                st = os.statvfs(base)
                if st.f_bavail * st.f_frsize < size + (16 << 20):
                    base = "/tmp"                               # a small /dev/shm (container default 64 MB): any file works
            except OSError:
                base = "/tmp"
            name[0] = "%s/sbx_frame_%d_%d" % (base, os.getpid(), id(self) & 0xffff)
            with open(name[0], "wb") as f:
                f.truncate(size)
        if self.world > 1:
            dist.broadcast_object_list(name, src=dst, group=group)
        self.path = name[0]
        self._file = open(self.path, "r+b")
        self._map = mmap.mmap(self._file.fileno(), size)
        self.array = np.frombuffer(self._map, dtype=np.float32, count=self.width * self.height * 4).reshape(self.height, self.width, 4)
        self._slots = np.frombuffer(self._map, dtype=np.int64, count=self.world, offset=frame_bytes)
        self._epoch = 0
        self.timeout_s = 120.0                                  # a rank that died must not leave the others spinning forever
        self.host_ptr = self.array.ctypes.data
        self.alias = renderer.host_frame_register(self.host_ptr, frame_bytes)
        if self.world > 1:
            dist.barrier(group)
        if self.rank == dst:
            os.unlink(self.path)          # every rank holds its mapping; the name is no longer needed

    def render(self, params, stripe_rows=DEFAULT_STRIPE_ROWS):
        """One frame, synchronous: on return (every rank) the frame is complete in host memory (self.array)."""
        assert params.width == self.width and params.height == self.height
        stream = torch.cuda.current_stream(torch.device("cuda", self.renderer.device)) if torch.cuda.is_available() else None
        self.renderer.render_frame(params, self.alias, shard=shard_of(self.rank, self.world, stripe_rows),
                                   stream=stream.cuda_stream if stream is not None else 0)
        if stream is not None:
            stream.synchronize()                                # this rank's stripes are in host memory
        t_wait = time.monotonic()
        self._epoch += 1
        self._slots[self.rank] = self._epoch
        slots, epoch = self._slots, self._epoch
        spins = 0
        while int(slots.min()) < epoch:                         # host barrier on the control page
            spins += 1
            if (spins & 0xfffff) == 0 and time.monotonic() - t_wait > self.timeout_s:
                raise RuntimeError("SharedHostFrame: rank(s) %s did not finish frame %d within %.0f s"
                                   % ([r for r in range(self.world) if int(slots[r]) < epoch], epoch, self.timeout_s))

    def close(self):
        if torch.cuda.is_available():
            torch.cuda.synchronize()
        if self.world > 1:
            dist.barrier(self.group)
        self.renderer.host_frame_unregister(self.host_ptr)
        self.array = None
        self._slots = None
        try:
            self._map.close()
        except BufferError:
            pass
        self._file.close()
