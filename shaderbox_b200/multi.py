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
    part straight into it over NVLink: the gather is fused into the render kernel.

    Behind the frame sit `world` 32-bit completion flags (same allocation, so the same IPC mapping): when rank r's
    launch has finished, its stream stores the frame number into flag r -- in rank `dst`'s memory, over the same
    NVLink path as its pixels (sbx_stream_write_flag; signal="kernel": the launch's last thread block stores it
    instead) -- and rank `dst`'s stream waits on the flags with cuStreamWaitValue32 (sbx_stream_wait_flags).  No
    collective and no host round trip sits between the last pixel and "frame complete".

    split = "rows": interleaved row stripes (include/sbx.h sbx_shard) -- the default: measured max/mean over 8 ranks
    1.01 on CLOUDS 1080p, and neighbouring warps stay neighbours in the frame (table lines shared in L1);
    split = "tiles": a checkerboard of 8x4 warp tiles dealt to the ranks (every rank renders the same share of EVERY
    row: balanced by construction, but 3-20 % slower per rank -- a rank's concurrent warps are spread 8 tiles apart)."""

    FLAG_BYTES = 4096

    def __init__(self, renderer, width, height, dst=0, group=None):
        self.renderer, self.width, self.height, self.dst, self.group = renderer, int(width), int(height), dst, group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.nbytes = self.width * self.height * 16
        self.flags_offset = (self.nbytes + 255) // 256 * 256
        self.owner = self.rank == dst
        self.epoch = 0
        handle = [None]
        if self.owner:
            self.ptr = renderer.frame_alloc(self.flags_offset + self.FLAG_BYTES)     # zero-filled: every flag starts at 0
            handle[0] = renderer.frame_export(self.ptr) if self.world > 1 else None
        if self.world > 1:
            dist.broadcast_object_list(handle, src=dst, group=group)
            if not self.owner:
                self.ptr = renderer.frame_import(handle[0])
            dist.barrier(group)
        self.flags_ptr = self.ptr + self.flags_offset
        dev = torch.device("cuda", renderer.device)
        self._done = torch.zeros(1, dtype=torch.float32, device=dev)

    def tensor(self):
        """[height, width, 4] float32 view of the frame (owner rank only)."""
        assert self.owner
        return torch.as_tensor(_CudaArray(self.ptr, (self.height, self.width, 4)), device=torch.device("cuda", self.renderer.device))

    def part(self, split="rows", stripe_rows=DEFAULT_STRIPE_ROWS):
        """Keyword arguments of Renderer.render_frame_part for this rank's part of the frame."""
        if split == "tiles":
            return {"shard": None, "tile_parts": self.world, "tile_part": self.rank}
        return {"shard": shard_of(self.rank, self.world, stripe_rows), "tile_parts": 1, "tile_part": 0}

    def launch(self, params, split="rows", stripe_rows=DEFAULT_STRIPE_ROWS, signal="flags", stream=None):
        """Enqueue this rank's part of one frame (asynchronous).  With signal="flags" the launch publishes its
        completion flag; complete() on the owner then orders "every part has landed" on its stream."""
        assert params.width == self.width and params.height == self.height
        if stream is None:
            stream = torch.cuda.current_stream(torch.device("cuda", self.renderer.device)).cuda_stream
        self.epoch += 1
        flag = self.flags_ptr + 4 * self.rank if self.world > 1 else 0
        self.renderer.render_frame_part(params, self.ptr, done_flag=flag if signal == "kernel" else 0, done_value=self.epoch,
                                        stream=stream, **self.part(split, stripe_rows))
        if signal == "flags" and flag:
            self.renderer.stream_write_flag(flag, self.epoch, stream=stream)

    def complete(self, signal="flags", stream=None):
        if self.world == 1:
            return
        if signal in ("flags", "kernel"):
            if self.owner:
                if stream is None:
                    stream = torch.cuda.current_stream(torch.device("cuda", self.renderer.device)).cuda_stream
                self.renderer.stream_wait_flags(self.flags_ptr, self.world, self.epoch, stream=stream)
        else:
            # a one-element all-reduce on the launching stream orders "every rank's kernel has finished"
            dist.all_reduce(self._done, group=self.group)

    def render(self, params, stripe_rows=DEFAULT_STRIPE_ROWS, split="rows", signal="flags"):
        """One frame: every rank launches its part into the shared frame; on the owner, work enqueued after this call
        on the current stream sees the complete frame.  (The other ranks do not wait: a caller that re-renders into
        the same frame before the owner has consumed it needs its own ordering -- a barrier, or two frames.)"""
        self.launch(params, split, stripe_rows, signal)
        self.complete(signal)

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
    """The full frame in HOST memory shared by the per-GPU processes of one box (an anonymous shared-memory file
    mapped by every rank, pinned and device-mapped by sbx_host_frame_register): each rank's render kernel stores its
    part straight into it over its own PCIe link.  This is the end-to-end path of an N-GPU host: no gather to one
    GPU, no device->host copy of the assembled frame.

    Completion: behind the frame sits a control page with one 32-bit slot per rank.  When rank r's launch has finished,
    its stream stores the frame number into slot r (sbx_stream_write_flag: cuStreamWriteValue32 behind a system-scope
    fence, travelling the same PCIe path as the pixels), and every rank's host thread polls the slots until all show
    the frame number -- no stream synchronise, no collective."""

    def __init__(self, renderer, width, height, dst=0, group=None):
        import mmap
        import os
        import tempfile

        import numpy as np

        self.renderer, self.width, self.height, self.dst, self.group = renderer, int(width), int(height), dst, group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.nbytes = self.width * self.height * 16
        page = mmap.PAGESIZE
        frame_bytes = (self.nbytes + page - 1) // page * page
        size = frame_bytes + page                               # + one control page: uint32 slot per rank
        name = [None]
        fd = unlink = None
        if self.rank == dst:
            # anonymous shmem (memfd: no name in any directory, independent of the size of /dev/shm, gone with its last
            # mapping); peers open it through /proc/<pid>/fd.  Fallback: an exclusive 0600 temp file in /dev/shm.
            try:
                fd = os.memfd_create("sbx_frame")
                os.ftruncate(fd, size)
                name[0] = "/proc/%d/fd/%d" % (os.getpid(), fd)
            except (AttributeError, OSError):
                fd, unlink = tempfile.mkstemp(prefix="sbx_frame_", dir="/dev/shm")   # O_EXCL, mode 0600
                os.ftruncate(fd, size)
                name[0] = unlink
        try:
            if self.world > 1:
                dist.broadcast_object_list(name, src=dst, group=group)
            if self.rank != dst:
                fd = os.open(name[0], os.O_RDWR)
            self._map = mmap.mmap(fd, size)
            if self.world > 1:
                dist.barrier(group)                             # every rank holds its mapping: the name can go
        finally:
            if unlink is not None:
                os.unlink(unlink)
            if fd is not None:
                os.close(fd)
        self.array = np.frombuffer(self._map, dtype=np.float32, count=self.width * self.height * 4).reshape(self.height, self.width, 4)
        self._slots = np.frombuffer(self._map, dtype=np.uint32, count=self.world, offset=frame_bytes)
        self._epoch = 0
        self.timeout_s = 20.0                                   # a rank that died must not leave the others spinning forever
        self.host_ptr = self.array.ctypes.data
        self.alias = renderer.host_frame_register(self.host_ptr, size)
        self.slots_alias = self.alias + frame_bytes
        if self.world > 1:
            dist.barrier(group)

    def render(self, params, stripe_rows=DEFAULT_STRIPE_ROWS, split="rows"):
        """One frame, synchronous: on return (every rank) the frame is complete in host memory (self.array)."""
        assert params.width == self.width and params.height == self.height
        stream = torch.cuda.current_stream(torch.device("cuda", self.renderer.device)).cuda_stream if torch.cuda.is_available() else 0
        self._epoch += 1
        part = ({"shard": None, "tile_parts": self.world, "tile_part": self.rank} if split == "tiles" else
                {"shard": shard_of(self.rank, self.world, stripe_rows), "tile_parts": 1, "tile_part": 0})
        self.renderer.render_frame_part(params, self.alias, stream=stream, **part)
        self.renderer.stream_write_flag(self.slots_alias + 4 * self.rank, self._epoch, stream=stream)
        slots, epoch = self._slots, self._epoch
        t_wait = time.monotonic()
        spins = 0
        while int(slots.min()) < epoch:                         # every rank's flag, written by its GPU
            spins += 1
            if (spins & 0xffff) == 0 and time.monotonic() - t_wait > self.timeout_s:
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
