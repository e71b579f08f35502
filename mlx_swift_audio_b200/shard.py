"""Data-parallel sharding of a batch of clips over the GPUs of one box.

Clips are independent (the only cross-frame state -- Whisper's max-8 clamp, Fun-ASR's CMVN, CAM++'s
mean-norm -- is per clip and stays on the clip's GPU), so every rank takes a contiguous block of
clips and there is NO collective on the data path.  ``gather_features`` is the optional step
for configurations whose consumer lives on one rank: NCCL on GPUs (NVLink 5 / NVSwitch), gloo in
the CPU tests.
"""
from __future__ import annotations


def shard_range(n_clips: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous block [start, stop) of clips owned by ``rank``: sizes differ by at most one and
    earlier ranks take the larger blocks."""
    if world <= 0 or not (0 <= rank < world) or n_clips < 0:
        raise ValueError("bad shard arguments")
    base, rem = divmod(n_clips, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def gather_features(local, n_clips: int, dst: int = 0, group=None):
    """Gathers per-rank feature blocks (rank r holds clips shard_range(n_clips, r, world), shape
    (n_local, ...)) onto rank ``dst`` in clip order.  Returns the full (n_clips, ...) tensor on ``dst`` and
    None elsewhere.  Uneven shards are handled by padding to the largest block."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    sizes = [shard_range(n_clips, r, world) for r in range(world)]
    max_n = max(b - a for a, b in sizes)
    pad = torch.zeros((max_n,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    bufs = [torch.empty_like(pad) for _ in range(world)] if rank == dst else None
    dist.gather(pad, bufs, dst=dst, group=group)
    if rank != dst:
        return None
    return torch.cat([bufs[r][: sizes[r][1] - sizes[r][0]] for r in range(world)], dim=0)


class FusedGather:
    """Gather fused into the producing kernels: rank ``dst`` owns ONE (n_clips, *per_clip_shape)
    float32 buffer, every rank maps it (CUDA IPC over NVLink 5 / NVSwitch peer memory) and passes its own slice as ``out=`` of
    the front end, so the kernel's epilogue stores ARE the transfer -- they overlap the FFT work tile by tile, there is no
    staging copy and no NCCL call on the data path.  torch.distributed only carries the 64-byte handle and the final barrier.

        ctx = api.Context(dev, torch.cuda.current_stream(dev).cuda_stream)      # (an own-stream Context also works: api orders it
        fg = FusedGather(ctx, n_clips, (3000, 128))                             #  against torch's current stream around every call)
        api.whisperLogMelSpectrogram(x_local, 128, ctx=ctx, out=fg.local_out())     # x_local: this rank's shard_range() clips
        full = fg.finish()          # rank dst: torch view of all clips' features; None elsewhere

    The buffer can be produced into again after ``reuse()`` (collective: the consumer calls it when it has finished reading);
    ``close()`` unmaps / frees it (collective: every rank calls it)."""

    def __init__(self, ctx, n_clips: int, per_clip_shape, dst: int = 0, group=None):
        import ctypes as C
        import torch.distributed as dist
        from . import _lib as L
        self.ctx, self.group, self.dst = ctx, group, dst
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.n_clips = int(n_clips)
        self.per_clip_shape = tuple(int(d) for d in per_clip_shape)
        self.clip_elems = 1
        for d in self.per_clip_shape:
            self.clip_elems *= d
        self.start, self.stop = shard_range(self.n_clips, self.rank, self.world)
        nbytes = max(1, self.n_clips * self.clip_elems * 4)
        lib = ctx.lib
        base = C.c_void_p()
        handle = [None]
        self._owner = self.rank == dst
        if self._owner:
            ctx.check(lib.b2a_device_alloc(ctx.h, C.byref(base), nbytes))
            buf = C.create_string_buffer(L.IPC_HANDLE_BYTES)
            ctx.check(lib.b2a_ipc_export(ctx.h, base, buf))
            handle[0] = buf.raw
        dist.broadcast_object_list(handle, src=dst, group=group)
        if not self._owner:
            ctx.check(lib.b2a_ipc_open(ctx.h, handle[0], C.byref(base)))
        self.base = int(base.value)
        self.nbytes = nbytes

    def local_out(self):
        """This rank's slice of the consumer's buffer, usable as ``out=`` of the api front ends."""
        from .api import DevicePtr
        return DevicePtr(self.base + self.start * self.clip_elems * 4, (self.stop - self.start,) + self.per_clip_shape)

    def finish(self):
        """Orders every producer's stores before the consumer's reads: stream sync on each rank, then a barrier.
        -> on ``dst`` a zero-copy torch view (n_clips, *per_clip_shape) of the gathered features, None elsewhere."""
        import torch.distributed as dist
        self.ctx.sync()
        dist.barrier(group=self.group)
        if not self._owner:
            return None
        return self._view(self.base, (self.n_clips,) + self.per_clip_shape)

    def _view(self, base: int, shape):
        """Zero-copy torch view of the consumer's buffer (overridden by the CPU test, whose "peer memory" is POSIX shared memory)."""
        import torch

        class _View:   # __cuda_array_interface__ v2: torch wraps the allocation without copying
            pass

        v = _View()
        v.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": "<f4", "data": (base, False), "version": 2, "strides": None}
        self._keep = v
        return torch.as_tensor(v, device=torch.device("cuda", self.ctx.device))

    def reuse(self):
        """Collective, before the buffer is produced into again: the consumer calls it once it is done reading the previous
        result (its reads are ordered on the context's stream), the producers before their next front-end call."""
        import torch.distributed as dist
        self.ctx.sync()
        dist.barrier(group=self.group)

    def close(self):
        import ctypes as C
        import torch.distributed as dist
        if self.base:
            self.ctx.sync()
            dist.barrier(group=self.group)   # nobody still writes / reads through a mapping that is about to go away
            if not self._owner:
                self.ctx.check(self.ctx.lib.b2a_ipc_close(self.ctx.h, C.c_void_p(self.base)))
            dist.barrier(group=self.group)
            if self._owner:
                self.ctx.check(self.ctx.lib.b2a_device_free(self.ctx.h, C.c_void_p(self.base)))
            self.base = 0
