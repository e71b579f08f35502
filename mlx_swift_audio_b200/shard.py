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
