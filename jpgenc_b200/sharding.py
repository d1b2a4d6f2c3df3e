"""Sharding of a batch of independent images over ranks.

A single image is never split (the reference emits no restart markers, src/Image.cpp:950-971); a batch shards by
image with NO data-path collective.  The only cross-rank step is bookkeeping: every rank learns the encoded size of
every frame so that (offset, size) in a concatenated output is known everywhere.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple


def frames_for_rank(n_frames: int, rank: int, world: int) -> range:
    """Contiguous, balanced partition: the first (n % world) ranks take one extra frame."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError("bad rank/world")
    base, extra = divmod(n_frames, world)
    start = rank * base + min(rank, extra)
    return range(start, start + base + (1 if rank < extra else 0))


def owner_of(frame: int, n_frames: int, world: int) -> int:
    base, extra = divmod(n_frames, world)
    edge = extra * (base + 1)
    return frame // (base + 1) if frame < edge else extra + (frame - edge) // max(base, 1)


def offsets_from_sizes(sizes: Sequence[int]) -> List[Tuple[int, int]]:
    out, at = [], 0
    for s in sizes:
        out.append((at, int(s)))
        at += int(s)
    return out


def gather_frame_sizes(local_sizes: Sequence[int], n_frames: int, device=None) -> List[int]:
    """all-gather the per-frame encoded sizes (frame order) with torch.distributed; works on gloo and nccl."""
    import torch
    import torch.distributed as dist

    world, rank = dist.get_world_size(), dist.get_rank()
    mine = frames_for_rank(n_frames, rank, world)
    assert len(local_sizes) == len(mine)
    cap = (n_frames + world - 1) // world
    buf = torch.zeros(cap, dtype=torch.int64, device=device)
    if len(mine):
        buf[: len(mine)] = torch.tensor(list(local_sizes), dtype=torch.int64, device=device)
    parts = [torch.zeros_like(buf) for _ in range(world)]
    dist.all_gather(parts, buf)
    sizes: List[int] = []
    for r in range(world):
        sizes.extend(int(v) for v in parts[r][: len(frames_for_rank(n_frames, r, world))].tolist())
    return sizes
