"""Frame-level data parallelism: one process per GPU, frames sharded contiguously, one collective.

The detection path has no cross-frame dependency (a frame is never split across GPUs), so the only
exchange step is the all-gather of per-frame keypoint counts from which every rank derives the global
CSR offsets of the batch result.  `torch.distributed` is plumbing: NCCL over
NVLink on the GPUs, gloo in the CPU tests.
"""
from __future__ import annotations

from typing import Tuple


def frame_shard(n_frames: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous block of frames owned by `rank`: [rank*F/G, (rank+1)*F/G)."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank / world size")
    return (rank * n_frames) // world, ((rank + 1) * n_frames) // world


def counts_from_offsets(offsets):
    """Per-frame keypoint counts from a rank-local CSR offsets tensor (F_local + 1,)."""
    return offsets[1:] - offsets[:-1]


def gather_frame_counts(local_counts, n_frames: int, group=None):
    """All-gathers per-frame counts (int64 tensor, this rank's frames in order) into the global (F,) tensor.

    Shards may differ by one frame, so each rank contributes a block padded to the largest shard.
    """
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    lo, hi = frame_shard(n_frames, rank, world)
    if local_counts.numel() != hi - lo:
        raise ValueError(f"rank {rank} owns {hi - lo} frames but passed {local_counts.numel()} counts")
    width = max(frame_shard(n_frames, r, world)[1] - frame_shard(n_frames, r, world)[0] for r in range(world))
    send = torch.zeros(width, dtype=torch.int64, device=local_counts.device)
    send[: hi - lo] = local_counts.to(torch.int64)
    recv = torch.empty(world * width, dtype=torch.int64, device=local_counts.device)
    dist.all_gather_into_tensor(recv, send, group=group)
    parts = []
    for r in range(world):
        a, b = frame_shard(n_frames, r, world)
        parts.append(recv[r * width: r * width + (b - a)])
    return torch.cat(parts)


def global_offsets(global_counts):
    """Exclusive scan: CSR offsets (F + 1,) of the batch result assembled from all ranks."""
    import torch

    out = torch.zeros(global_counts.numel() + 1, dtype=torch.int64, device=global_counts.device)
    torch.cumsum(global_counts, 0, out=out[1:])
    return out
