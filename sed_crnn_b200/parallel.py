"""Multi-GPU plumbing: one process per GPU, torch.distributed (NCCL over NVLink on the B200 box, gloo in
the CPU tests).  The reference has no distributed code at all (train_lightning.py:46 pins devices=1); what
is added is exactly what the path needs:

  * log-mel extraction shards by clip with NO collective (clip i -> rank i mod world);
  * CRNN training is pure data parallel: every rank holds the full (0.4-2.3 M parameter) model, runs its
    slice of the global batch, and the flat fp32 gradient buffer is sum-all-reduced ONCE per step; the
    1/world scaling is folded into the fused clip+Adam kernel.  BatchNorm statistics stay per replica
    (DDP semantics; the reference has no SyncBN).
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def world_info(group=None) -> tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


def shard_range(n: int, rank: int, world: int) -> range:
    """Contiguous, balanced slice of range(n) for `rank` (first n % world ranks get one extra)."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world {world}")
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return range(lo, lo + base + (1 if rank < rem else 0))


def clip_ids_for_rank(n_clips: int, rank: int, world: int) -> range:
    """Round-robin clip assignment for bulk feature extraction (no collective needed)."""
    return range(rank, n_clips, world)


def batch_slice(global_batch: int, rank: int, world: int) -> slice:
    if global_batch % world:
        raise ValueError(f"global batch {global_batch} is not divisible by world size {world}")
    per = global_batch // world
    return slice(rank * per, (rank + 1) * per)


def allreduce_sum_(flat: torch.Tensor, group=None) -> float:
    """In-place sum all-reduce of a flat gradient buffer; returns the factor (1/world) the optimizer must
    apply.  One message per step: the buffer is 1.5-9 MB, i.e. latency-bound on NVLink 5, so there is no
    bucketing."""
    rank, world = world_info(group)
    if world > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    return 1.0 / world


def broadcast_(flat: torch.Tensor, src: int = 0, group=None) -> None:
    """Make every replica start from rank `src`'s parameters."""
    _, world = world_info(group)
    if world > 1:
        dist.broadcast(flat, src=src, group=group)
