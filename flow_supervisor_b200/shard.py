"""Batch sharding of image pairs across ranks (one process per GPU).

The correlation path has no cross-sample term: every image pair owns
its volume, so N GPUs simply take disjoint slices of the batch and no data-path collective
is needed.  Inference gathers per-rank results at the end (one all_gather of small
tensors); training uses DistributedDataParallel's gradient all-reduce, which is outside
this path.  The functions are backend-agnostic (NCCL on the GPU box, gloo in the CPU tests).
"""
from __future__ import annotations

from typing import Callable, List, Sequence, Tuple

import torch
import torch.distributed as dist


def partition(n_items: int, world_size: int) -> List[Tuple[int, int]]:
    """Contiguous, balanced [start, stop) ranges: the first n_items % world_size ranks get
    one extra item.  Every item is covered exactly once, ranks may be empty."""
    if n_items < 0 or world_size < 1:
        raise ValueError("partition needs n_items >= 0 and world_size >= 1")
    base, extra = divmod(n_items, world_size)
    out, start = [], 0
    for r in range(world_size):
        stop = start + base + (1 if r < extra else 0)
        out.append((start, stop))
        start = stop
    return out


def shard(batch: torch.Tensor, rank: int, world_size: int) -> torch.Tensor:
    """This rank's slice of a batch-major tensor."""
    lo, hi = partition(batch.shape[0], world_size)[rank]
    return batch[lo:hi]


def gather_batch(local: torch.Tensor, n_items: int, group=None) -> torch.Tensor:
    """Reassemble a batch-major result from every rank's slice (uneven slices allowed).
    The only communication of the inference path; called once per forward, not per lookup."""
    if not dist.is_available() or not dist.is_initialized():
        return local
    world = dist.get_world_size(group)
    parts = partition(n_items, world)
    width = max(hi - lo for lo, hi in parts)
    pad = local.new_zeros((width,) + tuple(local.shape[1:]))
    pad[: local.shape[0]] = local
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad, group=group)
    return torch.cat([b[: hi - lo] for b, (lo, hi) in zip(bufs, parts)], dim=0)


def run_sharded(fn: Callable[[torch.Tensor, torch.Tensor], torch.Tensor], first: torch.Tensor,
                second: torch.Tensor, group=None) -> torch.Tensor:
    """Apply ``fn(first_slice, second_slice)`` (e.g. a RAFT forward using CorrBlock) to this
    rank's share of a batch of pairs and return the full batch of results on every rank."""
    if not dist.is_available() or not dist.is_initialized():
        return fn(first, second)
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    out = fn(shard(first, rank, world), shard(second, rank, world))
    return gather_batch(out, first.shape[0], group)
