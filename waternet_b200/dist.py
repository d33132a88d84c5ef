"""Batch sharding across GPUs: one process per GPU, images are independent units.

Every image is independent in the preprocess (per-image statistics) and in the
forward (no cross-sample op), so the path shards by batch with no data-path
collective; the only exchange is ONE all-gather of the output tensor (SURVEY.md
section 8e).  The weights (4.36 MB) are replicated.
"""
from __future__ import annotations

from typing import Callable, List, Optional, Tuple

import torch
import torch.distributed as dist


def shard_range(n: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous split of ``n`` images: returns ``(start, count)`` for ``rank``."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError(f"bad world/rank {world}/{rank}")
    base, extra = divmod(n, world)
    count = base + (1 if rank < extra else 0)
    start = rank * base + min(rank, extra)
    return start, count


def shard_counts(n: int, world: int) -> List[int]:
    return [shard_range(n, world, r)[1] for r in range(world)]


def all_gather_batch(local: torch.Tensor, counts: List[int], group=None) -> torch.Tensor:
    """All-gather per-rank batches of possibly different length along dim 0."""
    world = dist.get_world_size(group)
    if len(counts) != world:
        raise ValueError("counts must have one entry per rank")
    cap = max(counts)
    if local.shape[0] != counts[dist.get_rank(group)]:
        raise ValueError("local batch does not match its shard size")
    if cap == 0:
        return local
    padded = local
    if local.shape[0] != cap:
        pad = torch.zeros((cap - local.shape[0],) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        padded = torch.cat([local, pad], dim=0)
    padded = padded.contiguous()
    gathered = torch.empty((world * cap,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(gathered, padded, group=group)
    if all(c == cap for c in counts):
        return gathered
    return torch.cat([gathered[r * cap:r * cap + c] for r, c in enumerate(counts)], dim=0)


class PassGather:
    """All-gather of every rank's output batch, one pass of images at a time.

    ``Enhancer.submit(..., on_pass=gather.on_pass)`` calls :meth:`on_pass` on the copy-out stream as soon as a
    pass of the local batch is finished, so the collective of pass k runs under the kernels of pass k+1 (NCCL
    over NVLink; the only exchange of the path).  Every rank contributes the same batch size
    (weak scaling); ``gathered[r]`` is rank r's batch, ``result()`` the (world*B, ...) concatenation in rank order --
    bitwise what one process computing all world*B images returns.
    """

    def __init__(self, local_shape, dtype, device, group=None):
        self.group = group
        self.world = dist.get_world_size(group)
        self.gathered = torch.empty((self.world,) + tuple(local_shape), dtype=dtype, device=device)
        self.calls = 0

    def on_pass(self, local_chunk: torch.Tensor, a: int, b: int) -> None:
        outs = [self.gathered[r, a:b] for r in range(self.world)]  # each a contiguous block of rank r's batch
        dist.all_gather(outs, local_chunk.contiguous(), group=self.group)
        self.calls += 1

    def result(self) -> torch.Tensor:
        return self.gathered.view((-1,) + tuple(self.gathered.shape[2:]))


def run_sharded(batch: torch.Tensor, fn: Callable[[torch.Tensor], torch.Tensor], gather: bool = True,
                group=None) -> torch.Tensor:
    """Apply ``fn`` to this rank's contiguous shard of ``batch`` (dim 0) and all-gather the results.

    ``batch`` is the full batch (same on every rank); with ``gather=False`` only the
    local result is returned.  Without an initialised process group this is ``fn(batch)``.
    """
    if not (dist.is_available() and dist.is_initialized()):
        return fn(batch)
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    start, count = shard_range(batch.shape[0], world, rank)
    local = fn(batch[start:start + count])
    if not gather:
        return local
    return all_gather_batch(local, shard_counts(batch.shape[0], world), group)
