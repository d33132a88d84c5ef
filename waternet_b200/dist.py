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

    def finish(self) -> None:
        """Nothing to do: every all-gather is complete on the stream it was issued on."""

    def result(self) -> torch.Tensor:
        return self.gathered.view((-1,) + tuple(self.gathered.shape[2:]))


class PeerGather:
    """All-gather of every rank's output batch by **copy-engine pushes into peer memory** (NVLink / NVSwitch), pass by
    pass, under the next pass's kernels -- no kernel, no SM.

    Why not NCCL here: the convolution kernels are persistent and fill every SM's shared memory, so an NCCL kernel
    launched next to them cannot co-reside; it takes SMs at a kernel boundary and delays the next convolution launch
    by about its own duration (measured at N=2: 5.3 ms per step for four 25 MB all-gathers, `PassGather`).  Here every
    rank owns a `(world, B, ...)` buffer, exports it once over CUDA IPC (``torch.multiprocessing.reductions``, handles
    exchanged through the process group), and ``on_pass`` copies the finished pass ``local[a:b]`` into slot
    ``[rank, a:b]`` of every rank's buffer with ``cudaMemcpyPeerAsync`` on the caller's side stream.  ``finish()`` --
    once per step, after the last pass -- is a one-element NCCL all-reduce on that stream: when it completes, every
    rank's pushes (stream-ordered before it) have landed everywhere.  A consumer that reads ``result()`` must be done
    before the next step's first ``on_pass`` (the bench does not read it; double-buffer otherwise).

    Falls back to :class:`PassGather` when the ranks are not all on one node or IPC is unavailable
    (``PeerGather.create``).
    """

    def __init__(self, local_shape, dtype, device, group=None):
        from torch.multiprocessing.reductions import reduce_tensor
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.device = torch.device(device)
        self.gathered = torch.empty((self.world,) + tuple(local_shape), dtype=dtype, device=self.device)
        rebuild, args = reduce_tensor(self.gathered)
        handles = [None] * self.world
        dist.all_gather_object(handles, (rebuild, args), group=group)
        self.peers = []
        for r, (fn, a) in enumerate(handles):
            self.peers.append(self.gathered if r == self.rank else fn(*a))  # a tensor aliasing rank r's buffer
        self._flag = torch.zeros(1, device=self.device)
        self.calls = 0

    @classmethod
    def create(cls, local_shape, dtype, device, group=None):
        """PeerGather when every rank can open every other rank's buffer, else the NCCL PassGather."""
        ok = 1
        obj = None
        try:
            obj = cls(local_shape, dtype, device, group)
        except Exception:  # IPC refused (different nodes, container limits): every rank must agree on the fallback
            ok = 0
        flag = torch.tensor([ok], device=device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
        if int(flag.item()) == 1:
            return obj
        return PassGather(local_shape, dtype, device, group)

    def on_pass(self, local_chunk: torch.Tensor, a: int, b: int) -> None:
        for r in range(self.world):  # own slot first (local copy), then the peers round-robin from rank + 1
            p = (self.rank + r) % self.world
            self.peers[p][self.rank, a:b].copy_(local_chunk, non_blocking=True)
        self.calls += 1

    def finish(self) -> None:
        dist.all_reduce(self._flag, group=self.group)

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
