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


class _DeviceBlock:
    """Raw device memory as a torch tensor (``__cuda_array_interface__``); the owner keeps the memory alive."""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}


class PeerGather:
    """All-gather of every rank's output batch by **copy-engine pushes into peer memory** (NVLink / NVSwitch), pass by
    pass, under the next pass's kernels -- no kernel, no SM, and nothing submitted to the peer GPU.

    Two costs of doing this with a collective library, measured at N=2 (``tools/probe_gather.py``,
    ``profiles/r2_probe_gather.log``): the convolution kernels are persistent and fill every SM's shared memory, so an
    NCCL kernel launched next to them cannot co-reside -- it takes SMs at a kernel boundary and delays the next
    convolution launch by about its own duration (5.3 ms per step for four 25 MB all-gathers, `PassGather`); and a
    ``tensor.copy_`` into a peer's IPC-mapped tensor orders itself against the destination device's streams through
    the context THIS process holds on the peer GPU, which makes that GPU time-slice away from its owner (0.4 ms per
    copy with these kernels resident).  Here every rank owns one block -- ``(world, B, ...)`` result plus a flag word
    per rank -- from ``wn_peer_alloc``, sends its IPC handle through the process group, and maps every other rank's
    block into its OWN device context (``wn_peer_open``).  ``on_pass`` pushes the finished pass ``local[a:b]`` into
    slot ``[rank, a:b]`` of every rank's block with ``wn_memcpy_async`` (``cudaMemcpyAsync``: this GPU's copy engine
    writing through NVLink) on the caller's side stream.

    ``finish()`` -- once per step, after the last pass -- is the completion signal, again without a kernel: the step
    number is written to a local word by the stream's front end (``wn_stream_write_value32``), pushed into slot
    ``[rank]`` of every peer's flag array by the copy engine -- stream-ordered behind this step's pushes, so a peer
    that sees it has the data -- and the stream then holds (``wn_stream_wait_value32``, ``>=``) until every peer's word
    for this step has arrived in the local flag array.  A consumer that reads ``result()`` must be done before the
    next step's first ``on_pass`` (the bench does not read it; double-buffer otherwise).

    Falls back to :class:`PassGather` when the ranks are not all on one node or IPC is unavailable
    (``PeerGather.create``).
    """

    def __init__(self, local_shape, dtype, device, group=None):
        import ctypes
        from . import _lib
        self._check = _lib.check
        self._lib = _lib.load()
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.device = torch.device(device)
        self.local_shape = tuple(local_shape)
        self.dtype = dtype
        item = torch.empty((), dtype=dtype).element_size()
        self.slot_bytes = item * int(torch.Size(self.local_shape).numel())   # one rank's batch
        self.row_bytes = self.slot_bytes // max(1, self.local_shape[0])      # one image
        data_bytes = (self.world * self.slot_bytes + 255) // 256 * 256
        self._flags_off = data_bytes                   # int32[world]: [r] = last step rank r finished
        self._word_off = data_bytes + 256              # int32: this rank's step number (source of the flag pushes)
        self._bytes = data_bytes + 512
        self._base, self._peer_base = None, {}
        base = ctypes.c_void_p()
        handle = (ctypes.c_ubyte * _lib.PEER_HANDLE_BYTES)()
        with torch.cuda.device(self.device):
            try:
                self._check(self._lib.wn_peer_alloc(self._bytes, ctypes.byref(base), handle), "wn_peer_alloc")
                self._base = base.value
                mine = bytes(handle)
            except Exception:
                mine = None  # still take part in the handle exchange: the other ranks are waiting in it
            handles = [None] * self.world
            dist.all_gather_object(handles, mine, group=group)
            try:
                if any(hb is None for hb in handles):
                    raise _lib.WaterNetLibraryError("a rank could not allocate or export its peer buffer")
                for r, hb in enumerate(handles):
                    if r == self.rank:
                        continue
                    ptr = ctypes.c_void_p()
                    buf = (ctypes.c_ubyte * _lib.PEER_HANDLE_BYTES).from_buffer_copy(hb)
                    self._check(self._lib.wn_peer_open(buf, ctypes.byref(ptr)), "wn_peer_open")
                    self._peer_base[r] = ptr.value
            except Exception:
                self.close(collective=False)  # nothing of a half-built exchange stays mapped or allocated
                raise
        block = torch.as_tensor(_DeviceBlock(self._base, self._bytes), device=self.device)
        self._block = block
        self.gathered = block[:self.world * self.slot_bytes].view(dtype).view((self.world,) + self.local_shape)
        self.step = 0
        self.calls = 0

    @classmethod
    def create(cls, local_shape, dtype, device, group=None):
        """PeerGather when every rank can map every other rank's buffer, else the NCCL PassGather."""
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
        if obj is not None:
            obj.close(collective=False)
        return PassGather(local_shape, dtype, device, group)

    def _base_of(self, r: int) -> int:
        return self._base if r == self.rank else self._peer_base[r]

    def on_pass(self, local_chunk: torch.Tensor, a: int, b: int) -> None:
        if not local_chunk.is_contiguous() or local_chunk.dtype != self.dtype:
            raise ValueError("PeerGather.on_pass needs a contiguous chunk of the gathered dtype")
        stream = torch.cuda.current_stream(self.device).cuda_stream
        off = self.rank * self.slot_bytes + a * self.row_bytes
        nbytes = (b - a) * self.row_bytes
        with torch.cuda.device(self.device):
            for r in range(self.world):  # own slot first (local copy), then the peers round-robin from rank + 1
                p = (self.rank + r) % self.world
                self._check(self._lib.wn_memcpy_async(self._base_of(p) + off, local_chunk.data_ptr(), nbytes, stream),
                            "wn_memcpy_async")
        self.calls += 1

    def addresses(self, a: int):
        """Where images [a, ...) of THIS rank's batch belong in every rank's block (own block first): the
        ``peer_out`` of ``Engine.enhance`` -- the kernel that writes a pass's output stores it there as well, so the
        exchange needs no copy at all (``wn_enhance_u8_peers``; then ``signal()`` after the last pass)."""
        off = self.rank * self.slot_bytes + a * self.row_bytes
        return [self._base_of((self.rank + r) % self.world) + off for r in range(self.world)]

    def signal(self) -> None:
        """On the current stream, behind everything that wrote this step's output into the peers' blocks (kernel
        stores or ``on_pass`` pushes): publish this rank's step number to every peer."""
        self.step += 1
        value = self.step & 0xFFFFFFFF
        stream = torch.cuda.current_stream(self.device).cuda_stream
        word = self._base + self._word_off
        with torch.cuda.device(self.device):
            self._check(self._lib.wn_stream_write_value32(stream, word, value), "wn_stream_write_value32")
            for r in range(1, self.world):
                p = (self.rank + r) % self.world
                self._check(self._lib.wn_memcpy_async(self._peer_base[p] + self._flags_off + 4 * self.rank, word, 4,
                                                      stream), "wn_memcpy_async")

    def wait(self) -> None:
        """Hold the current stream until every peer has signalled the step this rank signalled last."""
        value = self.step & 0xFFFFFFFF
        stream = torch.cuda.current_stream(self.device).cuda_stream
        with torch.cuda.device(self.device):
            for r in range(1, self.world):
                p = (self.rank + r) % self.world
                self._check(self._lib.wn_stream_wait_value32(stream, self._base + self._flags_off + 4 * p, value),
                            "wn_stream_wait_value32")

    def finish(self) -> None:
        """Completion of this step's exchange on the current stream: signal every peer, wait for every peer."""
        self.signal()
        self.wait()

    def result(self) -> torch.Tensor:
        return self.gathered.view((-1,) + tuple(self.gathered.shape[2:]))

    def __del__(self):
        try:
            self.close(collective=False)
        except Exception:  # interpreter shutdown: the driver reclaims the memory with the process
            pass

    def close(self, collective: bool = True) -> None:
        """Unmap the peers' blocks and free this rank's (after every rank has stopped pushing, when collective).
        Tensors obtained from ``result()`` / ``gathered`` alias that memory: copy what must outlive the exchange."""
        if getattr(self, "_base", None) is None and not getattr(self, "_peer_base", None):
            return
        with torch.cuda.device(self.device):
            torch.cuda.synchronize(self.device)
            if collective and dist.is_initialized():
                dist.barrier(group=self.group)
            for ptr in self._peer_base.values():
                self._lib.wn_peer_close(ptr)
            self._peer_base = {}
            if collective and dist.is_initialized():
                dist.barrier(group=self.group)  # nobody still maps a block that is about to be freed
            self.gathered = self._block = None
            if self._base is not None:
                self._lib.wn_peer_free(self._base)
            self._base = None


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
