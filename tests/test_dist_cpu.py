"""Batch sharding + output all-gather on CPU (gloo, world_size 2): the N>1 host logic of bench.py."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from waternet_b200.dist import PassGather, all_gather_batch, run_sharded, shard_counts, shard_range


def test_shard_range_covers_the_batch_exactly():
    for n in (0, 1, 2, 7, 16, 128, 129):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and sum(c for _, c in spans) == n
            for (s0, c0), (s1, _) in zip(spans, spans[1:]):
                assert s0 + c0 == s1
            assert max(c for _, c in spans) - min(c for _, c in spans) <= 1
    assert shard_counts(128, 8) == [16] * 8
    with pytest.raises(ValueError):
        shard_range(4, 2, 2)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import preprocess as opre  # stand-in for the CUDA engine on a CPU box (tests only)
        rng = np.random.default_rng(0)
        batch = torch.from_numpy(rng.integers(0, 256, (n, 24, 40, 3), dtype=np.uint8))

        def fn(local):  # independent per-image work, like Engine.enhance
            return torch.from_numpy(np.stack([opre.gamma_correction(a.numpy()) for a in local])
                                    if len(local) else np.zeros((0, 24, 40, 3), np.uint8))

        full = run_sharded(batch, fn, gather=True)
        local = run_sharded(batch, fn, gather=False)
        start, count = shard_range(n, world, rank)
        ok = torch.equal(full, fn(batch)) and torch.equal(local, full[start:start + count])
        counts = shard_counts(n, world)
        again = all_gather_batch(local, counts)
        ok = ok and torch.equal(again, full)
        # the per-pass gather bench.py and Enhancer.submit(on_pass=...) use: equal batches, passes of 2 images
        per = 3
        mine = fn(batch[:per]) + rank  # rank-dependent content
        pg = PassGather(tuple(mine.shape), mine.dtype, mine.device)
        for a in range(0, per, 2):
            pg.on_pass(mine[a:a + 2], a, min(per, a + 2))
        want = torch.cat([fn(batch[:per]) + r for r in range(world)])
        ok = ok and pg.calls == 2 and torch.equal(pg.result(), want)
        # no CUDA, no peer memory: every rank must agree on the NCCL / gloo all-gather form instead of hanging
        from waternet_b200.dist import PeerGather
        fb = PeerGather.create(tuple(mine.shape), mine.dtype, mine.device)
        ok = ok and isinstance(fb, PassGather)
        for a in range(0, per, 2):
            fb.on_pass(mine[a:a + 2], a, min(per, a + 2))
        fb.finish()
        ok = ok and torch.equal(fb.result(), want)
        ret[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n", [4, 5, 1])
def test_sharded_result_equals_single_process_result(n):
    world = 2
    port = _free_port()
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_worker, args=(world, port, n, ret), nprocs=world, join=True)
        assert dict(ret) == {0: True, 1: True}


def test_run_sharded_without_process_group_is_identity():
    t = torch.arange(12).reshape(4, 3)
    assert torch.equal(run_sharded(t, lambda x: x * 2), t * 2)
