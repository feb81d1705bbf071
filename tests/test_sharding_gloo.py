"""CPU tests of the multi-GPU host logic (seal-embedded_b200/shard.py): the contiguous batch
partition and the optional collation all-gather, run with world_size 2 over gloo."""
from __future__ import annotations

import importlib
import os
import socket

import numpy as np
import pytest


def _shard():
    return importlib.import_module("seal-embedded_b200.shard")


def test_shard_range_partitions_exactly():
    sh = _shard()
    for batch in (0, 1, 2, 7, 8, 9, 65536, 262144, 131072, 1000003):
        for world in (1, 2, 3, 4, 8):
            parts = [sh.shard_range(batch, r, world) for r in range(world)]
            assert parts[0].first == 0 and parts[-1].stop == batch
            for a, b in zip(parts, parts[1:]):
                assert a.stop == b.first
            counts = [p.count for p in parts]
            assert max(counts) - min(counts) <= 1 and sum(counts) == batch
    for item in (0, 1, 4, 8):
        assert sh.shard_range(9, sh.owner_of(item, 9, 4), 4).first <= item < sh.shard_range(
            9, sh.owner_of(item, 9, 4), 4).stop
    with pytest.raises(ValueError):
        sh.shard_range(8, 2, 2)
    with pytest.raises(ValueError):
        sh.owner_of(9, 9, 4)


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank: int, world: int, port: int, batch: int, q):
    import torch
    import torch.distributed as dist

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        sh = importlib.import_module("seal-embedded_b200.shard")
        mine = sh.shard_range(batch, rank, world)
        np_, n = 2, 16
        # stand-in "ciphertexts": item b is filled with a function of b only, as independent items are
        items = torch.arange(mine.first, mine.stop, dtype=torch.int32).view(-1, 1, 1, 1)
        local = (items * 1000 + torch.arange(np_ * 2 * n, dtype=torch.int32).view(1, np_, 2, n)).contiguous()
        full = sh.all_gather_ciphertexts(local, batch)
        exp = (torch.arange(batch, dtype=torch.int32).view(-1, 1, 1, 1) * 1000 +
               torch.arange(np_ * 2 * n, dtype=torch.int32).view(1, np_, 2, n))
        ok = tuple(full.shape) == (batch, np_, 2, n) and bool(torch.equal(full, exp))
        # max-over-ranks timing reduction used by bench.py
        t = torch.tensor([float(rank + 1)], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ok = ok and float(t.item()) == float(world)
        q.put((rank, ok, mine.count))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("batch", [9, 2, 1, 64])
def test_all_gather_ciphertexts_gloo_world2(batch):
    import torch.multiprocessing as mp

    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, batch, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok, _ in res)
    assert sum(c for _, _, c in res) == batch
