"""CPU: crop sharding (LPT partition) and the ragged embedding gather over gloo, world_size 2."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from zoomearth_b200 import sharding


def test_partition_is_deterministic_and_balanced():
    rng = np.random.default_rng(4)
    grids = np.stack([np.ones(200, int), rng.integers(9, 74, 200) * 2, rng.integers(9, 74, 200) * 2], 1)
    cost = sharding.crop_cost(grids)
    for world in (1, 2, 4, 8):
        parts = sharding.partition(cost, world)
        assert parts == sharding.partition(cost, world)
        assert sorted(i for p in parts for i in p) == list(range(200))
        load = np.array([cost[p].sum() for p in parts])
        assert load.max() / load.mean() < 1.05


def _worker(rank, world, port, counts, D):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        cost = np.asarray(counts, dtype=np.float64)
        parts = sharding.partition(cost, world)
        mine = parts[rank]
        rows = [torch.full((counts[i], D), float(i)) + torch.arange(counts[i])[:, None] / 1000 for i in mine]
        local = torch.cat(rows) if mine else torch.zeros((0, D))
        out, got_counts = sharding.gather_embeddings(local, [counts[i] for i in mine], parts)
        expect = torch.cat([torch.full((counts[i], D), float(i)) + torch.arange(counts[i])[:, None] / 1000
                            for i in range(len(counts))])
        assert got_counts.tolist() == list(counts)
        assert torch.equal(out, expect)
    finally:
        dist.destroy_process_group()


def test_gather_embeddings_gloo_world2():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    counts = [5, 81, 324, 7, 16, 40, 3]
    mp.spawn(_worker, args=(2, port, counts, 8), nprocs=2, join=True)


def _rows_worker(rank, world, port, counts, D):
    """The ragged fused gather's row bookkeeping, emulated over gloo: every rank scatters its local embedding rows to
    sharded_rows(...) of a zero buffer; the SUM over ranks must be the global-order embeddings (each row written once)."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        parts = sharding.partition(np.asarray(counts, dtype=np.float64), world)
        starts, rows = sharding.sharded_rows(parts, counts)
        local = torch.cat([torch.full((counts[i], D), float(i)) + torch.arange(counts[i])[:, None] / 1000 for i in parts[rank]])
        buf = torch.zeros((int(starts[-1]), D))
        hits = torch.zeros(int(starts[-1]))
        buf[torch.from_numpy(rows[rank])] = local
        hits[torch.from_numpy(rows[rank])] += 1
        dist.all_reduce(buf)
        dist.all_reduce(hits)
        expect = torch.cat([torch.full((counts[i], D), float(i)) + torch.arange(counts[i])[:, None] / 1000
                            for i in range(len(counts))])
        assert torch.equal(hits, torch.ones_like(hits)) and torch.equal(buf, expect)
    finally:
        dist.destroy_process_group()


def test_sharded_rows_cover_the_gather_buffer_once_gloo_world2():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_rows_worker, args=(2, port, [5, 81, 324, 7, 16, 40, 3, 324, 99], 4), nprocs=2, join=True)


def test_sharded_rows_single_rank_is_identity():
    starts, rows = sharding.sharded_rows([[0, 1, 2]], [4, 2, 3])
    assert starts.tolist() == [0, 4, 6, 9] and rows[0].tolist() == list(range(9))
    starts, rows = sharding.sharded_rows([[2], [], [0, 1]], [4, 2, 3])
    assert rows[0].tolist() == [6, 7, 8] and rows[1].tolist() == [] and rows[2].tolist() == [0, 1, 2, 3, 4, 5]
