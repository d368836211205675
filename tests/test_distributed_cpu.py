"""CPU, world_size 2, gloo: the host side of the N>1 paths (hpe_b200/sharding.py).  The GPU kernels
are not involved: a rank's "step" is a deterministic numpy function keyed by the global particle
index, standing in for hpe_swarm_step, so that sharded == unsharded can be checked here."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from hpe_b200 import sharding

ROW = 27


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _mock_rows(idx, gen, table):
    """stand-in for one generation of the owned particles: depends on the global index, the
    generation and (like the informant lookup) on rows owned by other ranks"""
    n = table.shape[0]
    inf = (idx * 7 + gen * 13 + 1) % n
    rows = np.empty((len(idx), ROW))
    rows[:, 0] = 0.5 * table[idx, 0] + 0.25 * table[inf, 0] + np.cos(idx + gen)
    rows[:, 1:] = 0.9 * table[idx, 1:] + 0.1 * table[inf, 1:] + 0.01 * gen
    return rows


def _run_unsharded(n, gens):
    table = np.zeros((n, ROW))
    idx = np.arange(n)
    table[:] = np.cos(np.outer(idx + 1, np.arange(ROW) + 1))
    for g in range(1, gens):
        table = _mock_rows(idx, g, table.copy())
    return table


def _worker(rank, world, port, n, gens, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = sharding.shard_range(n, rank, world)
    idx = np.arange(lo, hi)
    table = torch.zeros((n, ROW), dtype=torch.float64)
    table[lo:hi] = torch.from_numpy(np.cos(np.outer(idx + 1, np.arange(ROW) + 1)))
    sharding.allgather_rows(table, n)                       # after "init"
    for g in range(1, gens):
        rows = _mock_rows(idx, g, table.numpy().copy())     # "step" on the owned rows
        table[lo:hi] = torch.from_numpy(rows)
        table[:lo] = float("nan"); table[hi:] = float("nan")   # foreign rows are stale until exchanged
        sharding.allgather_rows(table, n)
    t = sharding.max_over_ranks(float(rank + 1))
    s = sharding.sum_over_ranks(float(rank + 1))
    if rank == 0:
        np.save(out, np.concatenate([table.numpy().ravel(), [t, s]]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n", [64, 101])        # divisible and ragged shards
def test_sharded_table_exchange_equals_unsharded(tmp_path, n):
    world, gens = 2, 5
    out = str(tmp_path / "res.npy")
    mp.spawn(_worker, args=(world, _free_port(), n, gens, out), nprocs=world, join=True)
    res = np.load(out)
    table, t, s = res[:-2].reshape(n, ROW), res[-2], res[-1]
    assert np.array_equal(table, _run_unsharded(n, gens))
    assert t == 2.0 and s == 3.0


def test_shard_range_properties():
    for n in (0, 1, 7, 64, 65536, 262144, 16384):
        for world in (1, 2, 3, 4, 8):
            edges = [sharding.shard_range(n, r, world) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == n
            for (a, b), (c, d) in zip(edges[:-1], edges[1:]):
                assert b == c
            sizes = [b - a for a, b in edges]
            assert max(sizes) - min(sizes) <= 1


def test_single_process_helpers_are_no_ops():
    t = torch.arange(10 * ROW, dtype=torch.float64).view(10, ROW)
    assert sharding.allgather_rows(t.clone(), 10).equal(t)
    assert sharding.max_over_ranks(3.5) == 3.5 and sharding.world() == 1 and sharding.rank() == 0
