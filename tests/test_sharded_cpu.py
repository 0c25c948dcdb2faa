"""Multi-rank host logic on CPU: world_size-2 (and 3) gloo groups, no GPU.  The kernels are not involved:
the tables are deterministic functions of the unit id, so the gathered result is known exactly."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _table_for(first, count, rows=5, cols=3):
    ids = torch.arange(first, first + count, dtype=torch.float64).view(-1, 1, 1)
    return ids * 100 + torch.arange(rows, dtype=torch.float64).view(1, -1, 1) * 10 + torch.arange(cols, dtype=torch.float64).view(1, 1, -1)


def _worker(rank, world, port, n_units, q):
    from pymhealth_b200 import sharded
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        a, b = sharded.shard_range(n_units, rank, world)
        local = _table_for(a, b - a)
        full = sharded.gather_tables(local, n_units)
        ok_all = torch.equal(full, _table_for(0, n_units))
        to0 = sharded.gather_tables(local, n_units, dst=0)
        ok_dst = (to0 is None) if rank != 0 else torch.equal(to0, _table_for(0, n_units))
        # max-over-ranks timing reduction used by bench.py
        t = torch.tensor([float(rank + 1)])
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        q.put((rank, bool(ok_all), bool(ok_dst), float(t)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n_units", [(2, 7), (2, 8), (3, 10), (2, 1)])
def test_gather_tables_gloo(world, n_units):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_units, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ok_all, ok_dst, tmax in res:
        assert ok_all and ok_dst and tmax == world


def test_shard_ranges_partition_everything():
    from pymhealth_b200 import sharded
    for n in (0, 1, 7, 125, 1000, 10000):
        for w in (1, 2, 4, 8):
            rs = [sharded.shard_range(n, r, w) for r in range(w)]
            assert rs[0][0] == 0 and rs[-1][1] == n
            assert all(rs[i][1] == rs[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in rs]
            assert max(sizes) - min(sizes) <= 1 and sizes == sharded.shard_sizes(n, w)
