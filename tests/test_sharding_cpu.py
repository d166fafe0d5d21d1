"""World-size-2 gloo test of the multi-GPU host logic (sharding + result-table gather): no GPU
compute -- each rank fabricates the table of its shard as a known function of the variant id."""
import os
import socket

import numpy as np
import pytest

from pyseer_b200 import sharding as sh


def test_shard_ranges_cover_everything():
    for n in (0, 1, 7, 100, 6250001):
        for world in (1, 2, 3, 8):
            r = [sh.shard_range(n, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[k][1] == r[k + 1][0] for k in range(world - 1))
            sizes = [b - a for a, b in r]
            assert max(sizes) - min(sizes) <= 1


def test_pack_unpack_roundtrip():
    n = 13
    buf = np.zeros(n * sh.ROW_BYTES, dtype=np.uint8)
    cols = sh.unpack_table(buf, n)
    cols['pvalue'][:] = np.arange(n) * 0.5
    cols['flags'][:] = np.arange(n) + 7
    again = sh.unpack_table(buf.copy(), n)
    assert np.array_equal(again['pvalue'], np.arange(n) * 0.5)
    assert np.array_equal(again['flags'], np.arange(n) + 7)
    ptrs = sh.table_pointers(1000, n)
    assert ptrs['carriers'] == 1000 and ptrs['missing'] == 1000 + 4 * n


def _fake_table(first, last):
    n = last - first
    buf = np.zeros(n * sh.ROW_BYTES, dtype=np.uint8)
    cols = sh.unpack_table(buf, n)
    ids = np.arange(first, last)
    cols['carriers'][:] = ids % 97
    cols['pvalue'][:] = 1.0 / (1.0 + ids)
    cols['beta'][:] = np.sin(ids)
    cols['flags'][:] = ids % 5
    return buf


def _worker(rank, world, port, n_total, q):
    import torch
    import torch.distributed as dist
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        ranges = [sh.shard_range(n_total, k, world) for k in range(world)]
        first, last = ranges[rank]
        table = torch.from_numpy(_fake_table(first, last))
        out = sh.gather_tables(table, [b - a for a, b in ranges], dst=0)
        if rank == 0:
            merged = sh.merge_tables([t.numpy() for t in out], [b - a for a, b in ranges])
            ids = np.arange(n_total)
            ok = (np.array_equal(merged['carriers'], ids % 97) and
                  np.array_equal(merged['pvalue'], 1.0 / (1.0 + ids)) and
                  np.array_equal(merged['beta'], np.sin(ids)) and
                  np.array_equal(merged['flags'], ids % 5))
            q.put(bool(ok))
        else:
            assert out is None
        # max-over-ranks timing reduction as bench.py does it
        t = torch.tensor([float(rank + 1)], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        assert float(t[0]) == world
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('n_total', [1001, 64])
def test_gather_world2_gloo(n_total):
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_total, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert q.get(timeout=10) is True
