"""World-size-2 `gloo` tests (CPU) of the data-parallel host logic: bucketed all-reduce of
the flat gradient buffer launched from inside backward, row sharding, loss scaling."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from segmminterest_b200.dp import GradBuckets, shard_rows


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, n, bucket_bytes, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(100 + rank)
        flat = torch.randn(n, generator=g)
        mine = flat.clone()
        buckets = GradBuckets(flat, None, bucket_bytes)
        assert buckets.world == world
        buckets.begin()
        # backward finishes the flat buffer from the top (head, last layers) down to offset 0
        cuts = [int(n * f) for f in (0.9, 0.7, 0.65, 0.3, 0.05, 0.0)]
        for lo in cuts:
            buckets.ready(lo)
        buckets.finish()
        others = [torch.randn(n, generator=torch.Generator().manual_seed(100 + r)) for r in range(world)]
        expect = torch.stack(others).sum(0)
        ok = torch.allclose(flat, expect, rtol=0, atol=1e-6) and not torch.equal(flat, mine)
        out.put((rank, bool(ok), buckets.n_collectives))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("bucket_bytes,min_coll", [(1 << 30, 1), (4096, 3)])
def test_bucketed_allreduce_world2_gloo(bucket_bytes, min_coll):
    world, n = 2, 10000
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, bucket_bytes, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ok, ncoll in res:
        assert ok, f"rank {rank}: reduced gradient buffer differs from the sum over ranks"
        assert ncoll >= min_coll
    assert len({n for _, _, n in res}) == 1, "ranks must issue the same sequence of collectives"


def test_single_process_is_a_no_op():
    flat = torch.arange(10.0)
    b = GradBuckets(flat, None, 16)
    b.begin(); b.ready(5); b.ready(0); b.finish()
    assert b.n_collectives == 0 and torch.equal(flat, torch.arange(10.0))


def test_shard_rows_partitions_the_global_batch():
    for n, world in ((4096, 8), (1000, 3), (7, 2)):
        spans = [shard_rows(n, r, world) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        for (a0, a1), (b0, b1) in zip(spans, spans[1:]):
            assert a1 == b0 and a1 > a0


def _worker_skip(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from segmminterest_b200.dp import exchange_rows
        n = 1000
        flat = torch.full((n,), float(rank + 1))
        buckets = GradBuckets(flat, None, 256, skip=[(100, 50), (600, 100)])      # two "embedding tables" inside the flat buffer
        buckets.begin()
        for lo in (900, 640, 620, 300, 120, 0):
            buckets.ready(lo)
        buckets.finish()
        total = float(sum(range(1, world + 1)))
        keep = torch.zeros(n, dtype=torch.bool)
        keep[100:150] = True
        keep[600:700] = True
        ok = bool(torch.all(flat[keep] == rank + 1)) and bool(torch.all(flat[~keep] == total))
        # row-sparse exchange: every rank ends up with the concatenation of all ranks' (ids, rows)
        ids = torch.tensor([rank, 7, 7], dtype=torch.int64)
        rows = torch.full((3, 4), float(rank + 1))
        all_ids, all_rows = exchange_rows(ids, rows)
        ok = ok and all_ids.tolist() == [0, 7, 7, 1, 7, 7][: 3 * world] and all_rows[:, 0].tolist() == [1.0] * 3 + [2.0] * 3
        out.put((rank, ok))
    finally:
        dist.destroy_process_group()


def test_skipped_table_ranges_stay_local_and_rows_are_exchanged_world2_gloo():
    """the dense all-reduce leaves the embedding-table ranges alone (their gradients travel as (ids, rows))."""
    world = 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker_skip, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok in res), res
