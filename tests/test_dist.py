"""N > 1 host logic on CPU: frame-range sharding and the final feature gather over gloo, world size 2 (and 3)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from phdfx.dist import all_counts, gather_rows, shard_range


def test_shard_range_partitions_everything():
    for n in (0, 1, 7, 256, 2000, 200_000):
        for world in (1, 2, 3, 4, 8):
            ranges = [shard_range(n, r, world) for r in range(world)]
            assert ranges[0][0] == 0 and ranges[-1][1] == n
            for (a0, a1), (b0, b1) in zip(ranges, ranges[1:]):
                assert a1 == b0 and a0 <= a1
            assert sum(all_counts(n, world)) == n
            assert max(all_counts(n, world)) - min(all_counts(n, world)) <= max(1, (n + world - 1) // world)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_total, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    full = torch.arange(n_total * 5, dtype=torch.float32).view(n_total, 5)
    lo, hi = shard_range(n_total, rank, world)
    out = gather_rows(full[lo:hi].clone(), n_total, dst=0)
    if rank == 0:
        q.put(bool(torch.equal(out, full)))
    else:
        assert out is None
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,n_total", [(2, 10), (2, 7), (3, 4), (2, 1)])
def test_gather_rows_gloo(world, n_total):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_total, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert q.get(timeout=5) is True


def test_gather_rows_single_process_is_identity():
    x = torch.randn(4, 3)
    assert gather_rows(x, 4) is x
