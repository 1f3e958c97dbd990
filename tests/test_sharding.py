"""CPU: multi-GPU plumbing of the bench (contiguous sharding, barrier, max-over-ranks) on world_size-2 gloo."""
import os

import pytest

from pointslot_b200.shard import max_over_ranks, shard_range, sub_batches


def test_shard_range_partitions_exactly():
    for total in (0, 1, 7, 1000, 8000):
        for world in (1, 2, 3, 4, 8):
            seen = []
            for r in range(world):
                seen += list(shard_range(total, r, world))
            assert seen == list(range(total))
    assert [len(shard_range(8000, r, 8)) for r in range(8)] == [1000] * 8
    with pytest.raises(ValueError):
        shard_range(10, 2, 2)


def test_sub_batches():
    assert sub_batches(1000, 125) == [(i * 125, 125) for i in range(8)]
    assert sub_batches(10, 4) == [(0, 4), (4, 4), (8, 2)]
    assert sub_batches(0, 4) == []


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = list(shard_range(10, rank, world))
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)
    dist.barrier()
    slowest = max_over_ranks(10.0 + rank, dist)
    q.put((rank, gathered, slowest))
    dist.destroy_process_group()


def test_gloo_world2_sharding_and_timing():
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs: p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs: p.join(timeout=60)
    for rank, gathered, slowest in res:
        assert sum(gathered, []) == list(range(10))  # disjoint, complete, in order
        assert slowest == 11.0                        # max over ranks, identical on every rank
