"""CPU coverage of the N>1 path (host logic only, no kernels): contiguous shard ranges and the one collective of the
design -- an all_gather of four int64 counters per rank -- on the gloo backend with world_size 2 and 3."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from ppnet_b200 import sharding


def test_shard_range_partitions_exactly():
    for total in (0, 1, 7, 10000, 1000003):
        for world in (1, 2, 3, 4, 8):
            spans = [sharding.shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0
            for (f0, c0), (f1, _) in zip(spans, spans[1:]):
                assert f0 + c0 == f1                                   # contiguous, no overlap, no gap
            assert spans[-1][0] + spans[-1][1] == total
            assert max(c for _, c in spans) - min(c for _, c in spans) <= 1
    with pytest.raises(ValueError):
        sharding.shard_range(10, 2, 2)
    # bench schedule: ranks of one step tile a contiguous block, steps follow each other
    assert [sharding.step_range(3, r, 4, 100)[0] for r in range(4)] == [1200, 1300, 1400, 1500]


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, total, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        first, count = sharding.shard_range(total, rank, world)
        # stand-in for the kernels' counters: maps = count, valid = maps - rank, obstacles = 45/map, tries = 2/map + rank
        mine = torch.tensor([count, count - rank, 45 * count, 2 * count + rank], dtype=torch.int64)
        per_rank, totals, offs = sharding.gather_counts(mine)
        q.put((rank, first, count, per_rank.tolist(), totals, offs))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_counts_all_gather_gloo(world):
    total = 1001
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, total, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sum(c for _, _, c, _, _, _ in got) == total
    want_rows = [[c, c - r, 45 * c, 2 * c + r] for r, _, c, _, _, _ in got]
    for rank, first, count, per_rank, totals, offs in got:
        assert per_rank == want_rows                                   # every rank sees every rank's counters
        assert totals["maps"] == total and totals["valid_paths"] == total - sum(range(world))
        assert totals["accepted_obstacles"] == 45 * total
        assert offs["maps"] == first                                   # exclusive scan == start of this rank's shard
        assert offs["valid_paths"] == sum(w[1] for w in want_rows[:rank])


def test_gather_counts_single_process():
    per_rank, totals, offs = sharding.gather_counts(torch.tensor([5, 4, 200, 11], dtype=torch.int64))
    assert per_rank.tolist() == [[5, 4, 200, 11]] and totals["valid_paths"] == 4 and offs["maps"] == 0
    with pytest.raises(ValueError):
        sharding.gather_counts(torch.zeros(3, dtype=torch.int64))
