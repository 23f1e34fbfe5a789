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


def _digest_worker(rank, world, port, total, q):
    import numpy as np
    from oracle import ppnet_oracle as orc
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        first, count = sharding.shard_range(total, rank, world)
        # stand-in for a rank's generated maps: unit g is a pure function of g (as the Philox-keyed generator's outputs are)
        g = np.arange(first, first + count, dtype=np.int64)
        data = np.stack([g * 3 + 1, g * g % 1000003, g ^ 0x5555], axis=1).astype(np.int32)
        rows = (g % 4).astype(np.int32)
        d = (orc.digest_u32(data, first, salt=4) + orc.digest_u32(data, first, rows=rows, row_words=1, salt=5)) & 0xFFFFFFFFFFFFFFFF
        signed = d - (1 << 64) if d >= (1 << 63) else d
        combined, per_rank = sharding.combine_digests(torch.tensor([signed], dtype=torch.int64))
        q.put((rank, combined, per_rank))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_digest_combination_is_world_size_invariant_gloo(world):
    """Config 4's identity proof on the CPU: per-shard digests, all-gathered and summed mod 2^64, equal the digest of the whole
    range computed by one process (the kernels' digest is checked against the same restatement in the GPU tests)."""
    import numpy as np
    from oracle import ppnet_oracle as orc
    total = 1001
    g = np.arange(total, dtype=np.int64)
    data = np.stack([g * 3 + 1, g * g % 1000003, g ^ 0x5555], axis=1).astype(np.int32)
    want = (orc.digest_u32(data, 0, salt=4) + orc.digest_u32(data, 0, rows=(g % 4).astype(np.int32), row_words=1, salt=5)) & 0xFFFFFFFFFFFFFFFF
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_digest_worker, args=(r, world, port, total, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, combined, per_rank in got:
        assert combined == want and len(per_rank) == world
    assert sharding.combine_digests(torch.tensor([5], dtype=torch.int64)) == (5, [5])
