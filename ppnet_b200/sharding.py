"""Map-parallel sharding of the generator over ranks.

Every map is independent given its global index g (Philox counters are keyed by g), so the path shards with NO
data-path collective: rank r takes a contiguous range of global map indices.  The only exchange is one
all_gather of four int64 counters per rank {maps, valid_paths, accepted_obstacles, placement_tries}, from which every
rank derives the global totals and its own output offset (exclusive scan) -- e.g. where its records start in a
concatenated dataset.  Works on any torch.distributed backend (nccl on the GPUs; gloo in the CPU tests)."""
import torch
import torch.distributed as dist

COUNTER_NAMES = ("maps", "valid_paths", "accepted_obstacles", "placement_tries")


def shard_range(total, rank, world):
    """Contiguous split of [0, total): the first (total % world) ranks get one extra unit.  -> (first, count)."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank/world: %r/%r" % (rank, world))
    base, extra = divmod(int(total), int(world))
    first = rank * base + min(rank, extra)
    return first, base + (1 if rank < extra else 0)


def step_range(step, rank, world, maps_per_rank):
    """bench.py's weak-scaling schedule: step `it` covers [it*world*M, (it+1)*world*M), rank r the r-th slice."""
    return (step * world + rank) * maps_per_rank, maps_per_rank


def gather_counts(counters):
    """counters: int64[4] tensor on this rank (device for nccl, CPU for gloo).  -> (per_rank int64[world,4] on CPU,
    totals dict, this rank's exclusive-scan offsets dict)."""
    if counters.dtype != torch.int64 or counters.numel() != len(COUNTER_NAMES):
        raise ValueError("counters must be int64[%d]" % len(COUNTER_NAMES))
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        world, rank = dist.get_world_size(), dist.get_rank()
        parts = [torch.zeros_like(counters) for _ in range(world)]
        dist.all_gather(parts, counters.contiguous())
        per_rank = torch.stack(parts).cpu()
    else:
        rank, per_rank = 0, counters.reshape(1, -1).cpu()
    totals = {n: int(v) for n, v in zip(COUNTER_NAMES, per_rank.sum(0).tolist())}
    offs = per_rank[:rank].sum(0).tolist() if rank else [0] * len(COUNTER_NAMES)
    return per_rank, totals, {n: int(v) for n, v in zip(COUNTER_NAMES, offs)}


def combine_digests(digest):
    """The identity proof's collective: every rank contributes the 64-bit content digest of its shard (an int64 tensor of one
    element holding the uint64 bits, ppnet_digest_u32 / ops.digest_maps); the digests are all-gathered and summed modulo 2^64.
    Because the digest is additive over any split of the global map range, the result must not depend on the world size.
    -> (combined Python int, per-rank list)."""
    if digest.dtype != torch.int64 or digest.numel() != 1:
        raise ValueError("digest must be an int64[1] tensor")
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        parts = [torch.zeros_like(digest) for _ in range(dist.get_world_size())]
        dist.all_gather(parts, digest.contiguous())
    else:
        parts = [digest]
    per_rank = [int(p.item()) & 0xFFFFFFFFFFFFFFFF for p in parts]
    return sum(per_rank) & 0xFFFFFFFFFFFFFFFF, per_rank
