"""Batch sharding across ranks (one process per GPU). States are independent, so the global batch
is range-partitioned and there is no collective on the data path; the only
communication is the barrier / max-reduction of the timing in bench.py."""


def shard_range(total, rank, world):
    """Contiguous [first, first + count) of `total` states owned by `rank` (sizes differ by <= 1)."""
    base, rem = divmod(int(total), int(world))
    first = rank * base + min(rank, rem)
    return first, base + (1 if rank < rem else 0)


def weak_shard(per_rank, rank):
    """Weak scaling: every rank owns `per_rank` states of the global stream."""
    return rank * int(per_rank), int(per_rank)
