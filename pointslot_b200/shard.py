"""Partitioning of a batch of independent stereo pairs over the GPUs of one box.

Pairs never interact, so the path shards with NO data-path collective: rank r of `world` owns the
contiguous block shard_range(total, r, world) and writes results to the same slice of the output
arrays.  torch.distributed is used by the harness only for the start barrier and the max-over-ranks
of the measured time.
"""
from __future__ import annotations


def shard_range(total: int, rank: int, world: int) -> range:
    """Contiguous block partition: GPU g gets pairs [g*B/G, (g+1)*B/G)."""
    if not (0 <= rank < world) or total < 0:
        raise ValueError("bad shard arguments")
    return range(rank * total // world, (rank + 1) * total // world)


def sub_batches(n: int, batch: int):
    """(start, count) of the sub-batches one rank submits for its n pairs."""
    return [(lo, min(batch, n - lo)) for lo in range(0, n, batch)]


def max_over_ranks(ms: float, dist=None, device=None) -> float:
    """Device-measured elapsed time of the slowest rank (the number a multi-GPU result is quoted on)."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return ms
    import torch

    t = torch.tensor([ms], dtype=torch.float64, device=device or "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
