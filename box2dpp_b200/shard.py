"""World sharding across the GPUs of one box.

Worlds never exchange data (a b2World owns all of its state, Box2D/Dynamics/b2World.h:254-287), so the batch
axis is split into contiguous world-index blocks, one per rank / GPU, and the step path has NO collective.
The only cross-rank traffic is host-side bookkeeping around a run: a barrier before the timed region and a
max-reduction of the per-rank device times afterwards (bench.py).
"""
from __future__ import annotations


def shard_range(total_worlds: int, world_size: int, rank: int) -> tuple[int, int]:
    """(first world, world count) of `rank`: contiguous blocks, sizes differ by at most one, union = all worlds."""
    if world_size <= 0 or not 0 <= rank < world_size or total_worlds < 0:
        raise ValueError("bad shard arguments")
    base, extra = divmod(total_worlds, world_size)
    first = rank * base + min(rank, extra)
    return first, base + (1 if rank < extra else 0)


def owner_of(world: int, total_worlds: int, world_size: int) -> int:
    """Rank that owns `world` under shard_range."""
    if world_size <= 0 or not 0 <= world < total_worlds:
        raise ValueError("bad shard arguments")
    base, extra = divmod(total_worlds, world_size)
    split = extra * (base + 1)
    if world < split:
        return world // (base + 1)
    return extra + (world - split) // base


def reduce_max(value: float, dist=None, device=None) -> float:
    """max over ranks of a host float (device times are reduced as the max: the slowest GPU defines the step)."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return value
    import torch
    t = torch.tensor([value], dtype=torch.float64, device=device if device is not None else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def reduce_sum(value: float, dist=None, device=None) -> float:
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return value
    import torch
    t = torch.tensor([value], dtype=torch.float64, device=device if device is not None else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())
