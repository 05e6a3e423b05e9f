"""Station sharding across ranks: stations are independent, so the only multi-GPU logic is who owns which
station and how per-rank results are put together.  No collective touches the data path; torch.distributed is used
for the barrier around a timed region, the max over ranks of its duration and the sum of the units processed.

Used by bench.py (NCCL, one rank per GPU) and covered on CPU by tests/test_multirank_gloo.py (gloo, world_size 2)."""
from __future__ import annotations


def shard_range(n_total: int, world: int, rank: int) -> range:
    """Contiguous block of stations owned by `rank` when `n_total` stations are dealt over `world` ranks: sizes differ
    by at most one, lower ranks take the larger shards, every station has exactly one owner."""
    if world <= 0 or not 0 <= rank < world or n_total < 0:
        raise ValueError(f"bad shard request: n_total={n_total} world={world} rank={rank}")
    base, extra = divmod(n_total, world)
    lo = rank * base + min(rank, extra)
    return range(lo, lo + base + (1 if rank < extra else 0))


def weak_range(per_rank: int, rank: int) -> range:
    """Weak scaling (what bench.py runs): every rank owns `per_rank` stations, numbered globally."""
    if per_rank < 0 or rank < 0:
        raise ValueError("bad weak shard request")
    return range(rank * per_rank, (rank + 1) * per_rank)


def owner_of(station: int, n_total: int, world: int) -> int:
    """Inverse of shard_range."""
    if world <= 0 or n_total < 0 or not 0 <= station < n_total:
        raise ValueError(f"station {station} is not one of the {n_total} stations dealt over {world} ranks")
    base, extra = divmod(n_total, world)
    split = extra * (base + 1)
    if station < split:
        return station // (base + 1)
    return extra + (station - split) // base if base else world - 1


def job_throughput(units_local: float, seconds_local: float, dist=None, device=None) -> tuple[float, float, float]:
    """Whole-job throughput = (sum over ranks of the units processed) / (max over ranks of the duration).
    Returns (units_total, seconds_max, units_per_second).  `dist` is torch.distributed (initialised) or None."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return units_local, seconds_local, units_local / seconds_local
    import torch

    t = torch.tensor([seconds_local], dtype=torch.float64, device=device)
    u = torch.tensor([units_local], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.all_reduce(u, op=dist.ReduceOp.SUM)
    return float(u.item()), float(t.item()), float(u.item()) / float(t.item())
