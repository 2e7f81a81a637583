"""Multi-GPU partitioning of the path: the dataset is cut into `batch_size`-row blocks and block i
goes to rank i mod N (what Ray's actor pool does for the reference with one actor per GPU,
notebook :908-913). There is no exchange step on the inference path, hence no collective on the
data path; torch.distributed is used only to time (barrier, max over ranks) and, in the tests,
to gather the per-rank outputs back into block order."""
from __future__ import annotations

from typing import Any, List, Sequence


def shard_block_indices(n_blocks: int, rank: int, world: int) -> List[int]:
    if world < 1 or not 0 <= rank < world:
        raise ValueError(f"bad rank/world {rank}/{world}")
    return list(range(rank, n_blocks, world))


def restore_order(per_rank: Sequence[Sequence[Any]], n_blocks: int) -> List[Any]:
    """Inverse of shard_block_indices: per_rank[r][k] is the result of block r + k*world."""
    world = len(per_rank)
    out: List[Any] = [None] * n_blocks
    for r, items in enumerate(per_rank):
        idx = shard_block_indices(n_blocks, r, world)
        if len(idx) != len(items):
            raise ValueError(f"rank {r} returned {len(items)} blocks, expected {len(idx)}")
        for i, item in zip(idx, items):
            out[i] = item
    return out


def max_over_ranks(value: float, device=None) -> float:
    """Time-like scalars are reported as the max over ranks (device-timed, never wall clock)."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device or "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0])
