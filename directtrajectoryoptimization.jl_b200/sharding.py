"""Batch partitioning across GPUs / ranks. Problems are independent, so the only multi-GPU logic is
a contiguous split: shard i owns problems [i*ceil(B/n), min(B, (i+1)*ceil(B/n))) -- the same rule
dto_batch_create applies to its device list (csrc/dto_runtime.cpp). No collective is involved on
the data path; ranks only combine timing statistics."""
from __future__ import annotations

from typing import List, Tuple


def partition(total: int, shards: int) -> List[Tuple[int, int]]:
    """[(begin, size)] per shard; trailing shards may be empty."""
    if total < 0 or shards < 1:
        raise ValueError("partition: need total >= 0 and shards >= 1")
    chunk = (total + shards - 1) // shards
    out = []
    for i in range(shards):
        b = min(total, i * chunk)
        out.append((b, min(chunk, total - b)))
    return out


def rank_slice(total: int, rank: int, world: int) -> slice:
    b, n = partition(total, world)[rank]
    return slice(b, b + n)


def reduce_max(value: float, device: str = "cpu") -> float:
    """max over ranks of a scalar (timings); identity when torch.distributed is not initialised."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()):
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
