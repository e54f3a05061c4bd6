"""Row sharding across the GPUs of one box (one process per GPU, torch.distributed).

Spectra are independent, so the fit has NO data-path collective: rank g owns the contiguous
row block shard_rows(R, world, g); the only optional exchange is a final gather
of the small per-row outputs (coeff 8*K B, rms 4 B, status 4+4 B per row) to one rank.
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple


def shard_rows(num_rows: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Rows [start, stop) of `rank`: blocks of ceil(R/G) rows, the last ones possibly shorter/empty."""
    if world_size < 1 or not (0 <= rank < world_size):
        raise ValueError("bad world_size / rank")
    per = -(-num_rows // world_size)
    start = min(num_rows, rank * per)
    stop = min(num_rows, start + per)
    return start, stop


def gather_row_outputs(outputs: Dict[str, "object"], num_rows: int, dst: int = 0,
                       group=None) -> Optional[Dict[str, "object"]]:
    """Gathers per-row tensors (first dimension = this rank's rows) onto rank `dst` in global
    row order. Works with any backend (nccl on GPUs, gloo on CPU). Returns None on other ranks."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    per = -(-num_rows // world)
    result = {} if rank == dst else None
    for key, t in outputs.items():
        if t is None:
            continue
        pad_shape = (per,) + tuple(t.shape[1:])
        padded = torch.zeros(pad_shape, dtype=t.dtype, device=t.device)
        padded[: t.shape[0]] = t
        bucket = [torch.empty_like(padded) for _ in range(world)] if rank == dst else None
        dist.gather(padded, bucket, dst=dst, group=group)
        if rank == dst:
            parts = []
            for g in range(world):
                s, e = shard_rows(num_rows, world, g)
                parts.append(bucket[g][: e - s])
            result[key] = torch.cat(parts, dim=0)
    return result


def max_over_ranks(value: float, device=None, group=None) -> float:
    """Timing rule of bench.py: a multi-GPU step takes as long as its slowest rank."""
    import torch
    import torch.distributed as dist

    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())
