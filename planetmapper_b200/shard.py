"""
Sharding of the hot path across GPUs.

Frames of a time series and wavelength planes of a cube are independent, so the path
shards with NO data-path collective: rank r of world size W takes one contiguous
block of units.  ``torch.distributed`` (NCCL on GPUs, gloo in the CPU tests) is used
only for the barrier + max-over-ranks timing and for gathering per-shard checksums /
host results.
"""
from __future__ import annotations

import os


def shard_range(n_units: int, rank: int, world_size: int) -> tuple[int, int]:
    """Contiguous [start, stop) block of n_units for `rank`; sizes differ by <= 1."""
    if world_size <= 0 or not 0 <= rank < world_size:
        raise ValueError(f'bad rank/world_size {rank}/{world_size}')
    base, extra = divmod(int(n_units), world_size)
    start = rank * base + min(rank, extra)
    stop = start + base + (1 if rank < extra else 0)
    return start, stop


def env_rank_world() -> tuple[int, int, int]:
    """(rank, local_rank, world_size) from the torchrun environment (defaults 0, 0, 1)."""
    return (int(os.environ.get('RANK', 0)), int(os.environ.get('LOCAL_RANK', 0)),
            int(os.environ.get('WORLD_SIZE', 1)))


def gather_shard_results(obj, world_size: int):
    """all_gather of small python objects (checksums, timings); identity if W == 1."""
    if world_size == 1:
        return [obj]
    import torch.distributed as dist

    out = [None] * world_size
    dist.all_gather_object(out, obj)
    return out


def max_over_ranks(value: float, world_size: int, device=None) -> float:
    if world_size == 1:
        return float(value)
    import torch
    import torch.distributed as dist

    t = torch.tensor([float(value)], dtype=torch.float64, device=device or 'cpu')
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
