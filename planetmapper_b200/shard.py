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


def bind_rank_to_cores(local_rank: int, local_world: int, device_index: int | None = None) -> dict:
    """Give every rank of a node its own slice of the host cores (and, on multi-socket hosts, the cores
    of the NUMA node its GPU hangs off) BEFORE it allocates pinned memory, so that staging buffers are
    first touched - hence placed - next to the GPU and the ranks' copy / hand-off threads do not
    migrate across each other.  Returns what was done (for the bench line).

    The GPU's NUMA node comes from /sys/bus/pci/devices/<bdf>/numa_node; -1 (virtual machines, single
    socket) means no placement information, and the cores are simply split evenly."""
    info = {'local_rank': local_rank, 'local_world': local_world, 'numa_node': None}
    try:
        allowed = sorted(os.sched_getaffinity(0))
    except AttributeError:   # pragma: no cover
        return info
    cores = allowed
    if device_index is not None:
        try:
            import torch

            bdf = torch.cuda.get_device_properties(device_index).pci_bus_id   # torch >= 2.x
        except Exception:
            bdf = None
        if bdf:
            try:
                with open(f'/sys/bus/pci/devices/{str(bdf).lower()}/numa_node') as f:
                    node = int(f.read().strip())
                info['numa_node'] = node
                if node >= 0:
                    with open(f'/sys/devices/system/node/node{node}/cpulist') as f:
                        node_cores = _parse_cpulist(f.read())
                    local = [c for c in allowed if c in node_cores]
                    if local:
                        cores = local
            except (OSError, ValueError):
                pass
    per = max(1, len(cores) // max(1, local_world))
    mine = cores[(local_rank % max(1, len(cores) // per)) * per:][:per] or cores
    try:
        os.sched_setaffinity(0, mine)
        info['cores'] = mine
    except OSError:
        info['cores'] = allowed
    return info


def _parse_cpulist(text: str) -> set[int]:
    out: set[int] = set()
    for part in text.strip().split(','):
        if not part:
            continue
        lo, _, hi = part.partition('-')
        out.update(range(int(lo), int(hi or lo) + 1))
    return out


def gather_shard_results(obj, world_size: int):
    """all_gather of small python objects (checksums, timings); identity if W == 1."""
    if world_size == 1:
        return [obj]
    import torch.distributed as dist

    out = [None] * world_size
    dist.all_gather_object(out, obj)
    return out


def gather_blocks(local, n_units: int, world_size: int, rank: int, dst: int = 0, out=None):
    """Assemble the contiguous per-rank blocks of :func:`shard_range` into ONE tensor on rank `dst`
    ("cudaMemcpyPeer over NVLink 5 if the caller wants one device buffer").

    ``local``: this rank's block, shape (hi - lo, ...) for ``lo, hi = shard_range(n_units, rank, W)``.
    Returns the (n_units, ...) tensor on rank ``dst`` (``out`` if given), None elsewhere.  Point-to-point
    sends straight into the destination slices - with the NCCL backend that is a GPU-to-GPU copy
    over NVLink, with no staging buffer and no padding of unequal blocks.  This is result assembly
    after the path has run; the path itself has no exchange step.
    """
    import torch
    import torch.distributed as dist

    if world_size == 1:
        if out is None:
            return local
        out.copy_(local)
        return out
    if rank == dst:
        if out is None:
            out = torch.empty((n_units,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        ops = []
        for src in range(world_size):
            lo, hi = shard_range(n_units, src, world_size)
            if src == dst:
                out[lo:hi].copy_(local)
            elif hi > lo:
                ops.append(dist.P2POp(dist.irecv, out[lo:hi], src))
        if ops:
            for req in dist.batch_isend_irecv(ops):   # batched: the receives run concurrently
                req.wait()
        return out
    if local.shape[0] > 0:
        for req in dist.batch_isend_irecv([dist.P2POp(dist.isend, local.contiguous(), dst)]):
            req.wait()
    return None


def max_over_ranks(value: float, world_size: int, device=None) -> float:
    if world_size == 1:
        return float(value)
    import torch
    import torch.distributed as dist

    t = torch.tensor([float(value)], dtype=torch.float64, device=device or 'cpu')
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
