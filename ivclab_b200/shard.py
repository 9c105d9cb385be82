"""Multi-GPU sharding of independent units.

Images, sequences and GOPs are independent, so the path shards with NO data-path collective:
rank ``r`` of ``world`` owns a contiguous range of units, runs the kernels on its own GPU and
stream, and the host gathers the (small) per-unit results -- scan-index tensors, motion vectors,
PSNR scalars -- in unit order.  The gather uses ``torch.distributed`` object collectives (NCCL
process group on GPUs, gloo in the CPU tests); NVLink is not on the data path by design."""
from __future__ import annotations

from typing import Callable, List, Sequence, Tuple

import torch
import torch.distributed as dist

__all__ = ["shard_range", "shard_round_robin", "run_sharded", "gather_in_order"]


def shard_range(n_units: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous ``[lo, hi)`` of ``n_units`` for ``rank`` (cfg3: frame ranges); sizes differ by <= 1."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world {rank}/{world}")
    base, rem = divmod(n_units, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_round_robin(n_units: int, rank: int, world: int) -> List[int]:
    """Unit ``s`` -> rank ``s mod world`` (cfg4: one sequence per GPU, serial when world < n)."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world {rank}/{world}")
    return list(range(rank, n_units, world))


def gather_in_order(local: Sequence, owners: Sequence[int], n_units: int, dst: int = 0):
    """Host gather: every rank passes ``local`` results for the unit ids ``owners``; rank ``dst``
    gets the full list in unit order (others get None).  Works without an initialised process
    group (world == 1)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        out = [None] * n_units
        for u, v in zip(owners, local):
            out[u] = v
        return out
    payload = [(int(u), v.cpu() if isinstance(v, torch.Tensor) else v) for u, v in zip(owners, local)]
    bucket = [None] * dist.get_world_size() if dist.get_rank() == dst else None
    dist.gather_object(payload, bucket, dst=dst)
    if dist.get_rank() != dst:
        return None
    out = [None] * n_units
    for part in bucket:
        for u, v in part:
            out[u] = v
    return out


def run_sharded(n_units: int, work: Callable[[int], object], contiguous: bool = True, dst: int = 0):
    """Run ``work(unit)`` for this rank's units and gather the results on ``dst`` in unit order."""
    rank = dist.get_rank() if dist.is_available() and dist.is_initialized() else 0
    world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
    if contiguous:
        lo, hi = shard_range(n_units, rank, world)
        mine = list(range(lo, hi))
    else:
        mine = shard_round_robin(n_units, rank, world)
    return gather_in_order([work(u) for u in mine], mine, n_units, dst)
