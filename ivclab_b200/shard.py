"""Multi-GPU sharding of independent units.

Images, sequences and GOPs are independent, so the path shards with NO data-path collective:
rank ``r`` of ``world`` owns a contiguous range of units, runs the kernels on its own GPU and
stream, and the host gathers the (small) per-unit results -- symbol statistics, motion vectors,
PSNR scalars -- in unit order.  Bulk results travel as tensors (``gather_rows``: one ``dist.gather`` of
equal-sized blocks, device tensors under NCCL, host tensors under gloo), odd-shaped Python results as
objects (``gather_in_order``).  The gather is the only exchange; nothing on the coding path uses a collective."""
from __future__ import annotations

from typing import Callable, List, Sequence, Tuple

import torch
import torch.distributed as dist

__all__ = ["shard_range", "shard_round_robin", "run_sharded", "gather_in_order", "gather_rows", "SharedPinned"]


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


def gather_rows(local: torch.Tensor, n_units: int, dst: int = 0, axis: int = 0, out: torch.Tensor = None):
    """Host gather for contiguous-range sharding (``shard_range``): rank r holds the slice ``[lo_r, hi_r)`` of
    ``axis`` of a ``[.., n_units, ..]`` array; rank ``dst`` gets the whole array in unit order (others None).
    One ``dist.gather`` of tensors -- no pickling: every rank contributes a block padded to the largest shard
    (shards differ by at most one unit).  Under NCCL ``local`` is a CUDA tensor and the result arrives on ``dst``'s
    device; pass a pinned host tensor as ``out`` to have it copied there (the device-to-host read that ends the
    gather).  Without an initialised process group the input is returned (copied into ``out`` if given)."""
    on = dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
    if not on:
        if local.shape[axis] != n_units:
            raise ValueError(f"single process: the local shard must hold all {n_units} units, got {local.shape[axis]}")
        if out is not None:
            out.copy_(local)
            return out
        return local
    rank, world = dist.get_rank(), dist.get_world_size()
    lo, hi = shard_range(n_units, rank, world)
    if local.shape[axis] != hi - lo:
        raise ValueError(f"rank {rank} owns units [{lo}, {hi}) but holds {local.shape[axis]} rows")
    per = (n_units + world - 1) // world
    x = local.movedim(axis, 0).contiguous()
    if x.shape[0] < per:                                        # pad to the common block size
        pad = torch.zeros((per - x.shape[0],) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
        x = torch.cat([x, pad], 0)
    bucket = [torch.empty_like(x) for _ in range(world)] if rank == dst else None
    dist.gather(x, bucket, dst=dst)
    if rank != dst:
        return None
    parts = []
    for r in range(world):
        a, b = shard_range(n_units, r, world)
        parts.append(bucket[r][:b - a])
    full = torch.cat(parts, 0).movedim(0, axis)
    if out is not None:
        out.copy_(full, non_blocking=False)
        return out
    return full


class SharedPinned:
    """A host array in POSIX shared memory that every rank of a one-node run maps and page-locks: each rank's
    device-to-host copies of ITS units land directly in the array rank ``dst`` reads -- the "host gather of bitstream
    symbols and PSNR" of the path with no collective and no second copy (one process per GPU, one node).

    Collective constructor: every rank calls ``SharedPinned(shape, dtype, tag)``; rank ``dst`` creates the segment
    under ``/dev/shm``, the others attach after a barrier.  ``.tensor`` is a CPU tensor over the segment (pinned via
    ``cudaHostRegister`` when a CUDA device is in use; plain shared memory otherwise, e.g. in the gloo tests).
    ``close()`` (collective) unmaps and removes it.  Without a process group it is a private pinned buffer."""

    def __init__(self, shape, dtype, tag: str, dst: int = 0, register: bool = True):
        import mmap
        import os
        on = dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
        self._rank = dist.get_rank() if on else 0
        self._dst, self._on = dst, on
        self._registered = False
        nbytes = int(torch.tensor([], dtype=dtype).element_size())
        for d in shape:
            nbytes *= int(d)
        self._nbytes = max(nbytes, 1)
        name = [f"ivcb200_{os.getpid()}_{tag}"]
        if on:
            dist.broadcast_object_list(name, src=dst)
        self._path = os.path.join("/dev/shm", name[0])
        # no step raises before both barriers have been passed: a rank that fails must not leave the others waiting
        err, fd, self._map, self.tensor = None, -1, None, None
        try:
            if self._rank == dst:
                fd = os.open(self._path, os.O_CREAT | os.O_RDWR | os.O_TRUNC, 0o600)
                os.ftruncate(fd, self._nbytes)
        except Exception as e:
            err = e
        if on:
            dist.barrier()
        try:
            if err is None:
                if self._rank != dst:
                    fd = os.open(self._path, os.O_RDWR)
                self._map = mmap.mmap(fd, self._nbytes)
                os.close(fd)
                esz = torch.tensor([], dtype=dtype).element_size()
                self.tensor = torch.frombuffer(self._map, dtype=dtype, count=nbytes // esz).reshape(tuple(shape))
                if register and torch.cuda.is_available():
                    rc = torch.cuda.cudart().cudaHostRegister(self.tensor.data_ptr(), self._nbytes, 0)
                    if int(rc) != 0:
                        raise RuntimeError(f"cudaHostRegister failed on the shared segment: {rc}")
                    self._registered = True
        except Exception as e:
            err = e
        if on:
            dist.barrier()
        if err is not None:
            if self._rank == dst:
                try:
                    os.unlink(self._path)
                except OSError:
                    pass
            raise err

    def close(self):
        import os
        if self._on:
            dist.barrier()
        if self._registered:
            torch.cuda.cudart().cudaHostUnregister(self.tensor.data_ptr())
            self._registered = False
        self.tensor = None
        try:
            if self._map is not None:
                self._map.close()
        except BufferError:
            pass                                  # a caller still holds a view; the mapping goes with the process
        if self._rank == self._dst:
            try:
                os.unlink(self._path)
            except FileNotFoundError:
                pass
