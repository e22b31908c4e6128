"""Multi-GPU layout of the all-pairs path: one process per GPU, targets partitioned,
sources replicated.

Every output element is an independent sum over ALL sources (reference
src/P3D.cpp:335-339), so the only data that ever has to move between GPUs is the
source set when it arrives sharded (each rank owning ``n/world`` of the particles):
one all-gather of the raw rows over NCCL / NVLink per step, after which each rank
runs the unchanged single-GPU kernel on its own contiguous target range and keeps
its slice of the result.  No collective touches the results.

The reference itself has no multi-device strategy ("You can enable multiple
accelerators, but right now it doesn't do anything useful", reference README.md:149-150;
every host wrapper hard-codes active device 0, src/ocl_P3D.cpp:51-52).

``torch.distributed`` is plumbing here (process group, NCCL all-gather); on CPU the
same code runs over ``gloo`` for the tests.
"""
from __future__ import annotations

from typing import List, Tuple


def target_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced, exhaustive: rank r owns [n*r/world, n*(r+1)/world)."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world of {world}")
    return (n * rank) // world, (n * (rank + 1)) // world


def all_ranges(n: int, world: int) -> List[Tuple[int, int]]:
    return [target_range(n, r, world) for r in range(world)]


def shard_rows(rows, rank: int, world: int):
    """The rows of `rows` (numpy array or torch tensor) this rank owns."""
    lo, hi = target_range(rows.shape[0], rank, world)
    return rows[lo:hi]


def allgather_rows(local, n_total: int, group=None):
    """All-gather row shards laid out by :func:`target_range` into the full (n_total, cols)
    tensor, in rank order.  Shards may differ by one row; they are padded to a common
    length for ``all_gather_into_tensor`` and the padding is dropped on arrival."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    if world == 1:
        return local
    cols = local.shape[1]
    ranges = all_ranges(n_total, world)
    lo, hi = ranges[rank]
    if local.shape[0] != hi - lo:
        raise ValueError(f"rank {rank} holds {local.shape[0]} rows, expected {hi - lo}")
    longest = max(b - a for a, b in ranges)
    send = local
    if local.shape[0] != longest:
        send = torch.zeros((longest, cols), dtype=local.dtype, device=local.device)
        send[: hi - lo] = local
    recv = torch.empty((world * longest, cols), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(recv, send.contiguous(), group=group)
    if all(b - a == longest for a, b in ranges):
        return recv
    parts = [recv[r * longest: r * longest + (b - a)] for r, (a, b) in enumerate(ranges)]
    return torch.cat(parts, dim=0)


class ShardedM2M:
    """One rank's view of an all-pairs step on device-resident particles.

    ``step(op, reg, src_local, tgt_local, out_local, ...)``: all-gather the source rows,
    then ``cvtx_b200_m2m`` on this rank's targets.  ``backend`` is a
    :class:`cvortex_b200.device.DeviceBackend`; tensors are CUDA tensors of this
    rank's device."""

    def __init__(self, backend, device: int, n_src_total: int, group=None):
        self.backend, self.device, self.n_src_total, self.group = backend, device, n_src_total, group

    def gather_sources(self, src_local):
        import torch.distributed as dist
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(self.group) == 1:
            return src_local
        return allgather_rows(src_local, self.n_src_total, self.group)

    def step(self, op: str, reg: str, src_local, tgt_local, out_local, sigma: float, nu: float = 0.0,
             stream=None, src_full=None):
        import torch
        full = self.gather_sources(src_local) if src_full is None else src_full
        st = torch.cuda.current_stream().cuda_stream if stream is None else stream
        self.backend.m2m(op, reg, self.device, st, full, full.shape[0], tgt_local, tgt_local.shape[0],
                         out_local, sigma, nu)
        return full
