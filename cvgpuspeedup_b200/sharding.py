"""Multi-GPU use of the fused path: shard crops across ranks, optional gather into one contiguous tensor.

Crops are independent (one batch plane each, reference batch_operations.cuh:222-229) and the computation has no
exchange step, so the only collective is the optional final gather (BASELINE config 5):

  * rank r processes the contiguous crop range shard_range(n, r, world) -- contiguous so that its output is one
    contiguous [hi-lo, 3, H, W] slab;
  * every rank allocates the full [n, 3, H, W] tensor and writes its slab IN PLACE at plane offset lo (the fused
    kernel takes the destination pointer, so no staging copy exists);
  * gather_slabs() completes the tensor on every rank with ONE in-place all_gather_into_tensor (NCCL over
    NVLink/NVSwitch) when n divides evenly, else with one broadcast per rank.

One process per GPU (torchrun); torch.distributed is plumbing only.  The host logic runs under gloo on CPU
(tests/test_sharding.py); the kernel itself has no CPU fallback.
"""
from __future__ import annotations

from typing import Sequence, Tuple

from . import api


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced range of crop indices owned by `rank` (sizes differ by at most one)."""
    if world <= 0 or not 0 <= rank < world or n < 0:
        raise ValueError("bad shard arguments")
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def gather_slabs(out_full, n: int, group=None) -> None:
    """Complete `out_full` ([n, ...], same shape on every rank) from the per-rank slabs written in place."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    if world == 1:
        return
    if n % world == 0:
        lo, hi = shard_range(n, rank, world)
        dist.all_gather_into_tensor(out_full, out_full[lo:hi], group=group)  # in place: input aliases its slot
        return
    for r in range(world):  # uneven split: one broadcast per owner
        lo, hi = shard_range(n, r, world)
        if hi > lo:
            dist.broadcast(out_full[lo:hi], src=dist.get_global_rank(group, r) if group is not None else r, group=group)


def executeOperationsSharded(stream, crops: Sequence[api.GpuMat], dsize, ops, out_full, gather: bool = True, group=None,
                             **kw) -> Tuple[int, int]:
    """cvGS::executeOperations for a crop list sharded over the ranks of `group`.

    crops     the FULL crop list (the source image is resident on every GPU); only this rank's range is read
    out_full  [len(crops), 3, H, W] float32 CUDA tensor; this rank's planes are written in place
    Returns this rank's (lo, hi).  With gather=True every rank ends with the complete tensor."""
    import torch.distributed as dist
    n = len(crops)
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    lo, hi = shard_range(n, rank, world)
    if hi > lo:
        api.executeOperations(stream, api.resize(list(crops[lo:hi]), dsize, hi - lo), *ops,
                              api.split(out_full[lo:hi], dsize), **kw)
    if gather and world > 1:
        # the collective runs on the current stream of `out_full`'s device; order it after the kernel
        import torch
        cur = torch.cuda.current_stream()
        if isinstance(stream, torch.cuda.Stream) and stream != cur:
            cur.wait_stream(stream)
        gather_slabs(out_full, n, group)
    return lo, hi
