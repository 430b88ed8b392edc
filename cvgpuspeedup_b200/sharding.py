"""Multi-GPU use of the fused path: shard crops across ranks, optional gather into one contiguous tensor.

Crops are independent (one batch plane each, reference batch_operations.cuh:222-229) and the computation has no
exchange step, so the only collective is the optional final gather (BASELINE config 5):

  * rank r processes the contiguous crop range shard_range(n, r, world) -- contiguous so that its output is one
    contiguous [hi-lo, 3, H, W] slab;
  * every rank allocates the full [n, 3, H, W] tensor and writes its slab IN PLACE at plane offset lo (the fused
    kernel takes the destination pointer, so no staging copy exists);
  * gather_slabs() completes the tensor on every rank with ONE in-place all_gather_into_tensor (NCCL over
    NVLink/NVSwitch) when n divides evenly, else with one broadcast per rank;
  * or no collective at all: PeerTensor places the full tensor of every rank in memory the other ranks map into
    their own process (CUDA IPC), and the fused kernel stores each plane it computes into every rank's tensor
    (cvgs_b200_preproc_launch_replicated): the gather rides on the kernel's stores over NVLink while it computes;
    one tiny stream-ordered all-reduce afterwards tells every rank that all slabs have landed.

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


class PeerTensor:
    """One float32 CUDA tensor of the same shape per rank, each mapped into every other rank's address space.

    tensor   this rank's tensor (a torch view of library-allocated device memory)
    peers    device pointers, valid in THIS process, of the other ranks' tensors (rank order, own rank skipped)"""

    def __init__(self, shape, group=None):
        import ctypes as C
        import math
        import torch
        import torch.distributed as dist
        from . import _abi
        self._lib = _abi.load()
        self.shape = tuple(int(v) for v in shape)
        nbytes = 4 * math.prod(self.shape)
        ptr = C.c_void_p()
        _abi.check(self._lib.cvgs_b200_dev_alloc(C.byref(ptr), nbytes))
        self.ptr = int(ptr.value)
        self.peers, self._opened = [], []
        world = dist.get_world_size(group) if dist.is_initialized() else 1
        rank = dist.get_rank(group) if dist.is_initialized() else 0
        if world > 1:
            handle = C.create_string_buffer(64)
            _abi.check(self._lib.cvgs_b200_ipc_export(self.ptr, handle))
            handles = [None] * world
            dist.all_gather_object(handles, handle.raw, group=group)
            for r, h in enumerate(handles):
                if r == rank:
                    continue
                p = C.c_void_p()
                _abi.check(self._lib.cvgs_b200_ipc_open(C.create_string_buffer(h, 64), C.byref(p)))
                self.peers.append(int(p.value))
                self._opened.append(int(p.value))

        class _Holder:
            pass

        h = _Holder()
        h.__cuda_array_interface__ = {"shape": self.shape, "typestr": "<f4", "data": (self.ptr, False), "version": 2,
                                      "strides": None}
        self.tensor = torch.as_tensor(h, device="cuda")

    def close(self, group=None) -> None:
        """Unmap the peers' tensors, wait until every rank has done so, free this rank's."""
        import torch
        import torch.distributed as dist
        torch.cuda.synchronize()
        for p in self._opened:
            self._lib.cvgs_b200_ipc_close(p)
        self._opened, self.peers = [], []
        if dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.barrier(group=group)
        if self.ptr:
            self.tensor = None
            self._lib.cvgs_b200_dev_free(self.ptr)
            self.ptr = 0


def executeOperationsShardedFused(stream, crops: Sequence[api.GpuMat], dsize, ops, full: PeerTensor, group=None,
                                  signal=None, **kw) -> Tuple[int, int]:
    """The sharded launch with the gather fused into the kernel: this rank's planes are stored into its own tensor
    and, over NVLink, into every peer's (full.peers) by the same launch.  `signal` (a 1-element CUDA tensor) is
    all-reduced on the current stream afterwards: when that completes, every rank's kernel has finished, i.e. every
    rank's tensor is whole.  Returns this rank's (lo, hi)."""
    import torch
    import torch.distributed as dist
    n = len(crops)
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    lo, hi = shard_range(n, rank, world)
    plane_bytes = 4 * 3 * int(dsize[0]) * int(dsize[1])
    if hi > lo:
        api.executeOperations(stream, api.resize(list(crops[lo:hi]), dsize, hi - lo), *ops,
                              api.split(full.tensor[lo:hi], dsize), replicas=[p + lo * plane_bytes for p in full.peers], **kw)
    if world > 1:
        cur = torch.cuda.current_stream()
        _order_after(cur, stream)
        if signal is None:
            signal = torch.zeros(1, device="cuda")
        dist.all_reduce(signal, group=group)
    return lo, hi


def _order_after(cur, stream) -> None:
    """Make the current torch stream wait for `stream` (a torch stream, a raw cudaStream_t handle, or None)."""
    import torch
    if stream is None:
        return
    if isinstance(stream, int):
        stream = torch.cuda.ExternalStream(stream)
    if stream != cur:
        cur.wait_stream(stream)


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
        _order_after(torch.cuda.current_stream(), stream)
        gather_slabs(out_full, n, group)
    return lo, hi
