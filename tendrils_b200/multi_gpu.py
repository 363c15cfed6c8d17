"""Multi-GPU flow exchange: the ordered ring fold.

Particles are sharded by contiguous column blocks, i.e. by contiguous ranges of the draw order
p = x*PH + y.  The reference's flow blend is ordered and non-commutative (alpha "over" in
primitive order, src/index.js:267-268), so the ranks cannot simply sum their grids: the grid
travels rank 0 -> 1 -> ... -> P-1, each rank folding its own ordered fragment lists onto what
it received, and the last rank broadcasts the result.  This equals the single-GPU result bit
for bit.  torch.distributed (NCCL over NVLink on GPUs, gloo in the CPU tests) is only the wire.
"""
from __future__ import annotations

import contextlib


def ordered_ring_fold(rank, world_size, group, fold, flow_tensor, stream=None):
    """fold(): blend this rank's collected fragments onto the local flow grid.
    flow_tensor(): a torch tensor aliasing the local flow grid (device or CPU)."""
    import torch
    import torch.distributed as dist

    ctx = contextlib.nullcontext()
    if stream:
        ctx = torch.cuda.stream(torch.cuda.ExternalStream(stream))
    with ctx:
        t = flow_tensor()
        if rank > 0:
            dist.recv(t, src=_global_rank(group, rank - 1), group=group)
        fold()
        if rank < world_size - 1:
            dist.send(t, dst=_global_rank(group, rank + 1), group=group)
        dist.broadcast(t, src=_global_rank(group, world_size - 1), group=group)


def _global_rank(group, group_rank):
    import torch.distributed as dist
    if group is None:
        return group_rank
    return dist.get_global_rank(group, group_rank)


def gather_handles(mine: bytes, world_size, group, device):
    """all-gather the ranks' IPC handle blobs (rank order).  torch.distributed is only the courier."""
    import torch
    import torch.distributed as dist
    use_cuda = dist.get_backend(group) == "nccl"
    dev = torch.device("cuda", device) if use_cuda else torch.device("cpu")
    t = torch.tensor(list(mine), dtype=torch.uint8, device=dev)
    out = [torch.empty_like(t) for _ in range(world_size)]
    dist.all_gather(out, t, group=group)
    return [bytes(o.cpu().tolist()) for o in out]


class _CudaArray:
    def __init__(self, ptr, n, typestr="<f4"):
        self.__cuda_array_interface__ = {
            "shape": (n,), "typestr": typestr, "data": (ptr, False), "version": 3, "strides": None}


def wrap_device_buffer(ptr, n, device, typestr="<f4"):
    """A torch tensor aliasing `n` 4-byte elements of device memory owned by the C library."""
    import torch
    if n == 0:
        return torch.empty(0, dtype=torch.float32 if typestr == "<f4" else torch.int32, device=torch.device("cuda", device))
    return torch.as_tensor(_CudaArray(ptr, n, typestr), device=torch.device("cuda", device))
