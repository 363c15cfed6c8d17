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


def exchange_ring_handles(mine: bytes, rank, world_size, group, device):
    """all-gather the ranks' IPC handle blobs and return the one of rank+1 (the ring successor)."""
    return gather_handles(mine, world_size, group, device)[(rank + 1) % world_size]


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


def band_exchange_fold(particles):
    """The scalable ordered fold of a column-sharded run ("a2a").

    The grid is cut into one band of texels per rank.  Every rank's sorted fragment array is, band by band,
    a sequence of contiguous slices; one all-to-all delivers each slice to the band's owner, which blends the
    pieces it received in source-rank order (= draw order, the shards being contiguous column blocks) and the
    bands are all-gathered.  No rank waits for another rank's fold, unlike the ring; the result is the same,
    bit for bit.  NCCL over NVLink carries the slices; the folds are the library's own kernels."""
    import ctypes as C
    import torch
    import torch.distributed as dist
    from . import _native as N

    P, L, ctx, gl = particles, particles._L, particles._ctx, particles.gl
    world, rank, dev = gl.world_size, gl.rank, gl.device
    G = P.flow_shape[0] * P.flow_shape[1]
    band = G // world
    offs = (C.c_int64 * (world + 1))()
    N.check(ctx, L.tb_splat_band_offsets(ctx, world, band, offs))
    send_counts = [int(offs[b + 1] - offs[b]) for b in range(world)]
    n_send = int(offs[world])
    stream = torch.cuda.ExternalStream(P.stream_handle(), device=torch.device("cuda", dev))
    with torch.cuda.stream(stream):
        sc = torch.tensor(send_counts, dtype=torch.int64, device=torch.device("cuda", dev))
        rc = torch.empty_like(sc)
        dist.all_to_all_single(rc, sc, group=gl.group)
        recv_counts = [int(v) for v in rc.tolist()]
        n_recv = sum(recv_counts)
        sk, sv, rk, rv, vb = C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_int32()
        N.check(ctx, L.tb_splat_exchange_buffers(ctx, n_recv, C.byref(sk), C.byref(sv), C.byref(rk), C.byref(rv), C.byref(vb)))
        wpv = vb.value // 4                                            # 4-byte words per fragment value
        send_keys = wrap_device_buffer(sk.value, n_send, dev, "<i4")
        recv_keys = wrap_device_buffer(rk.value, n_recv, dev, "<i4")
        send_vals = wrap_device_buffer(sv.value, n_send * wpv, dev)
        recv_vals = wrap_device_buffer(rv.value, n_recv * wpv, dev)
        dist.all_to_all_single(recv_keys, send_keys, recv_counts, send_counts, group=gl.group)
        dist.all_to_all_single(recv_vals, send_vals, [c * wpv for c in recv_counts], [c * wpv for c in send_counts],
                               group=gl.group)
        off = 0
        for src in range(world):                                       # source-rank order = draw order
            N.check(ctx, L.tb_splat_fold_piece(ctx, off, recv_counts[src], rank * band, (rank + 1) * band))
            off += recv_counts[src]
        flow = particles._flow_tensor()
        dist.all_gather_into_tensor(flow, flow[rank * band * 4:(rank + 1) * band * 4], group=gl.group)
        N.check(ctx, L.tb_splat_exchange_done(ctx))
