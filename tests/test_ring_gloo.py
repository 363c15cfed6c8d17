"""world_size-2 gloo test of the multi-GPU host logic: column sharding + the ordered ring fold
(tendrils_b200/multi_gpu.py), with the CPU oracle standing in for the per-rank device work."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DT = 1000 / 60


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, R, G, steps, out_dir):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from oracle import oracle as O
    from tendrils_b200.multi_gpu import ordered_ring_fold
    from tendrils_b200.tendrils import shard_columns
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    cols = shard_columns(R, rank, world)
    P = O.make_params()
    cur, prev = O.spawn_ball(R, R, 0.3, 0.005, cols=cols), O.spawn_init(R, R)
    targets, flow = np.zeros((R, R, 4), np.float32), np.zeros((G, G, 4), np.float32)
    flow_t = torch.from_numpy(flow)              # aliases `flow`
    t = DT
    for _ in range(steps):
        t += DT
        new = O.integrate(P, cur, targets, flow, np.float32(t), np.float32(DT), cols=cols)   # local shard only
        prev, cur = cur, new
        ordered_ring_fold(rank, world, None,
                          fold=lambda: O.splat(P, cur, prev, flow, np.float32(t), cols=cols),
                          flow_tensor=lambda: flow_t)
    np.save(os.path.join(out_dir, f"flow_{rank}.npy"), flow)
    np.save(os.path.join(out_dir, f"state_{rank}.npy"), cur[cols[0]:cols[1]])
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_ring_fold_equals_single_process(oracle, tmp_path, world):
    import torch.multiprocessing as mp
    R, G, steps = 48, 32, 6
    port = _free_port()
    mp.spawn(_worker, args=(world, port, R, G, steps, str(tmp_path)), nprocs=world, join=True)
    O = oracle
    P = O.make_params()
    cur, prev = O.spawn_ball(R, R, 0.3, 0.005), O.spawn_init(R, R)
    targets, flow = np.zeros((R, R, 4), np.float32), np.zeros((G, G, 4), np.float32)
    t = DT
    for _ in range(steps):
        t += DT
        new = O.integrate(P, cur, targets, flow, np.float32(t), np.float32(DT))
        prev, cur = cur, new
        O.splat(P, cur, prev, flow, np.float32(t))
    flows = [np.load(tmp_path / f"flow_{r}.npy") for r in range(world)]
    for r in range(world):
        assert np.array_equal(flows[r], flow), f"rank {r} flow differs from the single-process result"
    got = np.concatenate([np.load(tmp_path / f"state_{r}.npy") for r in range(world)], 0)
    assert np.array_equal(got, cur)


# ---- the parallel band fold ("bands"): its host-side protocol on gloo, the oracle standing in for the kernels -------
def _bands_worker(rank, world, port, R, G, steps, out_dir):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from oracle import oracle as O
    from tendrils_b200.multi_gpu import gather_handles
    from tendrils_b200.tendrils import shard_columns
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    # the handle exchange of Particles._ensure_bands: every rank ends up with every rank's blob, in rank order
    blobs = gather_handles(bytes([rank, 7, 255 - rank]) * 5, world, None, 0)
    assert blobs == [bytes([r, 7, 255 - r]) * 5 for r in range(world)]
    cols = shard_columns(R, rank, world)
    P = O.make_params()
    cur, prev = O.spawn_ball(R, R, 0.3, 0.005, cols=cols), O.spawn_init(R, R)
    targets, flow = np.zeros((R, R, 4), np.float32), np.zeros((G, G, 4), np.float32)
    tiles = (G * G + 31) // 32
    mine = np.zeros(G * G, bool)
    for k in range(rank, tiles, world):                     # tb_splat_fold_bands: tile % world == rank
        mine[32 * k:32 * k + 32] = True
    mine = mine.reshape(G, G)
    t = DT
    for _ in range(steps):
        t += DT
        new = O.integrate(P, cur, targets, flow, np.float32(t), np.float32(DT), cols=cols)
        prev, cur = cur, new
        # "push": every owner receives every source's primitives (here: the source's state shard; the kernels send
        # the rasterised fragments) and folds them onto ITS tiles in source-rank order
        pair = torch.from_numpy(np.stack([cur, prev]))
        shards = [torch.empty_like(pair) for _ in range(world)]
        dist.all_gather(shards, pair)
        band = flow.copy()
        for src in range(world):
            c, p = shards[src][0].numpy(), shards[src][1].numpy()
            O.splat(P, c, p, band, np.float32(t), cols=shard_columns(R, src, world))
        # "publish": every rank's finished tiles go into every rank's grid
        out = torch.from_numpy(np.where(mine[..., None], band, 0).astype(np.float32))
        dist.all_reduce(out)                                # tiles are disjoint: the sum is a gather
        flow[...] = out.numpy()
    np.save(os.path.join(out_dir, f"flow_{rank}.npy"), flow)
    np.save(os.path.join(out_dir, f"state_{rank}.npy"), cur[cols[0]:cols[1]])
    dist.destroy_process_group()


@pytest.mark.parametrize("world,G", [(2, 32), (3, 25)])
def test_band_fold_protocol_equals_single_process(oracle, tmp_path, world, G):
    import torch.multiprocessing as mp
    R, steps = 48, 5
    mp.spawn(_bands_worker, args=(world, _free_port(), R, G, steps, str(tmp_path)), nprocs=world, join=True)
    O = oracle
    P = O.make_params()
    cur, prev = O.spawn_ball(R, R, 0.3, 0.005), O.spawn_init(R, R)
    targets, flow = np.zeros((R, R, 4), np.float32), np.zeros((G, G, 4), np.float32)
    t = DT
    for _ in range(steps):
        t += DT
        new = O.integrate(P, cur, targets, flow, np.float32(t), np.float32(DT))
        prev, cur = cur, new
        O.splat(P, cur, prev, flow, np.float32(t))
    for r in range(world):
        assert np.array_equal(np.load(tmp_path / f"flow_{r}.npy"), flow), f"rank {r} flow"     # (the sum may turn -0 into +0)
    got = np.concatenate([np.load(tmp_path / f"state_{r}.npy") for r in range(world)], 0)
    assert np.array_equal(got, cur)
