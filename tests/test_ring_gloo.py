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
