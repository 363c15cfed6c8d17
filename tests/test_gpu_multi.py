"""Multi-GPU parity (needs >= 2 GPUs, otherwise skipped): column-sharded contexts + the ordered ring
fold over NCCL through the public API must reproduce the single-process oracle bit for bit."""
import os
import socket
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, shape, G, steps, out_dir, ring, radius, prune):
    sys.path.insert(0, ROOT)
    if prune == "seg":
        os.environ.update({"TB_SEG_AT": "96", "TB_SEG_LEN": "64", "TB_SPLIT_AT": "512"})
    elif prune is not None:
        os.environ["TB_PRUNE"] = prune
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    import tendrils_b200 as T
    from tendrils_b200.spawn import spawnBall
    t = T.Tendrils(T.Device(G, G, device=rank, rank=rank, world_size=world, group=dist.group.WORLD, ring=ring))
    t.setup(shape); t.resize()
    spawnBall(t.gl, {"uniforms": {"radius": radius, "speed": 0.005}}).spawn(t)
    for _ in range(steps):
        t.timer.tick()
        t.step().draw()
    np.save(os.path.join(out_dir, f"flow_{rank}.npy"), t.flow.download())
    np.save(os.path.join(out_dir, f"state_{rank}.npy"), t.particles.buffers[0].download())
    np.save(os.path.join(out_dir, f"frags_{rank}.npy"), np.array([t.particles.stats()["last_fragments"], t.particles.segment_stats()["bins"]]))
    dist.destroy_process_group()


# ring, particle texture (columns, rows), flow grid, ball radius:
#   square and ragged grids; the TALL textures of the weak-scaling bench (columns sharded, many rows); a small ball so that
#   strips get crowded and the split map kicks in (the same map must come out on every rank)
#   prune "1": the opaque pruning across the ranks (TB_PRUNE=1; fewer fragments than the oracle rasterises)
#   prune "seg": the owners fold their crowded bins in segments (PARITY B4) -- forced onto bins above 96 fragments
@pytest.mark.parametrize("ring,shape,G,radius,prune", [("owners", (96, 96), 64, 0.3, None), ("owners", (96, 96), 50, 0.3, "1"),
                                                       ("dist", (96, 96), 64, 0.3, None), ("owners", (16, 2048), 128, 0.3, None),
                                                       ("owners", (64, 1024), 96, 0.02, "1"), ("owners", (8, 8192), 64, 0.3, None),
                                                       ("owners", (64, 512), 64, 0.05, "1"), ("owners", (64, 1024), 96, 0.02, "seg"),
                                                       ("owners", (96, 96), 64, 0.3, "seg")])
def test_sharded_draw_equals_oracle(oracle, tmp_path, ring, shape, G, radius, prune):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    world, steps = min(torch.cuda.device_count(), 8), 8
    while shape[0] % world:
        world //= 2
    mp.spawn(_worker, args=(world, _free_port(), list(shape), G, steps, str(tmp_path), ring, radius, prune), nprocs=world, join=True)
    O = oracle
    DT = 1000 / 60
    P = O.make_params()
    PW, PH = shape
    cur, prev = O.spawn_ball(PW, PH, radius, 0.005), O.spawn_init(PW, PH)
    targets, flow = np.zeros((PW, PH, 4), np.float32), np.zeros((G, G, 4), np.float32)
    t = DT
    n = 0
    for _ in range(steps):
        t += DT
        new = O.integrate(P, cur, targets, flow, np.float32(t), np.float32(DT))
        prev, cur = cur, new
        n = O.splat(P, cur, prev, flow, np.float32(t))
    for r in range(world):
        assert np.array_equal(np.load(tmp_path / f"flow_{r}.npy"), flow), f"rank {r} flow"
    got = np.concatenate([np.load(tmp_path / f"state_{r}.npy") for r in range(world)], 0)
    assert np.array_equal(got, cur)
    emitted = sum(int(np.load(tmp_path / f"frags_{r}.npy")[0]) for r in range(world))
    assert 0 < emitted <= n if prune == "1" else emitted == n
    if prune == "seg" and radius < 0.1:
        assert sum(int(np.load(tmp_path / f"frags_{r}.npy")[1]) for r in range(world)) > 0, "no bin was folded in segments"
