"""Mint the golden vectors under tests/golden/ from the CPU oracle (the reference ships none:
SURVEY.md section 4 / 8c, "parity unpinned").  They pin the oracle against accidental change and
give the GPU tests fixed inputs/outputs that do not need the oracle at run time.

    python tools/make_golden.py        # rewrites tests/golden/oracle_v1.npz
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import oracle as O  # noqa: E402
from util import synthetic_image  # noqa: E402

DT = 1000 / 60


def run(R, G, steps, **over):
    P = O.make_params(**over)
    cur, prev = O.spawn_ball(R, R, 0.3, 0.005), O.spawn_init(R, R)
    targets, flow = np.zeros((R, R, 4), np.float32), np.zeros((G, G, 4), np.float32)
    t = DT                       # the spawn pass ticked the timer once
    frags = []
    for _ in range(steps):
        t += DT
        new = O.integrate(P, cur, targets, flow, np.float32(t), np.float32(DT))
        prev, cur = cur, new
        frags.append(O.splat(P, cur, prev, flow, np.float32(t)))
    return cur, prev, flow, np.array(frags, np.int64)


def main():
    out = {}
    for R, G, steps in [(8, 16, 1), (8, 16, 10), (16, 16, 100), (64, 64, 10)]:
        cur, prev, flow, frags = run(R, G, steps)
        tag = f"sim_R{R}_G{G}_n{steps}"
        out[tag + "_cur"], out[tag + "_prev"], out[tag + "_flow"], out[tag + "_frags"] = cur, prev, flow, frags
    out["ball_R16_r1_s0"] = O.spawn_ball(16, 16, 1.0, 0.0)
    out["ball_R16_r03_s005"] = O.spawn_ball(16, 16, 0.3, 0.005)
    img = synthetic_image(24, 20)
    out["image_24x20"] = img
    S = O.make_spawn_pixels(spawnSize=(1, 1), jitter=(np.float32(2 / 32), np.float32(2 / 32)), speed=0.3, bias=1.0,
                            spawnMatrix=(-1, 0, 0, 0, 1, 0, 0, 0, 1))
    out["direct_R16"] = O.spawn_pixels_direct(S, 16, 16, img, np.float32(3 * DT))
    state = O.spawn_ball(16, 16, 0.5, 0.004)
    for v in O.SAMPLE_VARIANTS:
        out[f"sample_{v}_R16"] = O.spawn_pixels_sample(S, v, state, img, np.float32(5 * DT))
    xs = np.linspace(-7, 7, 57, dtype=np.float32)
    out["snoise_in"] = np.stack([xs, xs[::-1] * np.float32(0.37), xs * np.float32(1.91) + np.float32(0.123)], -1)
    out["snoise_out"] = np.array([O.snoise3(*v) for v in out["snoise_in"]], np.float32)
    ang = np.linspace(-20, 20, 401, dtype=np.float32)
    out["sin_in"] = ang
    out["sin_out"] = np.array([O.sin(a) for a in ang], np.float32)
    out["cos_out"] = np.array([O.cos(a) for a in ang], np.float32)
    co = np.stack([np.linspace(0.5, 300.5, 61, dtype=np.float32), np.linspace(7.5, 90.25, 61, dtype=np.float32)], -1)
    out["random_in"] = co
    out["random_out"] = np.array([O.random(*c) for c in co], np.float32)
    for PH in (8, 512, 2048, 4096):
        row, cur = O.vertex_table(PH)
        out[f"vtx_row_{PH}"], out[f"vtx_cur_{PH}"] = row.astype(np.int32), cur.astype(np.int8)
    path = os.path.join(ROOT, "tests", "golden", "oracle_v1.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes;", len(out), "arrays")


if __name__ == "__main__":
    main()
