"""Who crowds the hottest texels?  CPU-only study (the oracle) of the weak-scaling bench workload at reduced scale, keeping
the particles-per-texel ratio of cfg4 on N GPUs (16 N per texel): the blend alpha of a fragment is min(|vel| / speedLimit, 1)
(PARITY B1), and a texel chain started at +-FLT_MAX settles (PARITY B4) only after sum(-log(1 - alpha)) > ~110, from a
bracket as tight as the colours' range after ~25.  Prints, for the most crowded texels of the last draw, how many fragments
that takes against how many the texel gets.
    python tools/crowd_alpha.py [N=8] [R=512] [G=128] [steps=25]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from oracle import oracle as O  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 8
R = int(sys.argv[2]) if len(sys.argv) > 2 else 512
G = int(sys.argv[3]) if len(sys.argv) > 3 else 128
steps = int(sys.argv[4]) if len(sys.argv) > 4 else 25
PH = (16 * N * G * G) // R                                   # 16 N particles per texel
O.build()
P = O.make_params()
S = O.make_spawn_pixels(jitter=(np.float32(np.float32(1.0 / G) * 2),) * 2, spawnMatrix=(-1, 0, 0, 0, 1, 0, 0, 0, 1))
img = bench.synthetic_image(G)
cur = prev = O.spawn_init(R, PH)
targets, flow = np.zeros((R, PH, 4), np.float32), np.zeros((G, G, 4), np.float32)
dt, t = 1000 / 60, 0.0
for k in range(steps):
    if k == 0:
        t += dt
        S.speed = 0.3
        prev, cur = cur, O.spawn_pixels_direct(S, R, PH, img, t)
    t += dt
    prev, cur = cur, O.integrate(P, cur, targets, flow, t, dt)
    n = O.splat(P, cur, prev, flow, t, mt=True)
print(f"{R}x{PH} particles ({16 * N} per texel) on a {G}^2 grid, {steps} steps; fragments of the last draw: {n}")
# the line of particle p: prev -> cur; its fragments ~ the texels between (counted at the end point: crowds barely move)
pos, vel = cur[..., :2].reshape(-1, 2), cur[..., 2:].reshape(-1, 2)
alive = pos[:, 0] > -1e5
alpha = np.minimum(np.hypot(vel[:, 0], vel[:, 1]) / P.speedLimit, 1.0)
gx = np.floor((pos[:, 0] * P.viewSize[0] * 0.5 + 0.5) * G).astype(np.int64)
gy = np.floor((pos[:, 1] * P.viewSize[1] * 0.5 + 0.5) * G).astype(np.int64)
ok = alive & (gx >= 0) & (gx < G) & (gy >= 0) & (gy < G)
tex = (gy * G + gx)[ok]
a = alpha[ok].astype(np.float64)
count = np.bincount(tex, minlength=G * G)
print(f"particles per texel: mean {count.mean():.0f}, p99 {np.percentile(count, 99):.0f}, max {count.max()}")
print(f"alpha over all particles: median {np.median(a):.3f}, 10 % below {np.percentile(a, 10):.4f}, share at 1.0: {np.mean(a >= 1.0):.3f}")
print("hottest texels: particles, median alpha, fragments to settle from +-FLT_MAX (110 e-folds) / from a tight bracket (25)")
for tx in np.argsort(count)[::-1][:8]:
    at = a[tex == tx]
    e = -np.log1p(-np.minimum(at, 1 - 1e-7))                 # e-folds per fragment (draw order ~ particle order)
    c = np.cumsum(e)
    need = lambda folds: int(np.searchsorted(c, folds)) + 1 if c[-1] > folds else None
    print(f"  texel {tx:6d}: {len(at):6d} particles, median alpha {np.median(at):.4f}, settle after {need(110)} / {need(25)}")
