"""How many primitives of a bench workload are opaque at both ends (alpha == 1 at both vertices), step by step (runs on a GPU box).
    python tools/opaque_stats.py [workload] [steps]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import bench  # noqa: E402

wl = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "cfg3"]
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 70
t, first, sp = bench.build_sim(wl, 0, 1, 0, None)
sl = np.float32(t.state["speedLimit"])
for k in range(steps):
    if k == 0:
        first.spawn(t)
    elif wl["every"] and k % wl["every"] == 0:
        sp.spawn(t)
    t.timer.tick()
    t.step().draw()
    if k in (0, 2, 5, 10, 19, 30, 45, 59, 60, 61, 65, steps - 1):
        cur, prev = t.particles.buffers[0].download(), t.particles.buffers[1].download()
        rows = cur.shape[1] // 2 + 1                                  # the rows that draw (D6)
        a = lambda s: np.minimum(np.sqrt(s[:, :rows, 2] ** 2 + s[:, :rows, 3] ** 2) / sl, np.float32(1.0))
        ac, ap = a(cur), a(prev)
        both = (ac == 1) & (ap == 1)
        print(f"step {k:3d}: alpha==1 cur {np.mean(ac == 1):.3f} prev {np.mean(ap == 1):.3f} both {np.mean(both):.3f}  mean alpha {ac.mean():.3f}  "
              f"fragments {t.particles.stats()['last_fragments']}", flush=True)
