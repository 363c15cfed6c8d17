"""Distribution of the per-strip fragment bins of a bench workload after `steps` steps (runs on a GPU box).
    python tools/binstats.py [workload] [steps]"""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import bench  # noqa: E402
from tendrils_b200 import _native as N  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "cfg3"
if name.startswith("cfg4x"):                 # the global picture of a weak-scaling run on N GPUs, on one GPU: cfg3 with N x the rows
    wl = dict(bench.WORKLOADS["cfg3"], rows=4096 * int(name[5:]))
else:
    wl = bench.WORKLOADS[name]
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
t, first, sp = bench.build_sim(wl, 0, 1, 0, None)
L, ctx = N.load(), t.particles._ctx
for k in range(steps):
    if k == 0:
        first.spawn(t)
    elif wl["every"] and k % wl["every"] == 0:
        sp.spawn(t)
    t.timer.tick()
    t.step().draw()
    if k in (0, 1, 2, 5, 10, 19, 30, 59, 60, 61, 70, 100, steps - 1):
        mb = L.tb_debug_max_bins()
        off, info = np.zeros(mb + 1, np.uint32), np.zeros(mb, np.uint32)
        nb, sw, sh = C.c_int32(), C.c_int32(), C.c_int32()
        up = C.POINTER(C.c_uint32)
        N.check(ctx, L.tb_debug_bins(ctx, off.ctypes.data_as(up), info.ctypes.data_as(up), C.byref(nb), C.byref(sw), C.byref(sh)))
        n = np.diff(off[:nb.value + 1].astype(np.int64))
        split = np.bincount(info[:nb.value] >> 24, minlength=8)
        q = np.percentile(n, [50, 90, 99, 99.9, 100])
        single = n[(info[:nb.value] >> 24) == 7]
        if len(single):
            print(f"          single-texel bins: {len(single)}, fragments max {single.max()} p99 {np.percentile(single, 99):.0f} mean {single.mean():.0f}")
        print(f"step {k:4d}: bins {nb.value} (strip {sw.value}x{sh.value}; bins by log2 split {split.tolist()}) frags {n.sum():10d} mean {n.mean():8.1f} "
              f"p50 {q[0]:7.0f} p90 {q[1]:7.0f} p99 {q[2]:7.0f} p99.9 {q[3]:7.0f} max {q[4]:8.0f}  top8 {np.sort(n)[-8:][::-1].tolist()}", flush=True)
