"""Diagnostic: distribution of per-texel fragment-list lengths during the cfg3 bench workload."""
import ctypes as C
import sys
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np
import bench
from tendrils_b200 import _native as N

wl = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else 'cfg3']
t, first, sp = bench.build_sim(wl, 0, 1, 0, None)
L = N.load(); ctx = t.particles._ctx
G = wl['G'] * wl['G']
seg = np.zeros(2 * G, np.uint32)
for k in range(130):
    if k == 0: first.spawn(t)
    elif wl['every'] and k % wl['every'] == 0: sp.spawn(t)
    t.timer.tick(); t.step().draw()
    if k in (1, 5, 10, 20, 30, 40, 50, 59, 60, 61, 70, 90, 119, 121):
        N.check(ctx, L.tb_debug_segments(ctx, seg.ctypes.data_as(C.POINTER(C.c_uint32)), seg.size))
        n = (seg[1::2].astype(np.int64) - seg[0::2].astype(np.int64))
        nz = n[n > 0]
        w = n.reshape(-1, 32).max(1)       # per-warp max = serial length of the warp
        st = t.particles.buffers[0].download(); sp_ = np.hypot(st[..., 2], st[..., 3])
        print(f"k={k:3d} frags={t.particles.stats()['last_fragments']:9d} kept={n.sum():9d} texels>0 {len(nz):7d} mean {nz.mean():6.1f} "
              f"p99 {np.percentile(nz,99):6.0f} max {n.max():7d} sum(warp max)*32/kept {w.sum()*32/max(n.sum(),1):5.2f} "
              f"top5 {np.sort(n)[-5:]} speed mean {sp_.mean():.4f} frac>=limit {(sp_>=0.01).mean():.3f}", flush=True)
