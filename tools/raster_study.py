"""How far is RASTER-1 (spec/PARITY.md: the centre-sampled major-axis line rule both the oracle and the CUDA path
implement) from the EXACT diamond-exit rule of OpenGL ES 2.0 section 3.4.1?

The GL specification allows implementations to deviate from diamond-exit within limits (at most one pixel off in the
minor direction, no gaps, no doubled columns), so the rule is a decision, not a fact about the reference; this
script measures the decision against the ideal, in exact rational arithmetic, on random segments shaped like the
workload's (0 .. 12 pixels long).  Run: python tools/raster_study.py [n_segments]
"""
import os
import sys
from fractions import Fraction as F

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

f32 = np.float32
HALF = F(1, 2)


def diamond_exit(xa, ya, xb, yb, W, H):
    """Fragments of the segment a->b under the exact rule: pixel (i,j) is produced iff the segment intersects the open
    diamond |x-xc|+|y-yc| < 1/2 around its centre and b is not inside that diamond (the segment exits it)."""
    xa, ya, xb, yb = (F(float(v)) for v in (xa, ya, xb, yb))
    dx, dy = xb - xa, yb - ya
    out = set()
    if dx == 0 and dy == 0:
        return out
    i0, i1 = int(min(xa, xb)) - 1, int(max(xa, xb)) + 1
    j0, j1 = int(min(ya, yb)) - 1, int(max(ya, yb)) + 1
    for j in range(max(j0, 0), min(j1, H - 1) + 1):
        for i in range(max(i0, 0), min(i1, W - 1) + 1):
            xc, yc = F(i) + HALF, F(j) + HALF
            lo, hi = F(-10**9), F(10**9)                  # open interval of t with P(t) inside the diamond
            empty = False
            for sx in (1, -1):
                for sy in (1, -1):
                    # sx*(x-xc) + sy*(y-yc) < 1/2  with x = xa + t dx ...
                    a = sx * dx + sy * dy
                    b = sx * (xa - xc) + sy * (ya - yc)
                    if a == 0:
                        if not b < HALF:
                            empty = True
                    elif a > 0:
                        hi = min(hi, (HALF - b) / a)
                    else:
                        lo = max(lo, (HALF - b) / a)
            if empty or not lo < hi:
                continue
            if lo < 1 and hi > 0 and hi <= 1:             # touches the segment, and leaves before (or at) b
                out.add((i, j))
    return out


def raster1(oracle, xa, ya, xb, yb, W, H):
    """The oracle's fragments for the same segment, driven through one particle's prev -> cur line."""
    O = oracle
    R = 2
    P = O.make_params(speedLimit=1.0)
    cur, prev = O.spawn_init(R, R), O.spawn_init(R, R)
    prev[0, 0] = (xa, ya, 0.5, 0.0)
    cur[0, 0] = (xb, yb, 0.5, 0.0)
    flow = np.zeros((H, W, 4), f32)
    O.splat(P, cur, prev, flow, f32(1.0))
    return {(int(i), int(j)) for j, i in np.argwhere(flow[..., 3] != 0)}


def window(ndc, size):
    return f32(f32(f32(ndc) * f32(1.0)) * f32(size / 2)) + f32(size / 2)          # V3, viewSize = 1


def study(n=4000, W=64, H=64, seed=5):
    from oracle import oracle as O
    rng = np.random.default_rng(seed)
    same = extra = missing = 0
    worst = 0
    end_only = 0
    for _ in range(n):
        a = rng.uniform(-0.8, 0.8, 2).astype(f32)
        ang, ln = rng.uniform(0, 2 * np.pi), rng.uniform(0, 12) / (W / 2)
        b = (a + ln * np.array([np.cos(ang), np.sin(ang)])).astype(f32)
        xa, ya, xb, yb = window(a[0], W), window(a[1], H), window(b[0], W), window(b[1], H)
        ideal = diamond_exit(xa, ya, xb, yb, W, H)
        got = raster1(O, a[0], a[1], b[0], b[1], W, H)
        if ideal == got:
            same += 1
            continue
        diff = ideal ^ got
        extra += len(got - ideal)
        missing += len(ideal - got)
        worst = max(worst, len(diff))
        # is every differing pixel within one pixel of an end point?
        near = all(min(abs(i + .5 - float(x)) + abs(j + .5 - float(y)) for x, y in ((xa, ya), (xb, yb))) <= 1.5 for i, j in diff)
        end_only += 1 if near else 0
    return {"segments": n, "identical": same, "differing": n - same, "differ_only_near_endpoints": end_only,
            "fragments_extra": extra, "fragments_missing": missing, "worst_symmetric_difference": worst}


if __name__ == "__main__":
    print(study(int(sys.argv[1]) if len(sys.argv) > 1 else 4000))
