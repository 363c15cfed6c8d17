"""Property test of the splat's fused COUNT pass on the CPU: `count_fragments` (closed form, rides in k_integrate)
must equal the number of fragments `raster_line` emits (what k_splat_emit writes) for every segment -- a mismatch
would shift every later fragment's slot.  The two device functions are pure arithmetic, so the test cuts their text
out of tendrils_b200/csrc/tb_kernels.cuh (between the [raster-begin]/[raster-end] markers), compiles it with g++
-ffp-contract=off against a shim of the single-operation intrinsics, and throws adversarial segments at it.
Nothing here is used by the product."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

HARNESS = r'''
#include <cuda_runtime.h>
#include "cuda_intrinsics_shim.h"
#include "%(math)s"
namespace tb {
static constexpr float kInert = -1000000.0f;
%(body)s
}
extern "C" long long rh_check(long long n, const float *seg /* n x 4: NDC xa ya xb yb */, float vsx, float vsy, int W, int H,
                              long long *total, long long *first_bad) {
    long long bad = 0;
    *total = 0; *first_bad = -1;
    for (long long i = 0; i < n; ++i) {
        const float4 sa = make_float4(seg[4 * i], seg[4 * i + 1], 0.001f, 0.002f), sb = make_float4(seg[4 * i + 2], seg[4 * i + 3], 0.003f, 0.f);
        const unsigned fast = tb::count_fragments(sa, sb, vsx, vsy, W, H);
        unsigned slow = 0;
        if (tb::splat_vertex_ok(sa) && tb::splat_vertex_ok(sb)) {
            const float hw = 0.5f * (float)W, hh = 0.5f * (float)H;
            const float xa = (sa.x * vsx) * hw + hw, ya = (sa.y * vsy) * hh + hh, xb = (sb.x * vsx) * hw + hw, yb = (sb.y * vsy) * hh + hh;
            tb::raster_line(xa, ya, xb, yb, W, H, [&](int gx, int gy, float) { if (gx >= 0 && gx < W && gy >= 0 && gy < H) ++slow; else slow += 1000000; });
        }
        *total += slow;
        if (fast != slow) { if (*first_bad < 0) *first_bad = i; ++bad; }
    }
    return bad;
}
'''


@pytest.fixture(scope="module")
def rh(tmp_path_factory):
    src = open(os.path.join(ROOT, "tendrils_b200", "csrc", "tb_kernels.cuh")).read()
    body = src[src.index("// [raster-begin]"):src.index("// [raster-end]")].replace("__device__", "")
    d = tmp_path_factory.mktemp("rh")
    math = d / "tb_math_host.cuh"
    math.write_text(open(os.path.join(ROOT, "tendrils_b200", "csrc", "tb_math.cuh")).read().replace("__device__", ""))
    cpp = d / "raster_host.cpp"
    cpp.write_text(HARNESS % {"math": str(math), "body": body})
    out = d / "libraster_host.so"
    subprocess.run(["g++", "-O2", "-std=c++17", "-march=x86-64-v3", "-ffp-contract=off", "-fno-fast-math", "-fPIC", "-shared",
                    "-Wno-unknown-pragmas", "-Wno-unused-function", "-I/usr/local/cuda/include",
                    "-I", os.path.join(ROOT, "tests", "host_harness"), "-o", str(out), str(cpp)], check=True)
    L = C.CDLL(str(out))
    L.rh_check.restype = C.c_longlong
    L.rh_check.argtypes = [C.c_longlong, C.POINTER(C.c_float), C.c_float, C.c_float, C.c_int, C.c_int,
                           C.POINTER(C.c_longlong), C.POINTER(C.c_longlong)]
    return L


def segments(rng, n, W, H):
    """NDC segments built from WINDOW coordinates that stress the rule: centres (k + 0.5), pixel edges, the borders,
    just-off values (one ulp), tiny and long lines, lines leaving the grid on every side."""
    special = np.concatenate([np.arange(-2, W + 3, dtype=np.float64), np.arange(-2, W + 3) + 0.5,
                              [0.25, 0.75, 0.999999, 1.0, 1.000001, W - 1.000001, W - 1.0, W - 0.999999, W - 0.5, W - 0.25]])

    def coord(size, count):
        kind = rng.integers(0, 4, count)
        c = np.where(kind == 0, rng.choice(special, count) * size / W,
                     np.where(kind == 1, rng.uniform(-3, size + 3, count), rng.uniform(0, size, count)))
        c = c.astype(np.float32)
        jig = rng.integers(-2, 3, count)                                   # a few ulps either way
        return np.where(kind == 3, np.nextafter(c, np.where(jig > 0, np.inf, -np.inf).astype(np.float32)), c).astype(np.float32)

    xa, ya = coord(W, n), coord(H, n)
    mode = rng.integers(0, 4, n)
    length = np.where(mode == 0, rng.uniform(0, 2, n), np.where(mode == 1, rng.uniform(0, 12, n), rng.uniform(0, 1.5 * max(W, H), n)))
    ang = np.where(rng.integers(0, 3, n) == 0, rng.integers(0, 8, n) * (np.pi / 4), rng.uniform(0, 2 * np.pi, n))
    xb = np.where(mode == 3, coord(W, n), xa + length * np.cos(ang)).astype(np.float32)
    yb = np.where(mode == 3, coord(H, n), ya + length * np.sin(ang)).astype(np.float32)
    to_ndc = lambda w, size: ((w.astype(np.float64) - size / 2) / (size / 2)).astype(np.float32)
    return np.ascontiguousarray(np.stack([to_ndc(xa, W), to_ndc(ya, H), to_ndc(xb, W), to_ndc(yb, H)], 1))


@pytest.mark.parametrize("W,H,vs,seed", [(64, 64, (1.0, 1.0), 1), (1024, 1024, (1.0, 1.0), 2), (40, 24, (1.0, 40 / 24), 3),
                                         (7, 129, (129 / 7, 1.0), 4), (1, 1, (1.0, 1.0), 5), (2048, 2048, (1.0, 1.0), 6)])
def test_closed_form_count_equals_enumeration(rh, W, H, vs, seed):
    rng = np.random.default_rng(seed)
    n = 400_000
    seg = segments(rng, n, W, H)
    seg[:8] = [[np.nan, 0, 0, 0], [0, 0, np.inf, 0], [-1e6, -1e6, 0, 0], [0, 0, 0, 0], [0.5, 0.5, 0.5, 0.5],
               [-1, -1, 1, 1], [1, 1, -1, -1], [-1, 1, 1, -1]]
    total, first_bad = C.c_longlong(), C.c_longlong()
    bad = rh.rh_check(n, seg.ctypes.data_as(C.POINTER(C.c_float)), vs[0], vs[1], W, H, C.byref(total), C.byref(first_bad))
    assert bad == 0, (bad, first_bad.value, seg[first_bad.value] if first_bad.value >= 0 else None)
    assert total.value > n // 4 or W * H == 1                               # the segments do produce fragments


@pytest.mark.parametrize("W,H,vs", [(64, 64, (1.0, 1.0)), (1024, 1024, (1.0, 1.0)), (56, 63, (1.0, 56 / 63))])
def test_count_with_astronomic_coordinates(rh, W, H, vs):
    """Finite positions so large that the WINDOW coordinate overflows to +-Inf (|pos| beyond ~1e35): every t is NaN, the
    enumeration emits nothing and the closed form must say 0 too (it once counted the columns: stale slots)."""
    rng = np.random.default_rng(W)
    n = 300_000
    big = rng.choice([1e3, 1e7, 1e20, 1e30, 1e35, 1e36, 1e37, 3e38], (n, 4)) * rng.choice([-1, 1], (n, 4))
    seg = np.ascontiguousarray(np.where(rng.random((n, 4)) < 0.5, big, rng.uniform(-1.2, 1.2, (n, 4))).astype(np.float32))
    total, first_bad = C.c_longlong(), C.c_longlong()
    bad = rh.rh_check(n, seg.ctypes.data_as(C.POINTER(C.c_float)), vs[0], vs[1], W, H, C.byref(total), C.byref(first_bad))
    assert bad == 0, (bad, first_bad.value, seg[first_bad.value] if first_bad.value >= 0 else None)
