"""Property tests of the splat's per-line arithmetic on the CPU.  The count pass (k_splat_hist) and the emit pass
(k_splat_scatter) must agree on every line's fragments, or every later fragment of a bin lands in the wrong slot:
  * `count_fragments` / `prim_setup` (closed form) must equal the number of fragments `raster_line` emits;
  * `prim_fragment(j)`, j < n, must enumerate exactly raster_line's fragments (texel and interpolation parameter, bit for bit);
  * `prim_fragment_texel` (the count pass's division-free estimate with its exact fallback) must name the same texels.
The device functions are pure arithmetic, so the test cuts their text out of tendrils_b200/csrc/tb_kernels.cuh and
tb_splat.cuh (between the [raster-begin]/[raster-end] and [prim-begin]/[prim-end] markers), compiles it with g++
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
#include <algorithm>
#include <tuple>
#include <vector>
namespace tb {
static constexpr float kInert = -1000000.0f;
%(body)s
struct PrimGeom {
    float ma, mb, na, nb;
    int c0;
    uint32_t flags;
};
%(prim)s
%(texel)s
}
// prim_setup / prim_fragment / prim_fragment_texel against raster_line: returns the number of disagreeing segments
extern "C" long long rh_check_prims(long long n, const float *seg, float vsx, float vsy, int W, int H, long long *first_bad) {
    long long bad = 0;
    *first_bad = -1;
    for (long long i = 0; i < n; ++i) {
        const float4 sa = make_float4(seg[4 * i], seg[4 * i + 1], 0.001f, 0.002f), sb = make_float4(seg[4 * i + 2], seg[4 * i + 3], 0.003f, 0.f);
        tb::PrimGeom P;
        const unsigned cnt = tb::prim_setup(sa, sb, vsx, vsy, W, H, P);
        std::vector<std::tuple<int, int, unsigned>> want, got;
        if (tb::splat_vertex_ok(sa) && tb::splat_vertex_ok(sb)) {
            const float hw = 0.5f * (float)W, hh = 0.5f * (float)H;
            const float xa = (sa.x * vsx) * hw + hw, ya = (sa.y * vsy) * hh + hh, xb = (sb.x * vsx) * hw + hw, yb = (sb.y * vsy) * hh + hh;
            tb::raster_line(xa, ya, xb, yb, W, H, [&](int gx, int gy, float t) { want.emplace_back(gx, gy, __float_as_uint(t)); });
        }
        bool ok = cnt == want.size() && cnt == tb::count_fragments(sa, sb, vsx, vsy, W, H);
        for (unsigned j = 0; ok && j < cnt; ++j) {
            int gx, gy; float t;
            tb::prim_fragment(P, j, W, H, gx, gy, t);
            got.emplace_back(gx, gy, __float_as_uint(t));
            if (P.flags & 2u) {
                const float inv_dm = 1.0f / (P.mb - P.ma);
                const float eps = ((fabsf(P.na) + fabsf(P.nb - P.na)) + 1.0f) * 1.9073486328125e-06f;
                int ex, ey;
                tb::prim_fragment_texel(P, j, inv_dm, eps, ex, ey);
                ok = ok && ex == gx && ey == gy;
            }
        }
        std::sort(want.begin(), want.end());
        std::sort(got.begin(), got.end());
        ok = ok && want == got;
        if (!ok) { if (*first_bad < 0) *first_bad = i; ++bad; }
    }
    return bad;
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
    ssrc = open(os.path.join(ROOT, "tendrils_b200", "csrc", "tb_splat.cuh")).read()
    prim = ssrc[ssrc.index("// [prim-begin]"):ssrc.index("// [prim-end]")].replace("__device__", "").replace("__noinline__", "")
    texel = ssrc[ssrc.index("__device__ __forceinline__ void prim_fragment_texel("):ssrc.index("constexpr int kHistWarps")]
    cpp.write_text(HARNESS % {"math": str(math), "body": body, "prim": prim, "texel": texel.replace("__device__", "")})
    out = d / "libraster_host.so"
    subprocess.run(["g++", "-O2", "-std=c++17", "-march=x86-64-v3", "-ffp-contract=off", "-fno-fast-math", "-fPIC", "-shared",
                    "-Wno-unknown-pragmas", "-Wno-unused-function", "-I/usr/local/cuda/include",
                    "-I", os.path.join(ROOT, "tests", "host_harness"), "-o", str(out), str(cpp)], check=True)
    L = C.CDLL(str(out))
    L.rh_check.restype = C.c_longlong
    L.rh_check.argtypes = [C.c_longlong, C.POINTER(C.c_float), C.c_float, C.c_float, C.c_int, C.c_int,
                           C.POINTER(C.c_longlong), C.POINTER(C.c_longlong)]
    L.rh_check_prims.restype = C.c_longlong
    L.rh_check_prims.argtypes = [C.c_longlong, C.POINTER(C.c_float), C.c_float, C.c_float, C.c_int, C.c_int, C.POINTER(C.c_longlong)]
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


@pytest.mark.parametrize("W,H,vs,seed", [(64, 64, (1.0, 1.0), 11), (1024, 1024, (1.0, 1.0), 12), (40, 24, (1.0, 40 / 24), 13),
                                         (7, 129, (129 / 7, 1.0), 14), (1, 1, (1.0, 1.0), 15), (2048, 2048, (1.0, 1.0), 16)])
def test_emit_enumerates_what_the_count_counted(rh, W, H, vs, seed):
    """prim_setup + prim_fragment (k_splat_scatter) and prim_fragment_texel (k_splat_hist) against raster_line."""
    rng = np.random.default_rng(seed)
    n = 150_000
    seg = segments(rng, n, W, H)
    seg[:8] = [[np.nan, 0, 0, 0], [0, 0, np.inf, 0], [-1e6, -1e6, 0, 0], [0, 0, 0, 0], [0.5, 0.5, 0.5, 0.5],
               [-1, -1, 1, 1], [1, 1, -1, -1], [-1, 1, 1, -1]]
    big = rng.choice([1e3, 1e20, 1e35, 1e37, 3e38], (64, 4)) * rng.choice([-1, 1], (64, 4))
    seg[8:72] = np.where(rng.random((64, 4)) < 0.5, big, rng.uniform(-1.2, 1.2, (64, 4))).astype(np.float32)
    first_bad = C.c_longlong()
    bad = rh.rh_check_prims(n, seg.ctypes.data_as(C.POINTER(C.c_float)), vs[0], vs[1], W, H, C.byref(first_bad))
    assert bad == 0, (bad, first_bad.value, seg[first_bad.value] if first_bad.value >= 0 else None)
