"""The blend of one batch of 32 fragments (fold_prep / fold_apply of tendrils_b200/csrc/tb_splat.cuh: the rounds over the
in-batch rank with predecessor shuffles, the chained path for batches that pile up on few texels, its start at the last
fragment with alpha == 1) cut out of the product source unchanged and run on a CPU emulation of a warp -- one thread per
lane, shuffles / votes / match as barriers (tests/host_harness/warp_emu.h) -- against a plain sequential blend in
fragment order, on bins that are uniform, crowded, opaque, non-finite and signed-zero.  Nothing here is used by the product."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_fp = C.POINTER(C.c_float)
_up = C.POINTER(C.c_uint32)

HARNESS = r'''
#include <algorithm>
#include <cstddef>
#include <cuda_runtime.h>
#include "cuda_intrinsics_shim.h"
#include "warp_emu.h"
#include "%(math)s"
namespace tb {
using std::min; using std::max;
constexpr int kFoldTexels = 128;
constexpr uint32_t kKeyLocalMask = 0x000fffffu;
struct __attribute__((aligned(16))) Frag { float cx, cy, a; uint32_t key; };
constexpr int kFoldStage = 64;
constexpr uint32_t kFoldRounds = %(rounds)s;
struct FoldWarp { Frag stage[2][kFoldStage]; float4 term[32]; float om[32]; unsigned long long bar[2]; };
%(fold)s
}
using namespace tb;
static FoldWarp g_w;
static float4 g_tex[kFoldTexels];

// blend n fragments (x, y, a, key), in order, 32 at a time, onto tex[0..R) with texel indices key - lo
extern "C" void fh_fold(const float *frags, int n, int lo, int R, float time, float *tex) {
    for (int i = 0; i < R; ++i) g_tex[i] = make_float4(tex[4 * i], tex[4 * i + 1], tex[4 * i + 2], tex[4 * i + 3]);
    tb_run_warp(0, 0, [&] {
        const int lane = (int)threadIdx.x;
        for (int b0 = 0; b0 < n; b0 += 32) {
            const int i = b0 + lane;
            Frag f{0.f, 0.f, 0.f, 0u};
            if (i < n) { f.cx = frags[4 * i]; f.cy = frags[4 * i + 1]; f.a = frags[4 * i + 2]; f.key = __float_as_uint(frags[4 * i + 3]); }
            const FoldPrep p = fold_prep(f, i < n, (uint32_t)lo, time, lane);
            fold_apply(g_w, g_tex, p, i < n, lane);
        }
    });
    for (int i = 0; i < R; ++i) { tex[4 * i] = g_tex[i].x; tex[4 * i + 1] = g_tex[i].y; tex[4 * i + 2] = g_tex[i].z; tex[4 * i + 3] = g_tex[i].w; }
}

// The same bin folded in P segments (PARITY B4) the way k_splat_fold / k_splat_mend do it: segment 0 from the texels' values,
// segments 1.. on the two bracketing chains with their records, then the join.  stats: [0] texel-segments settled,
// [1] records written, [2] 1 if the join fell back to the plain fold.
static Frag g_rec[1 << 16];
static float4 g_out[16][kFoldTexels], g_hi[kFoldTexels];
static uint32_t g_cnt[16], g_mask[kFoldTexels / 32 + 1];
extern "C" void fh_fold_segments(const float *frags, int n, int lo, int R, float time, float *tex, int P, int *stats) {
    auto frag_at = [&](int i) { Frag f{frags[4 * i], frags[4 * i + 1], frags[4 * i + 2], __float_as_uint(frags[4 * i + 3])}; return f; };
    const int span = ((n + P - 1) / P + 63) / 64 * 64;
    stats[0] = stats[1] = stats[2] = 0;
    for (int part = 0; part < P; ++part) {
        const int b0 = std::min(part * span, n), b1 = std::min(b0 + span, n);
        if (part == 0) for (int i = 0; i < R; ++i) g_tex[i] = make_float4(tex[4 * i], tex[4 * i + 1], tex[4 * i + 2], tex[4 * i + 3]);
        tb_run_warp(0, 0, [&] {
            const int lane = (int)threadIdx.x;
            uint32_t n_rec = 0;
            if (part) seg_begin(g_tex, g_hi, g_mask, (uint32_t)R, lane);
            for (int s0 = b0; s0 < b1; s0 += 32) {
                const int i = s0 + lane;
                Frag f{0.f, 0.f, 0.f, 0u};
                if (i < b1) f = frag_at(i);
                if (part) seg_batch(g_w, g_tex, g_hi, g_mask, f, i < b1, (uint32_t)lo, time, lane, g_rec + b0, n_rec);
                else fold_apply(g_w, g_tex, fold_prep(f, i < b1, (uint32_t)lo, time, lane), i < b1, lane);
            }
            if (part) { seg_end(g_tex, g_hi, g_mask, (uint32_t)R, g_out[part], lane); if (lane == 0) g_cnt[part] = n_rec; }
            else for (int l = lane; l < R; l += 32) g_out[0][l] = g_tex[l];
        });
        if (part) { stats[1] += (int)g_cnt[part]; for (int l = 0; l < R; ++l) { const uint32_t ex = __float_as_uint(g_out[part][l].x); stats[0] += ex != kSegOpen && ex != kSegRedo; } }
    }
    for (int i = 0; i < R; ++i) g_tex[i] = g_out[0][i];
    bool redo = false;
    for (int part = 1; part < P && !redo; ++part) {
        const int b0 = std::min(part * span, n);
        tb_run_warp(0, 0, [&] {
            const int lane = (int)threadIdx.x;
            const bool r = mend_begin(g_tex, g_out[part], (uint32_t)R, g_mask, lane);
            if (r) { if (lane == 0) redo = true; return; }
            for (int s0 = 0; s0 < (int)g_cnt[part]; s0 += 32) {
                const int i = s0 + lane;
                Frag f{0.f, 0.f, 0.f, 0u};
                if (i < (int)g_cnt[part]) f = g_rec[b0 + i];
                mend_batch(g_w, g_tex, g_mask, f, i < (int)g_cnt[part], (uint32_t)lo, time, lane);
            }
            mend_end(g_tex, g_out[part], (uint32_t)R, g_mask, lane);
        });
    }
    stats[2] = redo;
    if (redo) { fh_fold(frags, n, lo, R, time, tex); return; }
    for (int i = 0; i < R; ++i) { tex[4 * i] = g_tex[i].x; tex[4 * i + 1] = g_tex[i].y; tex[4 * i + 2] = g_tex[i].z; tex[4 * i + 3] = g_tex[i].w; }
}
'''


@pytest.fixture(scope="module", params=["6", "1", "31"], ids=["rounds6", "rounds1", "rounds31"])
def fh(request, tmp_path_factory):
    """Built with the product's threshold between the two paths, with everything chained and with (almost) nothing chained."""
    return build_harness(tmp_path_factory.mktemp("fh"), request.param)


def build_harness(d, rounds):
    csrc = os.path.join(ROOT, "tendrils_b200", "csrc")
    math = d / "tb_math_host.cuh"
    math.write_text(open(os.path.join(csrc, "tb_math.cuh")).read().replace("__device__", ""))
    src = open(os.path.join(csrc, "tb_splat.cuh")).read()
    fold = src[src.index("// The order-independent half of one batch"):src.index("// [fold-host-end]")]
    assert "kFoldRounds" in fold and "asm" not in fold
    cpp = d / "fold_host.cpp"
    cpp.write_text(HARNESS % {"math": str(math), "fold": fold.replace("__device__", ""), "rounds": rounds})
    out = d / "libfold_host.so"
    subprocess.run(["g++", "-O2", "-std=c++20", "-march=x86-64-v3", "-ffp-contract=off", "-fno-fast-math", "-fPIC", "-shared", "-pthread",
                    "-Wno-unknown-pragmas", "-Wno-unused-function", "-Wno-attributes", "-I/usr/local/cuda/include",
                    "-I", os.path.join(ROOT, "tests", "host_harness"), "-o", str(out), str(cpp)], check=True)
    L = C.CDLL(str(out))
    L.fh_fold.restype = None
    L.fh_fold.argtypes = [_fp, C.c_int, C.c_int, C.c_int, C.c_float, _fp]
    L.fh_fold_segments.restype = None
    L.fh_fold_segments.argtypes = [_fp, C.c_int, C.c_int, C.c_int, C.c_float, _fp, C.c_int, C.POINTER(C.c_int)]
    return L


def plain_fold(frags, lo, time, tex):
    """spec/PARITY.md B2: dst = src*a + dst*(1-a), one rounding per operator, fragment order."""
    tex = tex.copy()
    f32 = np.float32
    with np.errstate(all="ignore"):
        for cx, cy, a, key in frags:
            t = int(np.float32(key).view(np.uint32)) - lo
            src = np.array([cx, cy, time, a], f32)
            tex[t] = (src * f32(a)).astype(f32) + (tex[t] * f32(f32(1.0) - f32(a))).astype(f32)
    return tex


def same_bits(a, b):
    """bit for bit, except that any NaN equals any NaN"""
    a, b = np.ascontiguousarray(a, np.float32), np.ascontiguousarray(b, np.float32)
    nan = np.isnan(a) & np.isnan(b)
    return np.array_equal(np.where(nan, 0, a.view(np.uint32)), np.where(nan, 0, b.view(np.uint32)))


def make_bin(R, lo, n, kind, seed, alpha_hi=1.0):
    rng = np.random.default_rng(seed)
    frags = np.zeros((n, 4), np.float32)
    frags[:, 0:2] = rng.normal(0, 0.01, (n, 2))
    frags[:, 2] = rng.uniform(0, alpha_hi, n)
    tex_id = rng.integers(0, R, n)
    if kind != "uniform" and R > 1:
        hot = rng.integers(0, R, 3)
        tex_id = np.where(rng.random(n) < 0.8, rng.choice(hot, n), tex_id)
    if kind in ("opaque", "wild"):
        frags[rng.random(n) < 0.4, 2] = 1.0                                         # overwrite the texel
        frags[rng.random(n) < 0.05, 2] = 0.0
        frags[rng.random(n) < 0.05, 0] = -0.0                                       # -0 colour: such a fragment must not start a chain
    if kind == "wild":
        bad = rng.random(n) < 0.03
        frags[bad, 0] = rng.choice([np.nan, np.inf, -np.inf, 3e38], bad.sum())
        frags[rng.random(n) < 0.02, 2] = rng.choice([np.nan, 1.5, -0.25], 1)[0]
    frags[:, 3] = (tex_id + lo).astype(np.uint32).view(np.float32)
    tex0 = rng.normal(0, 0.01, (R, 4)).astype(np.float32)
    tex0[:, 2] = rng.uniform(0, 1000, R)
    if kind == "wild":
        tex0[rng.integers(0, R, 2)] = np.float32(np.inf)
    return frags, tex0


@pytest.mark.parametrize("R,lo,n,kind,seed", [(128, 0, 700, "uniform", 1), (128, 0, 900, "hot", 2), (16, 48, 500, "hot", 3), (4, 124, 400, "hot", 4),
                                              (1, 77, 300, "hot", 5), (128, 0, 800, "opaque", 6), (16, 16, 600, "opaque", 7), (1, 0, 257, "opaque", 8),
                                              (128, 0, 640, "wild", 9), (4, 8, 333, "wild", 10), (128, 0, 31, "uniform", 11), (128, 0, 0, "uniform", 12)])
def test_batches_equal_the_plain_fold(fh, R, lo, n, kind, seed):
    time = np.float32(1234.5)
    frags, tex0 = make_bin(R, lo, n, kind, seed)
    want = plain_fold(frags, lo, time, tex0)
    got = tex0.copy()
    fh.fh_fold(frags.ctypes.data_as(_fp), n, lo, R, time, got.ctypes.data_as(_fp))
    assert same_bits(got, want)


@pytest.mark.parametrize("R,lo,n,P,kind,alpha_hi,seed,expect", [
    (1, 77, 6000, 4, "hot", 1.0, 21, "settles_most"),           # one crowded texel: every later segment settles it
    (4, 124, 9000, 8, "hot", 1.0, 22, "settles_most"),          # hot and lukewarm texels side by side
    (64, 0, 12000, 16, "hot", 1.0, 23, "settles"),
    (64, 64, 5000, 2, "uniform", 1.0, 24, "any"),          # ~40 fragments per texel and segment: most stay on record
    (4, 0, 8000, 4, "hot", 1e-4, 25, "records"),           # tiny alphas: nothing settles, everything is replayed
    (16, 16, 9000, 8, "opaque", 1.0, 26, "settles"),       # overwriting fragments settle a texel at once
    (16, 32, 7000, 4, "wild", 1.0, 27, "any"),             # NaN / Inf colours and texels: the serial fallback
    (2, 0, 100, 4, "hot", 1.0, 28, "any"),                 # segments shorter than a window, empty ones
    (8, 8, 0, 2, "hot", 1.0, 29, "any"),
])
def test_segments_equal_the_plain_fold(fh, R, lo, n, P, kind, alpha_hi, seed, expect):
    """PARITY B4: a bin folded in segments on bracketing chains + records + the join equals the sequential blend bit for bit."""
    time = np.float32(1234.5)
    frags, tex0 = make_bin(R, lo, n, kind, seed, alpha_hi)
    if expect == "records" or seed == 22:
        frags[::7, 0] = np.float32(-0.0)
    want = plain_fold(frags, lo, time, tex0)
    got = tex0.copy()
    stats = (C.c_int * 3)()
    fh.fh_fold_segments(frags.ctypes.data_as(_fp), n, lo, R, time, got.ctypes.data_as(_fp), P, stats)
    print("segments stats", R, n, P, kind, list(stats))
    assert same_bits(got, want)
    if expect.startswith("settles"):
        assert stats[0] > 0 and stats[2] == 0
    if expect == "settles_most":
        assert stats[1] < n // 2
    if expect == "records":
        assert stats[0] == 0 and stats[2] == 0 and stats[1] == n - min(n, ((n + P - 1) // P + 63) // 64 * 64)
