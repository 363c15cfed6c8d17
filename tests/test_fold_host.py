"""The warp-level fold kernels (k_splat_fold with its shared-memory chunk streaming and hot-texel worklist,
k_splat_fold_hot with its eight chains per warp) cut out of tendrils_b200/csrc/tb_kernels.cuh unchanged and run on a
CPU emulation of a warp -- one thread per lane, shuffles / votes / __syncwarp as barriers (tests/host_harness/
warp_emu.h), cp.async as a 16-byte copy -- against a plain sequential blend, in the three shapes the product launches
them: the whole grid in place, a chunk of a ring (src != dst, dst2, copy_all) and the strided tiles of the band fold.
Nothing here is used by the product."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_fp = C.POINTER(C.c_float)
_up = C.POINTER(C.c_uint32)

CP_ASYNC_HOST = r'''
inline void cp_async16(void *smem_dst, const void *gmem_src) { std::memcpy(smem_dst, gmem_src, 16); }
inline void cp_async4(void *smem_dst, const void *gmem_src) { std::memcpy(smem_dst, gmem_src, 4); }
inline void cp_async_wait_all() {}
'''


def fold_source(ksrc):
    """The fragment type and the fold kernels of tb_kernels.cuh, the cp.async helpers swapped for plain copies."""
    frag = ksrc[ksrc.index("#ifndef TB_FRAG_BYTES"):ksrc.index("// Line pair k of a column")]
    fold = ksrc[ksrc.index("// Pass 5: ordered alpha-over fold"):ksrc.index("// first index of the sorted key array")]
    a, b = fold.index("__device__ __forceinline__ void cp_async16("), fold.index("// [cp-async-end]")
    assert fold[a:b].count("asm volatile") == 4                              # exactly the cp.async helpers are swapped
    fold = fold[:a] + CP_ASYNC_HOST + fold[b:]
    assert "asm" not in fold.replace("// [cp-async-end]", "")
    return frag.replace("__device__", ""), fold.replace("__device__", "")


def compile_harness(d, name, text, frag_bytes, std="c++20"):
    cpp = d / f"{name}_{frag_bytes}.cpp"
    cpp.write_text(text)
    out = d / f"lib{name}_{frag_bytes}.so"
    subprocess.run(["g++", "-O2", f"-std={std}", "-march=x86-64-v3", "-ffp-contract=off", "-fno-fast-math", "-fPIC", "-shared", "-pthread",
                    f"-DTB_FRAG_BYTES={frag_bytes}", "-Wno-unknown-pragmas", "-Wno-unused-function", "-Wno-attributes",
                    "-I/usr/local/cuda/include", "-I", os.path.join(ROOT, "tests", "host_harness"), "-o", str(out), str(cpp)], check=True)
    return C.CDLL(str(out))

HARNESS = r'''
#include <algorithm>
#include <cstddef>
#include <cuda_runtime.h>
#include "cuda_intrinsics_shim.h"
#include "warp_emu.h"
#define __global__
#define __launch_bounds__(...)
#define __shared__ static
template <class T> static inline T __ldcs(const T *p) { return *p; }
namespace tb {
using std::min; using std::max;
%(frag)s
%(fold)s
}
using namespace tb;

// the two launches of fold() / ring_fold() / bands_fold() in tb_api.cu, one emulated warp at a time
extern "C" long long fh_fold(const uint32_t *seg, const float *vals, float *src, float *dst, float *dst2, int t_begin, int t_end,
                             int copy_all, int tile_first, int tile_stride, int n_warps, float time, uint32_t hot_threshold,
                             uint32_t *hot_scratch /* 2 + G words */) {
    FoldIO io{};
    io.src = (const float4 *)src; io.dst = (float4 *)dst; io.dst2 = (float4 *)dst2;
    io.t_begin = t_begin; io.t_end = t_end; io.copy_all = copy_all; io.tile_first = tile_first; io.tile_stride = tile_stride;
    hot_scratch[0] = hot_scratch[1] = 0;
    tb_host_blockDim = {(unsigned)(kFoldWarps * 32), 1, 1};
    const int blocks = (n_warps + kFoldWarps - 1) / kFoldWarps;
    for (int b = 0; b < blocks; ++b)
        for (int w = 0; w < kFoldWarps; ++w)
            tb_run_warp((unsigned)b, (unsigned)(w * 32), [&] {
                k_splat_fold(io, (const uint2 *)seg, (const FragVal *)vals, time, hot_scratch, hot_scratch + 2, hot_threshold);
            });
    tb_host_blockDim = {(unsigned)(kHotWarps * 32), 1, 1};
    for (int b = 0; b < 3; ++b)                                  // a small persistent grid; the first warps drain the list
        for (int w = 0; w < kHotWarps; ++w)
            tb_run_warp((unsigned)b, (unsigned)(w * 32), [&] {
                k_splat_fold_hot(io, (const uint2 *)seg, (const FragVal *)vals, time, hot_scratch, hot_scratch + 2, hot_scratch + 1);
            });
    return (long long)hot_scratch[0];
}
'''


@pytest.fixture(scope="module", params=[16, 12], ids=["frag16", "frag12"])
def fh(request, tmp_path_factory):
    """The harness built for the default 16-byte fragments and for -DTB_FRAG_BYTES=12 (a build option)."""
    d = tmp_path_factory.mktemp("fh")
    ksrc = open(os.path.join(ROOT, "tendrils_b200", "csrc", "tb_kernels.cuh")).read()
    frag, fold = fold_source(ksrc)
    L = compile_harness(d, "fold_host", HARNESS % {"frag": frag, "fold": fold}, request.param)
    L.fh_fold.restype = C.c_longlong
    L.fh_fold.argtypes = [_up, _fp, _fp, _fp, _fp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_uint32, _up]
    L.frag_floats = request.param // 4
    return L


def make_segments(rng, G, mean, hot_every, hot_len, gap_p=0.3):
    """A sorted fragment array and its segment table: per texel a run of fragments; `gap_p` of the runs start after a
    gap (fragments an opaque one cut away); every `hot_every`-th texel is long."""
    lens = rng.poisson(mean, G)
    lens[rng.random(G) < 0.25] = 0
    if hot_every:
        lens[::hot_every] = rng.integers(hot_len // 2, hot_len + 1, len(lens[::hot_every]))
    seg = np.zeros(2 * G, np.uint32)
    pos = 0
    for t in range(G):
        if lens[t] == 0:
            continue
        if rng.random() < gap_p:
            pos += int(rng.integers(1, 40))
        seg[2 * t], seg[2 * t + 1] = pos, pos + lens[t]
        pos += int(lens[t])
    F = pos + 8
    vals = np.zeros((F, 4), np.float32)
    vals[:, 0:2] = rng.normal(0, 0.01, (F, 2))
    vals[:, 2] = rng.uniform(0, 1, F)
    vals[rng.random(F) < 0.05, 2] = 1.0                                      # opaque fragments inside a run are just fragments
    vals[rng.random(F) < 0.02, 2] = 0.0
    vals[:, 3] = 12345.0                                                     # the pad lane must never matter
    return seg, vals


def plain_fold(seg, vals, src, time, texels):
    out = {}
    time = np.float32(time)
    for t in texels:
        b, e = int(seg[2 * t]), int(seg[2 * t + 1])
        if e <= b:
            continue
        d = src[t].copy()
        for i in range(b, e):
            cx, cy, a = vals[i, 0], vals[i, 1], vals[i, 2]
            c = np.array([cx, cy, time, a], np.float32)
            d = (c * a + d * (np.float32(1.0) - a)).astype(np.float32)       # spec/PARITY.md B2: three roundings per channel
        out[t] = d
    return out


def run(fh, seg, vals, src, dst, dst2, t_begin, t_end, copy_all, tile_first, tile_stride, n_warps, time, hot):
    G = src.shape[0]
    scratch = np.zeros(2 + G, np.uint32)
    vals = np.ascontiguousarray(vals[:, :fh.frag_floats])                    # 16-byte fragments carry a pad lane, 12-byte ones do not
    p = lambda a: None if a is None else a.ctypes.data_as(_fp)
    n_hot = fh.fh_fold(seg.ctypes.data_as(_up), p(vals), p(src), p(dst), p(dst2), t_begin, t_end, copy_all, tile_first, tile_stride,
                       n_warps, np.float32(time), hot, scratch.ctypes.data_as(_up))
    return n_hot


@pytest.mark.parametrize("G,mean,hot_every,hot_len,threshold", [(256, 6, 37, 700, 96), (200, 20, 0, 0, 96), (96, 3, 5, 300, 8),
                                                                (33, 40, 11, 1500, 96), (64, 1, 0, 0, 0xffffffff)])
def test_whole_grid_in_place(fh, G, mean, hot_every, hot_len, threshold):
    rng = np.random.default_rng(G + mean)
    seg, vals = make_segments(rng, G, mean, hot_every, hot_len)
    flow0 = rng.normal(0, 0.01, (G, 4)).astype(np.float32)
    flow = flow0.copy()
    n_hot = run(fh, seg, vals, flow, flow, None, 0, G, 0, 0, 1, (G + 31) // 32, 77.5, threshold)
    want = flow0.copy()
    for t, d in plain_fold(seg, vals, flow0, 77.5, range(G)).items():
        want[t] = d
    assert np.array_equal(flow.view(np.uint32), want.view(np.uint32))
    lens = seg[1::2].astype(np.int64) - seg[0::2]
    assert n_hot == int((lens > threshold).sum())
    if hot_every and threshold < 0xffffffff:
        assert n_hot > 0


@pytest.mark.parametrize("t_begin,t_end", [(0, 128), (128, 301), (64, 96)])
def test_ring_chunk_copies_and_forwards(fh, t_begin, t_end):
    """ring_fold: src = the inbox, dst = the next rank's inbox, dst2 = rank 0's grid, every texel of the chunk written."""
    G = 301
    rng = np.random.default_rng(t_end)
    seg, vals = make_segments(rng, G, 8, 23, 400)
    src = rng.normal(0, 0.01, (G, 4)).astype(np.float32)
    dst, dst2 = np.full((G, 4), 7.0, np.float32), np.full((G, 4), 9.0, np.float32)
    n_warps = (t_end - t_begin + 31) // 32
    run(fh, seg, vals, src, dst, dst2, t_begin, t_end, 1, 0, 1, n_warps, 3.25, 96)
    want = np.full((G, 4), 7.0, np.float32)
    want[t_begin:t_end] = src[t_begin:t_end]
    for t, d in plain_fold(seg, vals, src, 3.25, range(t_begin, t_end)).items():
        want[t] = d
    assert np.array_equal(dst.view(np.uint32), want.view(np.uint32))
    want2 = np.full((G, 4), 9.0, np.float32)
    want2[t_begin:t_end] = want[t_begin:t_end]
    assert np.array_equal(dst2.view(np.uint32), want2.view(np.uint32))


@pytest.mark.parametrize("world,rank,G", [(2, 1, 256), (3, 0, 250), (3, 2, 250), (8, 5, 1000), (5, 1, 33), (5, 4, 33)])
def test_band_tiles(fh, world, rank, G):
    """bands_fold: this rank owns the 32-texel tiles with tile % world == rank and must leave every other texel alone."""
    rng = np.random.default_rng(world * 100 + rank)
    seg, vals = make_segments(rng, G, 10, 13, 500)
    flow0 = rng.normal(0, 0.01, (G, 4)).astype(np.float32)
    flow = flow0.copy()
    tiles = (G + 31) // 32
    mine = (tiles - rank + world - 1) // world
    run(fh, seg, vals, flow, flow, None, 0, G, 0, rank, world, mine, 5.0, 96)
    owned = [t for t in range(G) if (t // 32) % world == rank]
    want = flow0.copy()
    for t, d in plain_fold(seg, vals, flow0, 5.0, owned).items():
        want[t] = d
    assert np.array_equal(flow.view(np.uint32), want.view(np.uint32))
    assert len(owned) < 40 or any(seg[2 * t + 1] > seg[2 * t] for t in owned)


STRESS = int(os.environ.get("TB_STRESS", "0"))        # TB_STRESS=n: n extra random configurations (off in the normal run)


@pytest.mark.parametrize("seed", range(STRESS))
def test_stress_random_configurations(fh, seed):
    rng = np.random.default_rng(10_000 + seed)
    G = int(rng.integers(1, 700))
    seg, vals = make_segments(rng, G, float(rng.choice([0.5, 3, 12, 40])), int(rng.choice([0, 3, 17, 64])), int(rng.integers(97, 2500)),
                              gap_p=float(rng.choice([0.0, 0.3, 0.9])))
    threshold = int(rng.choice([0, 8, 96, 96, 96, 0xffffffff]))
    world = int(rng.integers(1, 9))
    rank = int(rng.integers(0, world))
    flow0 = rng.normal(0, 0.01, (G, 4)).astype(np.float32)
    flow = flow0.copy()
    tiles = (G + 31) // 32
    mine = (tiles - rank + world - 1) // world
    run(fh, seg, vals, flow, flow, None, 0, G, 0, rank, world, mine, 9.75, threshold)
    want = flow0.copy()
    for t, d in plain_fold(seg, vals, flow0, 9.75, [t for t in range(G) if (t // 32) % world == rank]).items():
        want[t] = d
    assert np.array_equal(flow.view(np.uint32), want.view(np.uint32)), (seed, G, threshold, world, rank)


@pytest.mark.parametrize("seed", range(STRESS))
def test_stress_ring_chunks(fh, seed):
    rng = np.random.default_rng(70_000 + seed)
    G = int(rng.integers(1, 600))
    seg, vals = make_segments(rng, G, float(rng.choice([0.5, 5, 25])), int(rng.choice([0, 7, 40])), int(rng.integers(97, 1500)))
    chunks = int(rng.integers(1, 6))
    per = ((G + chunks - 1) // chunks + 127) // 128 * 128                  # ring_fold: texels per chunk, CTA aligned
    src = rng.normal(0, 0.01, (G, 4)).astype(np.float32)
    dst, dst2 = np.full((G, 4), 7.0, np.float32), np.full((G, 4), 9.0, np.float32)
    use2 = bool(rng.random() < 0.5)
    want = np.full((G, 4), 7.0, np.float32)
    for k in range(chunks):
        t0, t1 = min(G, k * per), min(G, (k + 1) * per)
        if t0 >= t1:
            continue
        run(fh, seg, vals, src, dst, dst2 if use2 else None, t0, t1, 1, 0, 1, (t1 - t0 + 127) // 128 * 4, 1.5, 96)
        want[t0:t1] = src[t0:t1]
        for t, d in plain_fold(seg, vals, src, 1.5, range(t0, t1)).items():
            want[t] = d
    assert np.array_equal(dst.view(np.uint32), want.view(np.uint32)), seed
    if use2:
        assert np.array_equal(dst2.view(np.uint32), want.view(np.uint32)), seed
