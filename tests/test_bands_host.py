"""The band fold of a sharded run (tb_splat_fold_bands) end to end on the CPU: its kernels -- k_bands_lengths,
k_bands_offsets, k_bands_push, the strided k_splat_fold / k_splat_fold_hot, k_bands_publish -- cut out of
tendrils_b200/csrc/tb_kernels.cuh unchanged, with N "ranks" living in one process (a peer pointer is just a pointer)
and the phases run in the order the barriers enforce.  Every rank's grid must equal the plain fold of all ranks'
fragments, per texel in rank order -- for world sizes and grid sizes the 2-GPU parity test cannot reach.
Nothing here is used by the product."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from test_fold_host import compile_harness, fold_source, make_segments

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_fp = C.POINTER(C.c_float)
_up = C.POINTER(C.c_uint32)

HARNESS = r'''
#include <algorithm>
#include <cstddef>
#include <vector>
#include <cuda_runtime.h>
#include "cuda_intrinsics_shim.h"
#include "warp_emu.h"
#define __global__
#define __launch_bounds__(...)
#define __shared__ static
template <class T> static inline T __ldcs(const T *p) { return *p; }
namespace tb {
using std::min; using std::max;
constexpr uint32_t kOpaqueBit = 0x80000000u;
%(frag)s
%(fold)s
%(bands)s
}
using namespace tb;

template <class F> static void threads(long long n, unsigned per_block, F body) {     // kernels without warp intrinsics
    tb_host_blockDim = {per_block, 1, 1};
    for (long long b = 0; b < (n + per_block - 1) / per_block; ++b)
        for (unsigned t = 0; t < per_block; ++t) { tb_host_blockIdx = {(unsigned)b, 0, 0}; tb_host_threadIdx = {t, 0, 0}; body(); }
}

// world ranks in one process.  seg[j], keys[j], vals[j] (n_frag[j] fragments): rank j's sorted fragments; flow: world grids.
extern "C" int bh_bands_fold(int world, int G, const uint32_t *const *seg, const uint32_t *const *keys, const float *const *vals,
                             const uint32_t *n_frag, float *const *flow, uint32_t cap, float time, uint32_t hot_threshold) {
    const int tiles = (G + 31) / 32;
    std::vector<std::vector<FragVal>> merged(world, std::vector<FragVal>(cap));
    std::vector<std::vector<uint32_t>> dst(world, std::vector<uint32_t>(G, 0xdeadbeefu)), seg_m(world, std::vector<uint32_t>(2 * (size_t)G, 0u));
    BandSources S{}; BandSinks D{}; BandPeers P{};
    for (int j = 0; j < world; ++j) {
        S.seg[j] = (const uint2 *)seg[j]; D.merged[j] = merged[j].data(); D.dst[j] = dst[j].data(); P.flow[j] = (float4 *)flow[j];
    }
    P.n = world;
    int overflow = 0;
    // (barrier 0)  owners: lengths -> scan -> offsets
    for (int r = 0; r < world; ++r) {
        const int mine = (tiles - r + world - 1) / world;
        const long long warps = (long long)mine * world;
        std::vector<uint32_t> len((size_t)mine * 32 * world + 1, 0u), off(len.size());
        threads(warps * 32, 256, [&] { k_bands_lengths(S, world, r, mine, G, len.data()); });
        uint32_t run = 0;
        for (size_t i = 0; i < len.size(); ++i) { off[i] = run; run += len[i]; }
        threads(warps * 32, 256, [&] { k_bands_offsets(D, world, r, mine, G, off.data(), cap, seg_m[r].data(), &overflow); });
    }
    if (overflow) return -1;
    // (barrier 1)  sources: push
    for (int j = 0; j < world; ++j)
        if (n_frag[j])
            threads(n_frag[j], 256, [&] { k_bands_push(keys[j], (const FragVal *)vals[j], n_frag[j], (const uint2 *)seg[j], dst[j].data(), D, world, cap); });
    // (barrier 2)  owners: fold their tiles of the merged array, publish
    for (int r = 0; r < world; ++r) {
        const int mine = (tiles - r + world - 1) / world;
        FoldIO io{};
        io.src = (const float4 *)flow[r]; io.dst = (float4 *)flow[r]; io.dst2 = nullptr;
        io.t_begin = 0; io.t_end = G; io.copy_all = 0; io.tile_first = r; io.tile_stride = world;
        std::vector<uint32_t> hot(2 + (size_t)G, 0u);
        tb_host_blockDim = {(unsigned)(kFoldWarps * 32), 1, 1};
        for (int b = 0; b < (mine + kFoldWarps - 1) / kFoldWarps; ++b)
            for (int w = 0; w < kFoldWarps; ++w)
                tb_run_warp((unsigned)b, (unsigned)(w * 32), [&] {
                    k_splat_fold(io, (const uint2 *)seg_m[r].data(), merged[r].data(), time, hot.data(), hot.data() + 2, hot_threshold);
                });
        tb_host_blockDim = {(unsigned)(kHotWarps * 32), 1, 1};
        for (int w = 0; w < kHotWarps; ++w)
            tb_run_warp(0u, (unsigned)(w * 32), [&] {
                k_splat_fold_hot(io, (const uint2 *)seg_m[r].data(), merged[r].data(), time, hot.data(), hot.data() + 2, hot.data() + 1);
            });
    }
    for (int r = 0; r < world; ++r) {                            // publish after ALL folds: a rank folds against its own grid only
        const int mine = (tiles - r + world - 1) / world;
        P.me = r;
        threads((long long)mine * 32, 256, [&] { k_bands_publish((const float4 *)flow[r], P, G); });
    }
    return 0;
}
'''


@pytest.fixture(scope="module", params=[16, 12], ids=["frag16", "frag12"])
def bh(request, tmp_path_factory):
    d = tmp_path_factory.mktemp("bh")
    ksrc = open(os.path.join(ROOT, "tendrils_b200", "csrc", "tb_kernels.cuh")).read()
    frag, fold = fold_source(ksrc)
    bands = (ksrc[ksrc.index("// Band fold of a sharded run"):ksrc.index("// all-rank barrier over peer memory")] +
             ksrc[ksrc.index("// copy this rank's finished tiles"):ksrc.index("// Full-grid alpha-over of an RGBA layer")])
    L = compile_harness(d, "bands_host", HARNESS % {"frag": frag, "fold": fold, "bands": bands.replace("__device__", "")}, request.param)
    L.bh_bands_fold.restype = C.c_int
    L.bh_bands_fold.argtypes = [C.c_int, C.c_int, C.POINTER(_up), C.POINTER(_up), C.POINTER(_fp), _up, C.POINTER(_fp), C.c_uint32,
                                C.c_float, C.c_uint32]
    L.frag_floats = request.param // 4
    return L


def keys_of(seg, n):
    """The sorted key array behind a segment table: every fragment's texel, the cut-away ones (the gap before a run)
    included -- they belong to the texel whose run follows them."""
    keys = np.zeros(n, np.uint32)
    pos = 0
    for t in range(seg.shape[0] // 2):
        b, e = int(seg[2 * t]), int(seg[2 * t + 1])
        if e > b:
            keys[pos:e] = t
            pos = e
    keys[pos:] = max(seg.shape[0] // 2 - 1, 0)
    return keys


@pytest.mark.parametrize("world,G,mean,hot_every,cap_slack", [(2, 256, 6, 37, 64), (3, 250, 10, 13, 64), (8, 1000, 4, 101, 64),
                                                               (5, 33, 30, 4, 64), (4, 128, 0, 0, 64), (3, 250, 10, 13, -1)])
def test_band_fold_equals_plain_fold_in_rank_order(bh, world, G, mean, hot_every, cap_slack):
    rng = np.random.default_rng(world * 1000 + G)
    segs, keys, vals = [], [], []
    for j in range(world):
        s, v = make_segments(rng, G, mean, hot_every if j % 2 == 0 else 0, 400) if mean else (np.zeros(2 * G, np.uint32), np.zeros((8, 4), np.float32))
        n = int(s[1::2].max()) if mean else 0
        segs.append(s); vals.append(np.ascontiguousarray(v)); keys.append(keys_of(s, max(n, 1)))
    n_frag = np.array([int(s[1::2].max()) for s in segs], np.uint32)
    flow0 = rng.normal(0, 0.01, (G, 4)).astype(np.float32)
    flows = [flow0.copy() for _ in range(world)]
    # the plain fold: per texel the sources side by side in rank order
    want = flow0.copy()
    time = np.float32(42.5)
    kept = np.zeros(world, np.int64)
    for t in range(G):
        d = want[t].copy()
        touched = False
        for j in range(world):
            for i in range(int(segs[j][2 * t]), int(segs[j][2 * t + 1])):
                a = vals[j][i, 2]
                c = np.array([vals[j][i, 0], vals[j][i, 1], time, a], np.float32)
                d = (c * a + d * (np.float32(1.0) - a)).astype(np.float32)
                touched = True
                kept[(t // 32) % world] += 1
        if touched:
            want[t] = d
    cap = int(kept.max()) + cap_slack                                        # -1: one owner's merged fragments do not fit
    arr = lambda ptr_t, xs, cast: (ptr_t * world)(*[x.ctypes.data_as(cast) for x in xs])
    packed = [np.ascontiguousarray(v[:, :bh.frag_floats]) for v in vals]     # the harness's fragment layout (with / without pad lane)
    rc = bh.bh_bands_fold(world, G, arr(_up, segs, _up), arr(_up, keys, _up), arr(_fp, packed, _fp), n_frag.ctypes.data_as(_up),
                          arr(_fp, flows, _fp), max(cap, 1), time, 96)
    if cap_slack < 0:
        assert rc == -1                                                      # the overflow flag the host turns into an error
        return
    assert rc == 0
    for r in range(world):
        assert np.array_equal(flows[r].view(np.uint32), want.view(np.uint32)), f"rank {r}"


STRESS = int(os.environ.get("TB_STRESS", "0"))        # TB_STRESS=n: n extra random configurations (off in the normal run)


@pytest.mark.parametrize("seed", range(STRESS))
def test_stress_random_configurations(bh, seed):
    rng = np.random.default_rng(30_000 + seed)
    world = int(rng.integers(2, 17))
    G = int(rng.integers(1, 900))
    test_band_fold_equals_plain_fold_in_rank_order(bh, world, G, float(rng.choice([0, 0.5, 4, 15])), int(rng.choice([0, 5, 29, 101])),
                                                   int(rng.choice([0, 1, 64])))
