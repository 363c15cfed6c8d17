"""The collect half of the ordered flow splat -- the D6 pair table (build_pairs, host code of tb_api.cu), the count pass
both ways (k_splat_count and the count fused into k_integrate), k_splat_emit, k_splat_bounds with its opaque cut --
cut out of the product source unchanged, compiled for the CPU by this test, completed with a plain stable sort and a
plain sequential blend, and compared bit for bit with the oracle's splat on hostile states.  The device-only parts
(CUB scan / sort, the warp-level fold kernels) are what the GPU parity tests cover.  Nothing here is used by the product."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from test_integrate_host import hostile_state
from test_math_host import PACKED_HOST_PRIMITIVES

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_fp = C.POINTER(C.c_float)

HARNESS = r'''
#include <algorithm>
#include <cstddef>
#include <numeric>
#include <vector>
#include <cuda_runtime.h>
#include "cuda_intrinsics_shim.h"
#include "%(math)s"
#include "%(noise)s"
#include "%(abi)s"
struct HostIdx { unsigned x, y, z; };
static HostIdx tb_host_blockIdx, tb_host_threadIdx, tb_host_blockDim;
#define blockIdx tb_host_blockIdx
#define threadIdx tb_host_threadIdx
#define blockDim tb_host_blockDim
#define __global__
#define __launch_bounds__(...)
template <class T> static inline T __ldcs(const T *p) { return *p; }
template <class T> static inline T __ldg(const T *p) { return *p; }
template <class T, class V> static inline void __stcs(T *p, V v) { *p = v; }
static inline uint32_t atomicMax(uint32_t *p, uint32_t v) { const uint32_t o = *p; if (v > o) *p = v; return o; }
namespace tb {
static constexpr float kInert = -1000000.0f;
%(kernels)s
namespace {
%(pairs)s
}
}
using namespace tb;

template <class K, class A> static void launch(K kernel, const A &args, long long threads, unsigned by = 1) {
    tb_host_blockDim = {256, 1, 1};
    for (unsigned y = 0; y < by; ++y)
        for (long long b = 0; b < (threads + 255) / 256; ++b)
            for (unsigned t = 0; t < 256; ++t) {
                tb_host_blockIdx = {(unsigned)b, y, 0}; tb_host_threadIdx = {t, 0, 0};
                kernel(args);
            }
}

// the D6 pair table of the product (build_pairs), flattened: k, row_a, cur_a, row_b, cur_b per active pair
extern "C" int sh_pairs(int PH, int *out5, int cap) {
    const std::vector<PairEntry> pairs = build_pairs(PH);
    if ((int)pairs.size() > cap) return -1;
    for (size_t i = 0; i < pairs.size(); ++i) {
        out5[5 * i] = pairs[i].k; out5[5 * i + 1] = pairs[i].row_a & 0x7fffffff; out5[5 * i + 2] = pairs[i].row_a < 0;
        out5[5 * i + 3] = pairs[i].row_b & 0x7fffffff; out5[5 * i + 4] = pairs[i].row_b < 0;
    }
    return (int)pairs.size();
}

// One integrate step prev -> cur (so that the count can ride in it), then the splat of (cur, prev) into flow.
// Returns the fragment count, -1 if the fused count disagrees with k_splat_count.
extern "C" long long sh_step_and_splat(const float *state18, int PW, int PH, int W, int H, const float *prev, float *cur,
                                       const float *targets, float *flow, float time, float dt, long long *kept, int partial, int *mode) {
    tb_state S; std::memcpy(&S, state18, sizeof(S));
    const std::vector<PairEntry> pairs = build_pairs(PH);
    const int n_pairs = (int)pairs.size();
    const long long n_prims = (long long)PW * n_pairs;
    // row -> pair table of the fused count: tb_create (tb_api.cu), checked against the source text by the test
    std::vector<int32_t> rp((size_t)PH, -1), odd;
    for (size_t k = 0; k < pairs.size(); ++k) {
        const int ra = pairs[k].row_a & 0x7fffffff, rb = pairs[k].row_b & 0x7fffffff;
        const bool ca = pairs[k].row_a < 0, cb = pairs[k].row_b < 0;
        const bool rides = ra == rb && ca != cb && k < (1u << 30) && rp[(size_t)ra] == -1;
        if (rides) rp[(size_t)ra] = (int32_t)((uint32_t)k | ((cb ? 1u : 2u) << 30));
        else odd.push_back((int32_t)k);
    }
    const bool fuse_count = !pairs.empty() && odd.empty();
    const bool fuse_partial = partial && !pairs.empty() && !odd.empty() && odd.size() * 16 <= pairs.size();
    const bool fuse = fuse_count || fuse_partial;
    *mode = fuse_count ? 1 : (fuse_partial ? 2 : 0);
    std::vector<uint32_t> fused((size_t)n_prims + 1, 0u), counted((size_t)n_prims + 1, 0u);
    IntegrateArgs I{};
    I.S = S; I.in = (const float4 *)prev; I.out = (float4 *)cur; I.targets = (const float4 *)targets; I.flow = (const float4 *)flow;
    I.PW = PW; I.PH = PH; I.W = W; I.H = H; I.col0 = 0; I.cols = PW; I.time = time; I.dt = dt;
    I.use_targets = 1; I.use_noise = 1; I.packed_noise = 1; I.pow2_res = 0;
    I.pk.one = 1.0f; I.pk.neg_one = -1.0f; I.pk.neg_zero = -0.0f;
    I.row_pair = rp.data(); I.prim_off = fuse ? fused.data() : nullptr; I.n_pairs = n_pairs;
    launch(k_integrate<kFused>, I, PH, (unsigned)PW);
    std::vector<uint32_t> keys0, keys1; std::vector<FragVal> vals0, vals1;
    SplatArgs A{};
    A.cur = (const float4 *)cur; A.prev = (const float4 *)prev; A.pairs = pairs.data(); A.n_pairs = n_pairs; A.PH = PH; A.cols = PW;
    A.W = W; A.H = H; A.vsx = S.viewSize[0]; A.vsy = S.viewSize[1]; A.speedLimit = S.speedLimit;
    A.prim_off = counted.data();
    if (fuse_partial) {                                        // tb_step: the pairs that could not ride in k_integrate
        SplatArgs O2 = A; O2.prim_off = fused.data();
        tb_host_blockDim = {256, 1, 1};
        const long long n = (long long)PW * (long long)odd.size();
        for (long long b = 0; b < (n + 255) / 256; ++b)
            for (unsigned t = 0; t < 256; ++t) { tb_host_blockIdx = {(unsigned)b, 0, 0}; tb_host_threadIdx = {t, 0, 0}; k_splat_count_odd(O2, odd.data(), (int)odd.size()); }
    }
    launch(k_splat_count, A, n_prims);
    if (fuse && fused != counted) return -1;
    uint32_t total = 0;                                        // exclusive scan; slot n_prims holds the total
    for (long long i = 0; i <= n_prims; ++i) { const uint32_t c = counted[(size_t)i]; counted[(size_t)i] = total; total += c; }
    keys0.resize(total); vals0.resize(total);
    A.keys = keys0.data(); A.vals = vals0.data(); A.cap = total; A.total = &counted[(size_t)n_prims];
    launch(k_splat_emit, A, n_prims);
    std::vector<uint32_t> order(total);
    std::iota(order.begin(), order.end(), 0u);
    std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return (keys0[a] & ~kOpaqueBit) < (keys0[b] & ~kOpaqueBit); });
    keys1.resize(total); vals1.resize(total);
    for (uint32_t i = 0; i < total; ++i) { keys1[i] = keys0[order[i]]; vals1[i] = vals0[order[i]]; }
    const size_t G = (size_t)W * H;
    std::vector<uint32_t> seg(2 * G, 0u);
    if (total) {
        tb_host_blockDim = {256, 1, 1};
        for (long long b = 0; b < ((total + 3) / 4 + 255) / 256; ++b)
            for (unsigned t = 0; t < 256; ++t) {
                tb_host_blockIdx = {(unsigned)b, 0, 0}; tb_host_threadIdx = {t, 0, 0};
                k_splat_bounds(keys1.data(), total, seg.data());
            }
    }
    *kept = 0;
    for (size_t t = 0; t < G; ++t)                              // the plain fold: spec/PARITY.md B2, from the cut of B3
        for (uint32_t i = seg[2 * t]; i < seg[2 * t + 1]; ++i) {
            const FragVal f = vals1[i];
            const float c[4] = {f.cx, f.cy, time, f.a}, om = 1.0f - f.a;
            for (int k = 0; k < 4; ++k) { const float t1 = c[k] * f.a, t2 = flow[4 * t + k] * om; flow[4 * t + k] = t1 + t2; }
            ++*kept;
        }
    return (long long)total;
}
'''


@pytest.fixture(scope="module", params=[16, 12], ids=["frag16", "frag12"])
def sh(request, tmp_path_factory):
    d = tmp_path_factory.mktemp("sh")
    csrc = os.path.join(ROOT, "tendrils_b200", "csrc")
    math = d / "tb_math_host.cuh"
    math.write_text(open(os.path.join(csrc, "tb_math.cuh")).read().replace("__device__", ""))
    nsrc = open(os.path.join(csrc, "tb_noise2.cuh")).read()
    a, b = nsrc.index("__device__ __forceinline__ F2 pack2("), nsrc.index("struct P2 {")
    noise = d / "tb_noise2_host.cuh"
    noise.write_text((nsrc[:a] + PACKED_HOST_PRIMITIVES + nsrc[b:]).replace("__device__", ""))
    ksrc = open(os.path.join(csrc, "tb_kernels.cuh")).read()
    kernels = ksrc[ksrc.index("#ifndef TB_FRAG_BYTES"):ksrc.index("// Pass 5: ordered alpha-over fold")].replace("__device__", "")
    asrc = open(os.path.join(csrc, "tb_api.cu")).read()
    pairs = asrc[asrc.index("int host_texel(float u, int size) {"):asrc.index("// column sampled by vertex column i")]
    # the harness repeats tb_create's row -> pair loop: make sure the product still has it verbatim
    for line in ("const bool rides = ra == rb && ca != cb && k < (1u << 30) && rp[static_cast<size_t>(ra)] == -1;",
                 "else odd.push_back(static_cast<int32_t>(k));",
                 "c->fuse_count = !pairs.empty() && odd.empty();",
                 "!pairs.empty() && !odd.empty() && odd.size() * 16 <= pairs.size();"):
        assert line in asrc, line
    cpp = d / "splat_host.cpp"
    cpp.write_text(HARNESS % {"math": str(math), "noise": str(noise), "abi": os.path.join(ROOT, "include", "tendrils_b200.h"),
                              "kernels": kernels, "pairs": pairs})
    out = d / "libsplat_host.so"
    subprocess.run(["g++", "-O2", "-std=c++17", "-march=x86-64-v3", "-ffp-contract=off", "-fno-fast-math", "-fPIC", "-shared",
                    f"-DTB_FRAG_BYTES={request.param}",                       # 12: the build option without the pad lane
                    "-Wno-unknown-pragmas", "-Wno-unused-function", "-Wno-attributes", "-I/usr/local/cuda/include",
                    "-I", os.path.join(ROOT, "tests", "host_harness"), "-o", str(out), str(cpp)], check=True)
    L = C.CDLL(str(out))
    L.frag_bytes = request.param
    L.sh_pairs.restype = C.c_int
    L.sh_pairs.argtypes = [C.c_int, C.POINTER(C.c_int), C.c_int]
    L.sh_step_and_splat.restype = C.c_longlong
    L.sh_step_and_splat.argtypes = [_fp, C.c_int, C.c_int, C.c_int, C.c_int, _fp, _fp, _fp, _fp, C.c_float, C.c_float,
                                    C.POINTER(C.c_longlong), C.c_int, C.POINTER(C.c_int)]
    return L


@pytest.mark.parametrize("PW,PH,W,H,speed_limit,seed", [(32, 64, 24, 16, 0.2, 1), (20, 50, 40, 40, 0.3, 2), (5, 7, 3, 2, 0.2, 3),
                                                        (48, 48, 64, 8, 0.3, 4), (9, 64, 1, 1, 0.01, 5), (8, 1, 16, 16, 0.1, 6),
                                                        (64, 64, 8, 8, 0.5, 7), (16, 32, 24, 16, 0.01, 8),
                                                        # heights whose D6 table draws some rows TWICE: no fused count there
                                                        (30, 47, 56, 63, 2.0, 9), (20, 83, 32, 32, 0.5, 10), (24, 23, 16, 16, 0.3, 11)])
def test_collect_on_host_equals_oracle(sh, oracle, PW, PH, W, H, speed_limit, seed, partial=0, want_mode=None):
    rng = np.random.default_rng(seed)
    O = oracle
    P = O.make_params(viewSize=(1.0, W / H) if W >= H else (H / W, 1.0), speedLimit=speed_limit, target=0.001)
    S = np.array([getattr(P, n) for n, _ in P._fields_[:16]] + [P.viewSize[0], P.viewSize[1]], np.float32)
    prev = hostile_state(rng, PW, PH)
    prev[..., 2:4] *= np.float32(speed_limit / 0.006)                       # some particles at the speed limit: opaque fragments
    targets = np.zeros((PW, PH, 4), np.float32)
    flow0 = rng.normal(0, 0.01, (H, W, 4)).astype(np.float32)
    flow0[..., 2] = rng.uniform(0, 100, (H, W))
    time, dt = np.float32(117.0), np.float32(1000 / 60)
    p = lambda a: a.ctypes.data_as(_fp)
    stats = []
    for step in range(3):                                                   # the flow written by one step feeds the next
        with np.errstate(all="ignore"):
            want_cur = O.integrate(P, prev, targets, flow0, time, dt)
            want_flow = flow0.copy()
            n = O.splat(P, want_cur, prev, want_flow, time)
        cur = np.zeros_like(prev)
        flow = flow0.copy()
        kept = C.c_longlong()
        mode = C.c_int()
        got_n = sh.sh_step_and_splat(p(S), PW, PH, W, H, p(prev), p(cur), p(targets), p(flow), time, dt, C.byref(kept), partial,
                                     C.byref(mode))
        assert want_mode is None or mode.value == want_mode
        assert got_n != -1, "the count fused into k_integrate disagrees with k_splat_count"
        stats.append((n, kept.value))
        assert got_n == n and 0 <= kept.value <= n
        same = lambda a, b: (np.isnan(a) == np.isnan(b)).all() and np.array_equal(a[~np.isnan(a)], b[~np.isnan(b)])
        assert same(cur, want_cur), f"state, step {step}"
        assert same(flow, want_flow), f"flow, step {step}"
        prev, flow0, time = want_cur, want_flow, np.float32(time + dt)
    if seed in (1, 2, 4, 7):                                                # the substantial cases: many fragments, and the cut bites
        assert min(n for n, _ in stats) > 200 and any(k < n for n, k in stats), stats


@pytest.mark.parametrize("PW,PH", [(6, 43), (4, 133), (3, 1000), (2, 8192)])
def test_partial_fused_count(sh, oracle, PW, PH):
    """TB_FUSE_PARTIAL (experimental): the same-row pairs ride in k_integrate, k_splat_count_odd counts the few others
    (rows drawn twice, pairs joining two rows): together they must equal the full count pass, and the splat the oracle."""
    test_collect_on_host_equals_oracle(sh, oracle, PW, PH, 32, 24, 0.3, 500 + PH, partial=1, want_mode=2)
    test_collect_on_host_equals_oracle(sh, oracle, PW, PH, 32, 24, 0.3, 500 + PH, partial=0, want_mode=0)


STRESS = int(os.environ.get("TB_STRESS", "0"))        # TB_STRESS=n: n extra random configurations (off in the normal run)


@pytest.mark.parametrize("seed", range(STRESS))
def test_stress_random_configurations(sh, oracle, seed):
    rng = np.random.default_rng(20_000 + seed)
    PW, PH = int(rng.integers(1, 40)), int(rng.integers(1, 90))
    W, H = int(rng.integers(1, 70)), int(rng.integers(1, 70))
    test_collect_on_host_equals_oracle(sh, oracle, PW, PH, W, H, float(rng.choice([0.01, 0.1, 0.5, 2.0])), 1000 + seed,
                                       partial=seed % 2)


def test_pair_table_equals_oracle_vertex_table(sh, oracle):
    """build_pairs (tb_api.cu) against the oracle's vertex table (D6) for every height up to 3000 and the large ones."""
    if sh.frag_bytes != 16:
        pytest.skip("host table: independent of the fragment layout")
    for PH in list(range(1, 3001)) + [4096, 5000, 8192, 12288, 16384, 28672, 32768, 65536, 100003]:
        row, cur = oracle.vertex_table(PH)
        want = [(k, int(row[2 * k]), int(cur[2 * k]), int(row[2 * k + 1]), int(cur[2 * k + 1])) for k in range(PH)
                if not (row[2 * k] == row[2 * k + 1] and cur[2 * k] == cur[2 * k + 1])]
        buf = (C.c_int * (5 * PH))()
        n = sh.sh_pairs(PH, buf, PH)
        got = [tuple(buf[5 * i:5 * i + 5]) for i in range(n)]
        assert got == want, PH
