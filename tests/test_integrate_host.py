"""The whole integrate kernel (k_integrate of tendrils_b200/csrc/tb_kernels.cuh: logic.frag with its fast paths, the
packed twin noise, the fused / split launch shapes) compiled for the CPU by this test and compared bit for bit with
the oracle on adversarial states -- NaN, Inf, huge, inert, motionless particles, hostile targets and flow texels --
under every combination of the host-side shortcuts (spec/PARITY.md I5).  The kernel text is cut out of the product
source unchanged; blockIdx / threadIdx, the cache-hinted loads and the three PTX primitives of the packed noise are
shimmed.  Nothing here is used by the product."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from test_math_host import PACKED_HOST_PRIMITIVES

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_fp = C.POINTER(C.c_float)

HARNESS = r'''
#include <cstddef>
#include <cuda_runtime.h>
#include "cuda_intrinsics_shim.h"
#include "%(math)s"
#include "%(noise)s"
#include "%(abi)s"
struct HostIdx { unsigned x, y, z; };
static HostIdx tb_host_blockIdx, tb_host_threadIdx;
#define blockIdx tb_host_blockIdx
#define threadIdx tb_host_threadIdx
#define __global__
#define __launch_bounds__(...)
template <class T> static inline T __ldcs(const T *p) { return *p; }
template <class T> static inline T __ldg(const T *p) { return *p; }
template <class T, class V> static inline void __stcs(T *p, V v) { *p = v; }
namespace tb {
static constexpr float kInert = -1000000.0f;
%(body)s
}
// flags exactly as tb_step derives them (tb_api.cu); `split` runs the noise launch then the finish launch
extern "C" void ih_integrate(const float *state18, int PW, int PH, int W, int H, const float *in, float *out, const float *targets,
                             const float *flow, float time, float dt, int use_targets, int use_noise, int packed, int split) {
    tb::IntegrateArgs A{};
    std::memcpy(&A.S, state18, sizeof(tb_state));
    A.in = reinterpret_cast<const float4 *>(in); A.out = reinterpret_cast<float4 *>(out);
    A.targets = reinterpret_cast<const float4 *>(targets); A.flow = reinterpret_cast<const float4 *>(flow);
    A.PW = PW; A.PH = PH; A.W = W; A.H = H; A.col0 = 0; A.cols = PW; A.time = time; A.dt = dt;
    A.use_targets = use_targets; A.use_noise = use_noise; A.packed_noise = packed;
    auto pow2 = [](int v) { return v > 0 && (v & (v - 1)) == 0; };
    A.pow2_res = pow2(PW) && pow2(PH);
    A.inv_resx = 1.0f / (float)PW; A.inv_resy = 1.0f / (float)PH; A.inv_n = 1.0f / ((float)PW * (float)PH);
    A.pk.one = 1.0f; A.pk.neg_one = -1.0f; A.pk.neg_zero = -0.0f;
    float2 *wander = new float2[(size_t)PW * PH];
    A.wander = wander;
    auto launch = [&](int mode) {
        for (unsigned by = 0; by < (unsigned)PW; ++by)
            for (unsigned bx = 0; bx < (unsigned)((PH + 255) / 256); ++bx)
                for (unsigned tx = 0; tx < 256; ++tx) {
                    tb_host_blockIdx = {bx, by, 0}; tb_host_threadIdx = {tx, 0, 0};
                    if (mode == 0) tb::k_integrate<tb::kFused>(A);
                    if (mode == 1) tb::k_integrate<tb::kNoise>(A);
                    if (mode == 2) tb::k_integrate<tb::kFinish>(A);
                }
    };
    if (split) { launch(1); launch(2); } else launch(0);
    delete[] wander;
}
'''


@pytest.fixture(scope="module")
def ih(tmp_path_factory):
    d = tmp_path_factory.mktemp("ih")
    csrc = os.path.join(ROOT, "tendrils_b200", "csrc")
    math = d / "tb_math_host.cuh"
    math.write_text(open(os.path.join(csrc, "tb_math.cuh")).read().replace("__device__", ""))
    nsrc = open(os.path.join(csrc, "tb_noise2.cuh")).read()
    a, b = nsrc.index("__device__ __forceinline__ F2 pack2("), nsrc.index("struct P2 {")
    noise = d / "tb_noise2_host.cuh"
    noise.write_text((nsrc[:a] + PACKED_HOST_PRIMITIVES + nsrc[b:]).replace("__device__", ""))
    ksrc = open(os.path.join(csrc, "tb_kernels.cuh")).read()
    body = ksrc[ksrc.index("struct PairEntry {"):ksrc.index("constexpr int kMaxBandRanks")].replace("__device__", "")
    cpp = d / "integrate_host.cpp"
    cpp.write_text(HARNESS % {"math": str(math), "noise": str(noise), "abi": os.path.join(ROOT, "include", "tendrils_b200.h"),
                              "body": body})
    out = d / "libintegrate_host.so"
    subprocess.run(["g++", "-O2", "-std=c++17", "-march=x86-64-v3", "-ffp-contract=off", "-fno-fast-math", "-fPIC", "-shared",
                    "-Wno-unknown-pragmas", "-Wno-unused-function", "-Wno-attributes", "-I/usr/local/cuda/include",
                    "-I", os.path.join(ROOT, "tests", "host_harness"), "-o", str(out), str(cpp)], check=True)
    L = C.CDLL(str(out))
    L.ih_integrate.restype = None
    L.ih_integrate.argtypes = [_fp, C.c_int, C.c_int, C.c_int, C.c_int, _fp, _fp, _fp, _fp, C.c_float, C.c_float,
                               C.c_int, C.c_int, C.c_int, C.c_int]
    return L


def hostile_state(rng, PW, PH):
    st = np.zeros((PW, PH, 4), np.float32)
    st[..., 0:2] = rng.uniform(-1.2, 1.2, (PW, PH, 2))
    st[..., 2:4] = rng.normal(0, 0.006, (PW, PH, 2))
    flat = st.reshape(-1, 4)
    n = flat.shape[0]
    pick = lambda k: rng.choice(n, k, replace=False)
    flat[pick(n // 10)] = (-1e6, -1e6, 0.003, -0.001)                       # inert (velocity kept)
    flat[pick(n // 20), 2:4] = 0.0                                          # motionless: 0/0 (PARITY I2)
    for v in (np.nan, np.inf, -np.inf, 1e7, -3e6, 1e30, 2.5e6, 1.9e6, 1e36, -3e38):   # wild positions: the shortcuts must not apply
        flat[pick(max(n // 60, 1)), rng.integers(0, 2)] = v
    flat[pick(n // 40), 2] = np.nan
    flat[pick(n // 40), 3] = np.inf
    flat[pick(n // 30), 0] = -1e6                                           # only ONE coordinate at the sentinel: not inert
    return st


def is_tame(v, lim):
    return bool(np.isfinite(v) and abs(v) < lim)


# (overrides of the reference defaults, hostile targets?, time, dt)
CASES = [
    ({}, False, 7 * 1000 / 60, 1000 / 60),
    ({"target": 0.002, "varyTarget": 1.5}, False, 7 * 1000 / 60, 1000 / 60),
    ({"target": 0.002, "varyTarget": 1.5, "noiseWeight": 0.0}, True, 7 * 1000 / 60, 1000 / 60),
    ({"noiseWeight": 0.0}, False, 7 * 1000 / 60, 1000 / 60),
    ({"noiseWeight": 0.0, "target": 0.0}, False, 123456.7, 1000 / 60),
    ({"noiseWeight": 0.0, "target": 0.0}, True, 50.0, 16.0),                # non-finite targets: the target shortcut is off
    ({"target": 0.0, "flowWeight": 0.0, "damping": 0.0}, False, 50.0, 16.0),
    ({"noiseWeight": 0.0, "varyNoise": 1e7}, False, 50.0, 16.0),            # wild variance: the noise shortcut is off
    ({"noiseWeight": 0.0, "noiseScale": float("inf")}, False, 50.0, 16.0),
    ({"speedLimit": 0.0}, False, 50.0, 16.0),
    ({"forceWeight": float("nan")}, False, 50.0, 16.0),
    ({"noiseScale": 1e6, "noiseSpeed": 10.0}, False, 9.0e5, 16.0),          # lattice coordinates beyond the packed domain
]


@pytest.mark.parametrize("shape", [(16, 32), (12, 300), (5, 7)])           # pow2 / not / tiny (division vs reciprocal)
@pytest.mark.parametrize("case", range(len(CASES)))
def test_integrate_kernel_on_host_equals_oracle(ih, oracle, shape, case):
    over, hostile_targets, time, dt = CASES[case]
    PW, PH = shape
    W, H = 24, 16
    rng = np.random.default_rng(100 * case + PW)
    O = oracle
    P = O.make_params(viewSize=(1.0, W / H), **over)
    st = hostile_state(rng, PW, PH)
    targets = np.zeros((PW, PH, 4), np.float32)
    targets[..., 0:2] = rng.uniform(-0.5, 0.5, (PW, PH, 2))
    if hostile_targets:
        targets.reshape(-1, 4)[rng.choice(PW * PH, 3, replace=False), 0] = (np.nan, np.inf, 1e30)
    flow = rng.normal(0, 0.01, (H, W, 4)).astype(np.float32)
    flow[..., 2] = rng.uniform(0, time, (H, W))
    flow[0, 0] = 0
    flow[1, 2, 0] = np.nan
    flow[2, 3, 2] = np.inf
    with np.errstate(all="ignore"):
        want = O.integrate(P, st, targets, flow, np.float32(time), np.float32(dt))
    S = np.array([getattr(P, n) for n, _ in P._fields_[:16]] + [P.viewSize[0], P.viewSize[1]], np.float32)
    # the shortcut flags, as tb_step sets them (the conditions are checked against the source text below)
    targets_finite = bool(np.isfinite(targets).all())
    use_targets = not (S[14] == 0.0 and is_tame(S[15], 1e6) and targets_finite)
    use_noise = not (S[6] == 0.0 and is_tame(S[7], 1e6) and is_tame(S[10], 1e6) and is_tame(S[11], 1e6) and
                     is_tame(S[12], 1e6) and is_tame(S[13], 1e6) and is_tame(np.float32(time), 1e9) and is_tame(np.float32(dt), 1e6))
    p = lambda a: a.ctypes.data_as(_fp)
    for packed in (1, 0):
        for split in (0, 1):
            for flags in ((use_targets, use_noise), (True, True)):
                got = np.full_like(st, 777.0)
                ih.ih_integrate(p(S), PW, PH, W, H, p(st), p(got), p(targets), p(flow), np.float32(time), np.float32(dt),
                                int(flags[0]), int(flags[1]), packed, split)
                same = (np.isnan(got) == np.isnan(want)).all() and np.array_equal(got[~np.isnan(got)], want[~np.isnan(want)])
                if not same:
                    bad = np.argwhere(~((got == want) | (np.isnan(got) & np.isnan(want))))[0]
                    raise AssertionError((case, shape, packed, split, flags, bad, st[bad[0], bad[1]], got[bad[0], bad[1]],
                                          want[bad[0], bad[1]]))


def test_harness_flags_follow_tb_step():
    """The two shortcut conditions above are copies of tb_step's: fail if the product's change."""
    api = open(os.path.join(ROOT, "tendrils_b200", "csrc", "tb_api.cu")).read()
    assert "A.use_targets = !(S.target == 0.0f && tame(S.varyTarget, 1e6f) && c->targets_finite);" in api
    assert ("A.use_noise = !(S.noiseWeight == 0.0f && tame(S.varyNoise, 1e6f) && tame(S.noiseScale, 1e6f) &&\n"
            "                    tame(S.varyNoiseScale, 1e6f) && tame(S.noiseSpeed, 1e6f) && tame(S.varyNoiseSpeed, 1e6f) &&\n"
            "                    tame(time, 1e9f) && tame(dt, 1e6f));") in api
    assert "bool tame(float v, float lim) { return std::isfinite(v) && std::fabs(v) < lim; }" in api


def test_harness_is_sensitive(ih, oracle):
    """Not a vacuous comparison: taking a shortcut where it does not apply changes the result."""
    PW, PH, W, H = 16, 32, 24, 16
    rng = np.random.default_rng(0)
    O = oracle
    P = O.make_params(viewSize=(1.0, W / H), target=0.002, varyTarget=1.5)  # noise on, target pull on
    st = hostile_state(rng, PW, PH)
    targets = np.zeros((PW, PH, 4), np.float32)
    targets[..., 0:2] = rng.uniform(-0.5, 0.5, (PW, PH, 2))
    flow = rng.normal(0, 0.01, (H, W, 4)).astype(np.float32)
    with np.errstate(all="ignore"):
        want = O.integrate(P, st, targets, flow, np.float32(100.0), np.float32(16.0))
    S = np.array([getattr(P, n) for n, _ in P._fields_[:16]] + [P.viewSize[0], P.viewSize[1]], np.float32)
    p = lambda a: a.ctypes.data_as(_fp)
    differing = {}
    for flags in ((1, 1), (1, 0), (0, 1)):
        got = np.zeros_like(st)
        ih.ih_integrate(p(S), PW, PH, W, H, p(st), p(got), p(targets), p(flow), np.float32(100.0), np.float32(16.0),
                        flags[0], flags[1], 1, 0)
        differing[flags] = int((~((got == want) | (np.isnan(got) & np.isnan(want)))).sum())
    assert differing[(1, 1)] == 0 and differing[(1, 0)] > 100 and differing[(0, 1)] > 100, differing


STRESS = int(os.environ.get("TB_STRESS", "0"))        # TB_STRESS=n: n extra random configurations (off in the normal run)


@pytest.mark.parametrize("seed", range(STRESS))
def test_stress_random_parameters(ih, oracle, seed):
    rng = np.random.default_rng(40_000 + seed)
    names = ["damping", "speedLimit", "forceWeight", "varyForce", "flowWeight", "varyFlow", "noiseWeight", "varyNoise", "flowDecay",
             "noiseScale", "varyNoiseScale", "noiseSpeed", "varyNoiseSpeed", "target", "varyTarget"]
    over = {}
    for n in names:
        r = rng.random()
        if r < 0.25:
            over[n] = 0.0
        elif r < 0.5:
            over[n] = float(rng.choice([1e-6, 0.01, 1.0, 50.0, 1e5, 1e7, -1.0, np.inf, np.nan]))
    CASES.append((over, bool(rng.random() < 0.3), float(rng.choice([0.0, 16.7, 1e4, 2e9])), float(rng.choice([0.0, 16.7, 1e7]))))
    try:
        test_integrate_kernel_on_host_equals_oracle(ih, oracle, (int(rng.integers(1, 20)), int(rng.integers(1, 70))), len(CASES) - 1)
    finally:
        CASES.pop()
