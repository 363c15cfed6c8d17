"""The spawn kernels (ball, direct, the five sample variants), the optical-flow kernel and the layer blend, cut out of
tendrils_b200/csrc/tb_kernels.cuh unchanged, compiled for the CPU by this test and compared bit for bit with the
oracle over parameter sweeps wider than the GPU parity tests drive (rotated / flipped spawn matrices, large times,
extreme jitter and bias, images and states holding negative, huge and non-finite values).  Nothing here is used by
the product."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from test_integrate_host import hostile_state
from util import synthetic_image

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_fp = C.POINTER(C.c_float)

HARNESS = r'''
#include <cstddef>
#include <cuda_runtime.h>
#include "cuda_intrinsics_shim.h"
#include "%(math)s"
#include "%(abi)s"
struct HostIdx { unsigned x, y, z; };
static HostIdx tb_host_blockIdx, tb_host_threadIdx, tb_host_blockDim;
#define blockIdx tb_host_blockIdx
#define threadIdx tb_host_threadIdx
#define blockDim tb_host_blockDim
#define __global__
#define __launch_bounds__(...)
template <class T> static inline T __ldg(const T *p) { return *p; }
namespace tb {
static constexpr float kInert = -1000000.0f;
%(kernels)s
using namespace tb;

template <class K, class A> static void launch(K kernel, const A &args, long long threads) {
    tb_host_blockDim = {256, 1, 1};
    for (long long b = 0; b < (threads + 255) / 256; ++b)
        for (unsigned t = 0; t < 256; ++t) {
            tb_host_blockIdx = {(unsigned)b, 0, 0}; tb_host_threadIdx = {t, 0, 0};
            kernel(args);
        }
}

extern "C" void ph_ball(int PW, int PH, float radius, float speed, float *out) {
    SpawnArgs A{}; A.out = (float4 *)out; A.PW = PW; A.PH = PH; A.p0 = 0; A.n = (long long)PW * PH; A.radius = radius; A.speed = speed;
    launch(k_spawn_ball, A, A.n);
}
// spawner13 = spawnSize[2], jitter[2], speed, bias, spawnMatrix[9] (tb_pixel_spawner)
extern "C" void ph_pixels(const float *spawner15, int apply, int vignette, int samples, int PW, int PH, const float *state,
                          const float *image, int IW, int IH, int xmajor, float time, float flowDecay, float *out) {
    SpawnArgs A{};
    std::memcpy(&A.U, spawner15, sizeof(tb_pixel_spawner));
    A.out = (float4 *)out; A.state = (const float4 *)state; A.image = (const float4 *)image;
    A.PW = PW; A.PH = PH; A.IW = IW; A.IH = IH; A.image_xmajor = xmajor; A.p0 = 0; A.n = (long long)PW * PH;
    A.time = time; A.flowDecay = flowDecay; A.apply = apply; A.vignette = vignette; A.samples = samples;
    if (samples == 0) launch(k_spawn_direct, A, A.n); else launch(k_spawn_sample, A, A.n);
}
extern "C" void ph_optical(const float *params9, const unsigned char *view, const unsigned char *last, int IW, int IH, float *flow, int W, int H) {
    OpticalArgs A{};
    std::memcpy(&A.U, params9, sizeof(tb_optical_flow_params));
    A.flow = (float4 *)flow; A.view = (const uchar4 *)view; A.last = (const uchar4 *)last; A.W = W; A.H = H; A.IW = IW; A.IH = IH;
    launch(k_optical_flow, A, (long long)W * H);
}
extern "C" void ph_blend(float *flow, const float *layer, int G) {
    tb_host_blockDim = {256, 1, 1};
    for (int b = 0; b < (G + 255) / 256; ++b)
        for (unsigned t = 0; t < 256; ++t) { tb_host_blockIdx = {(unsigned)b, 0, 0}; tb_host_threadIdx = {t, 0, 0}; k_blend_layer((float4 *)flow, (const float4 *)layer, G); }
}
'''


@pytest.fixture(scope="module")
def ph(tmp_path_factory):
    d = tmp_path_factory.mktemp("ph")
    csrc = os.path.join(ROOT, "tendrils_b200", "csrc")
    math = d / "tb_math_host.cuh"
    math.write_text(open(os.path.join(csrc, "tb_math.cuh")).read().replace("__device__", ""))
    ksrc = open(os.path.join(csrc, "tb_kernels.cuh")).read()
    kernels = ksrc[ksrc.index("// Full-grid alpha-over of an RGBA layer"):].replace("__device__", "")
    assert kernels.rstrip().endswith("}  // namespace tb")                   # the cut runs to the end of the namespace
    asrc = open(os.path.join(csrc, "tb_api.cu")).read()
    for line in ("case TB_SPAWN_DIRECT:        A.apply = APPLY_COLOR;     A.vignette = 1; A.samples = 0; break;",
                 "case TB_SPAWN_BEST_SAMPLE:   A.apply = APPLY_COLOR;     A.vignette = 1; A.samples = 6; break;",
                 "case TB_SPAWN_BRIGHT_SAMPLE: A.apply = APPLY_BRIGHTEST; A.vignette = 0; A.samples = 6; break;",
                 "case TB_SPAWN_COLOR_SAMPLE:  A.apply = APPLY_COLOR;     A.vignette = 0; A.samples = 3; break;",
                 "case TB_SPAWN_DATA_SAMPLE:   A.apply = APPLY_IDENTITY;  A.vignette = 1; A.samples = 2; break;",
                 "case TB_SPAWN_FLOW_SAMPLE:   A.apply = APPLY_FLOW;      A.vignette = 0; A.samples = 5; break;"):
        assert line in asrc                                                   # VARIANTS below mirror tb_spawn_pixels
    cpp = d / "spawn_host.cpp"
    cpp.write_text(HARNESS % {"math": str(math), "abi": os.path.join(ROOT, "include", "tendrils_b200.h"), "kernels": kernels})
    out = d / "libspawn_host.so"
    subprocess.run(["g++", "-O2", "-std=c++17", "-march=x86-64-v3", "-ffp-contract=off", "-fno-fast-math", "-fPIC", "-shared",
                    "-Wno-unknown-pragmas", "-Wno-unused-function", "-Wno-attributes", "-I/usr/local/cuda/include",
                    "-I", os.path.join(ROOT, "tests", "host_harness"), "-o", str(out), str(cpp)], check=True)
    L = C.CDLL(str(out))
    L.ph_ball.argtypes = [C.c_int, C.c_int, C.c_float, C.c_float, _fp]
    L.ph_pixels.argtypes = [_fp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _fp, _fp, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, _fp]
    L.ph_optical.argtypes = [_fp, C.c_void_p, C.c_void_p, C.c_int, C.c_int, _fp, C.c_int, C.c_int]
    L.ph_blend.argtypes = [_fp, _fp, C.c_int]
    for f in (L.ph_ball, L.ph_pixels, L.ph_optical, L.ph_blend):
        f.restype = None
    return L


def same(a, b):
    return (np.isnan(a) == np.isnan(b)).all() and np.array_equal(a[~np.isnan(a)], b[~np.isnan(b)])


p = lambda a: a.ctypes.data_as(_fp)


@pytest.mark.parametrize("PW,PH,radius,speed", [(16, 16, 0.3, 0.005), (7, 300, 1.0, 0.0), (257, 3, 1e4, -2.0), (1, 1, 0.0, 1.0)])
def test_ball_spawn(ph, oracle, PW, PH, radius, speed):
    got = np.zeros((PW, PH, 4), np.float32)
    ph.ph_ball(PW, PH, radius, speed, p(got))
    assert same(got, oracle.spawn_ball(PW, PH, radius, speed))


# name: (apply, vignette, samples) -- tb_spawn_pixels; 0 samples = the direct shader
VARIANTS = {"direct": (0, 1, 0), "best": (0, 1, 6), "bright": (1, 0, 6), "color": (0, 0, 3), "data": (2, 1, 2), "flow": (3, 0, 5)}
SPAWNERS = [  # spawnSize, jitter, speed, bias, spawnMatrix (column-major mat3), time
    ((1.0, 1.0), (0.002, 0.002), 1.0, 1.0, (-1, 0, 0, 0, 1, 0, 0, 0, 1), 16.7),
    ((0.9, 1.3), (0.05, 0.0), 0.3, 0.2, (0.6, 0.8, 0, -0.8, 0.6, 0, 0.1, -0.2, 1), 9.87e5),
    ((2.0, -0.5), (0.0, 0.5), -2.0, 5.0, (0, 1, 0, 1, 0, 0, 0, 0, 1), 0.0),
    ((1.0, 1.0), (1e3, 1e-9), 1e-3, 0.0, (1, 0, 0, 0, 1, 0, 0, 0, 1), 3.3e6),
]


@pytest.mark.parametrize("variant", sorted(VARIANTS))
@pytest.mark.parametrize("sp", range(len(SPAWNERS)))
def test_pixel_spawners(ph, oracle, variant, sp):
    O = oracle
    size, jitter, speed, bias, mat, time = SPAWNERS[sp]
    rng = np.random.default_rng(31 * sp + len(variant))
    PW, PH = (12, 20) if sp % 2 else (16, 16)
    state = hostile_state(rng, PW, PH)
    if variant == "flow":                                                    # spawnData = the flow grid
        image = rng.normal(0, 0.01, (9, 14, 4)).astype(np.float32)
        image[..., 2] = rng.uniform(0, 2 * max(time, 1.0), (9, 14))
        image[0, 0] = 0
        image[3, 5, 0] = np.nan
    elif variant == "data":                                                  # spawnData = the particle texture, x-major
        image = state
    else:
        image = synthetic_image(11, 7)
        image[2, 3] = (-0.5, 2.0, 1e9, 1.0)
        image[4, 1] = (np.nan, 0.5, 0.5, np.inf)
        image[5, 5] = 0
    S = O.make_spawn_pixels(spawnSize=size, jitter=jitter, speed=speed, bias=bias, spawnMatrix=mat, flowDecay=0.005)
    apply, vig, samples = VARIANTS[variant]
    with np.errstate(all="ignore"):
        if variant == "direct":
            want = O.spawn_pixels_direct(S, PW, PH, image, np.float32(time))
        elif variant == "data":
            want = O.spawn_pixels_sample(S, variant, state, np.ascontiguousarray(image.transpose(1, 0, 2)), np.float32(time))
        else:
            want = O.spawn_pixels_sample(S, variant, state, image, np.float32(time))
    U = np.array(list(size) + list(jitter) + [speed, bias] + list(mat), np.float32)
    got = np.zeros((PW, PH, 4), np.float32)
    xmajor = 1 if variant == "data" else 0
    IW, IH = (PW, PH) if xmajor else (image.shape[1], image.shape[0])
    img = np.ascontiguousarray(image, np.float32)
    ph.ph_pixels(p(U), apply, vig, samples, PW, PH, p(state), p(img), IW, IH, xmajor, np.float32(time), np.float32(0.005), p(got))
    assert same(got, want), (variant, sp)


@pytest.mark.parametrize("case", range(4))
def test_optical_flow_and_layer_blend(ph, oracle, case):
    O = oracle
    rng = np.random.default_rng(900 + case)
    (W, H), (iw, ih) = [((24, 16), (30, 18)), ((7, 33), (5, 5)), ((64, 64), (64, 64)), ((1, 1), (3, 2))][case]
    params = [dict(viewSize=(1.0, 1.5), scaleUV=(-1, -1), offset=0.1, lambda_=0.001, speed=0.08, speedLimit=0.01, time=150.0),
              dict(viewSize=(2.0, 1.0), scaleUV=(1, -1), offset=1.0, lambda_=0.0, speed=1.0, speedLimit=1.0, time=1.0),
              dict(viewSize=(1.0, 1.0), scaleUV=(0.5, 3.0), offset=0.013, lambda_=1e-6, speed=-4.0, speedLimit=1e-4, time=9.9e5),
              dict(viewSize=(1.0, 1.0), scaleUV=(-1, -1), offset=0.0, lambda_=0.001, speed=0.08, speedLimit=0.0, time=5.0)][case]
    view = rng.integers(0, 256, (ih, iw, 4), dtype=np.uint8)
    last = np.clip(view.astype(np.int32) + rng.integers(-60, 61, view.shape), 0, 255).astype(np.uint8)
    if case == 1:
        last = view.copy()                                                   # no motion: 0/sqrt(0 + lambda = 0) = NaN
    base = rng.normal(0, 0.01, (H, W, 4)).astype(np.float32)
    want = base.copy()
    with np.errstate(all="ignore"):
        O.optical_flow(want, view, last, **params)
    got = base.copy()
    U = np.array(list(params["viewSize"]) + list(params["scaleUV"]) + [params["offset"], params["lambda_"], params["speed"],
                                                                         params["speedLimit"], params["time"]], np.float32)
    ph.ph_optical(p(U), view.ctypes.data, last.ctypes.data, iw, ih, p(got), W, H)
    assert same(got, want)
    # the layer blend is spec/PARITY.md B2 on every texel
    layer = rng.normal(0, 0.5, (H, W, 4)).astype(np.float32)
    layer[..., 3] = rng.uniform(-0.2, 1.2, (H, W))
    a, om = layer[..., 3:4], np.float32(1.0) - layer[..., 3:4]
    want2 = (layer * a + got * om).astype(np.float32)
    ph.ph_blend(p(got), p(np.ascontiguousarray(layer)), W * H)
    assert same(got, want2)


STRESS = int(os.environ.get("TB_STRESS", "0"))        # TB_STRESS=n: n extra random configurations (off in the normal run)


@pytest.mark.parametrize("seed", range(STRESS))
def test_stress_random_spawners(ph, oracle, seed):
    rng = np.random.default_rng(60_000 + seed)
    pick = lambda xs: float(rng.choice(xs))
    mat = tuple(float(v) for v in rng.choice([0, 1, -1, 0.5, 2.5, 1e3], 9))
    SPAWNERS.append(((pick([1, 0.3, -2, 0]), pick([1, 0.3, -2, 1e4])), (pick([0, 0.002, 0.5, 1e3]), pick([0, 0.002, 0.5])),
                     pick([1, 0.3, -2, 0, 1e4]), pick([1, 0, 0.2, 5, -1, np.inf]), mat, pick([0, 16.7, 9.87e5, 3.3e6, 1e9])))
    try:
        test_pixel_spawners(ph, oracle, sorted(VARIANTS)[seed % len(VARIANTS)], len(SPAWNERS) - 1)
    finally:
        SPAWNERS.pop()
