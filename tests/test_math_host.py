"""The scalar device arithmetic of tendrils_b200/csrc/tb_math.cuh (TSIN-1 sin/cos, the glsl-random hash, the scalar
simplex noise) compiled for the CPU by this test and compared bit for bit with the oracle's functions on millions of
inputs, including the corners a simulation run never visits.  Nothing here is used by the product."""
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
extern "C" { float or_sin(float); float or_cos(float); float or_random(float, float); float or_snoise3(float, float, float); }
static inline bool same(float a, float b) {
    unsigned x, y; std::memcpy(&x, &a, 4); std::memcpy(&y, &b, 4);
    return x == y || (a != a && b != b);                     // any NaN equals any NaN (spec/PARITY.md)
}
extern "C" long long mh_sincos(long long n, const float *x, long long *first_bad) {
    long long bad = 0; *first_bad = -1;
    for (long long i = 0; i < n; ++i) {
        float s, c; tb::sincos_t1(x[i], s, c);
        if (!same(s, or_sin(x[i])) || !same(c, or_cos(x[i]))) { if (*first_bad < 0) *first_bad = i; ++bad; }
    }
    return bad;
}
extern "C" long long mh_random(long long n, const float *xy, long long *first_bad) {
    long long bad = 0; *first_bad = -1;
    for (long long i = 0; i < n; ++i)
        if (!same(tb::grandom(xy[2 * i], xy[2 * i + 1]), or_random(xy[2 * i], xy[2 * i + 1]))) { if (*first_bad < 0) *first_bad = i; ++bad; }
    return bad;
}
extern "C" long long mh_snoise(long long n, const float *xyz, long long *first_bad) {
    long long bad = 0; *first_bad = -1;
    for (long long i = 0; i < n; ++i)
        if (!same(tb::snoise3(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]), or_snoise3(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]))) {
            if (*first_bad < 0) *first_bad = i; ++bad;
        }
    return bad;
}
'''


@pytest.fixture(scope="module")
def mh(tmp_path_factory, oracle):
    d = tmp_path_factory.mktemp("mh")
    math = d / "tb_math_host.cuh"
    math.write_text(open(os.path.join(ROOT, "tendrils_b200", "csrc", "tb_math.cuh")).read().replace("__device__", ""))
    cpp = d / "math_host.cpp"
    cpp.write_text(HARNESS % {"math": str(math)})
    out = d / "libmath_host.so"
    odir = os.path.join(ROOT, "oracle", "_build")
    subprocess.run(["g++", "-O2", "-std=c++17", "-march=x86-64-v3", "-ffp-contract=off", "-fno-fast-math", "-fPIC", "-shared",
                    "-Wno-unknown-pragmas", "-Wno-unused-function", "-I/usr/local/cuda/include",
                    "-I", os.path.join(ROOT, "tests", "host_harness"), "-o", str(out), str(cpp),
                    "-L", odir, "-ltendrils_oracle", f"-Wl,-rpath,{odir}"], check=True)
    L = C.CDLL(str(out))
    for f in (L.mh_sincos, L.mh_random, L.mh_snoise):
        f.restype = C.c_longlong
        f.argtypes = [C.c_longlong, C.POINTER(C.c_float), C.POINTER(C.c_longlong)]
    return L


def run(fn, a, per):
    a = np.ascontiguousarray(a, np.float32)
    first = C.c_longlong()
    bad = fn(a.size // per, a.ctypes.data_as(C.POINTER(C.c_float)), C.byref(first))
    assert bad == 0, (bad, first.value, a.reshape(-1, per)[first.value] if first.value >= 0 else None)


def test_sin_cos_whole_domain(mh):
    rng = np.random.default_rng(1)
    k = np.arange(-70000, 70000, dtype=np.float64) * (np.pi / 2)              # quadrant boundaries and their neighbours
    near = np.concatenate([np.nextafter(k.astype(np.float32), np.float32(s)) for s in (-np.inf, np.inf)] + [k.astype(np.float32)])
    x = np.concatenate([rng.uniform(-1e5, 1e5, 2_000_000), rng.normal(0, 4, 500_000), near,
                        [0.0, -0.0, 1e5, -1e5, 100000.01, np.nan, np.inf, -np.inf, 1e-30, -1e-30, 3.14, 12582912.0]])
    run(mh.mh_sincos, x, 1)


def test_random_hash(mh):
    rng = np.random.default_rng(2)
    a = np.concatenate([rng.uniform(-2, 2, (1_000_000, 2)), rng.uniform(0, 4097, (1_000_000, 2)),        # uv seeds, gl_FragCoord
                        rng.uniform(-1e6, 1e6, (300_000, 2)), np.array([[0, 0], [np.nan, 1], [1e30, 1e30], [-1e6 + 0.3, -1e6 + 0.7]])])
    run(mh.mh_random, a, 2)


def test_scalar_simplex_noise(mh):
    rng = np.random.default_rng(3)
    lattice = rng.integers(-50, 50, (200_000, 3)).astype(np.float64) + rng.choice([0.0, 1 / 3, 1 / 6, 0.5], (200_000, 3))
    a = np.concatenate([rng.uniform(-4, 4, (1_500_000, 3)), rng.uniform(-300, 300, (500_000, 3)), lattice,
                        rng.uniform(-1e5, 1e5, (100_000, 3)), np.array([[0, 0, 0], [np.nan, 0, 0], [1e30, 0, 0], [-0.0, 0.0, -0.0]])])
    run(mh.mh_snoise, a, 3)


# ---- the packed (FFMA2) twin noise of tb_noise2.cuh ----------------------------------------------------------------
# Its three inline-PTX primitives are swapped for host equivalents -- mov.b64 pack/unpack as bit packing, and
# fma.rn.f32x2 as one correctly rounded fmaf per lane, which is what the instruction is -- and everything built on
# them (the exact-integer FMAs, the dropped floors, the sign / select rewrites) runs unchanged against the oracle.
PACKED_HOST_PRIMITIVES = r'''
inline F2 pack2(float lo, float hi) {
    F2 r; unsigned a, b; std::memcpy(&a, &lo, 4); std::memcpy(&b, &hi, 4);
    r.v = (unsigned long long)a | ((unsigned long long)b << 32);
    return r;
}
inline F2 splat2(float x) { return pack2(x, x); }
inline void unpack2(F2 a, float &lo, float &hi) {
    const unsigned x = (unsigned)(a.v & 0xffffffffull), y = (unsigned)(a.v >> 32);
    std::memcpy(&lo, &x, 4); std::memcpy(&hi, &y, 4);
}
inline F2 fma2(F2 a, F2 b, F2 c) {
    float a0, a1, b0, b1, c0, c1;
    unpack2(a, a0, a1); unpack2(b, b0, b1); unpack2(c, c0, c1);
    return pack2(fmaf(a0, b0, c0), fmaf(a1, b1, c1));
}
'''

PACKED_HARNESS = r'''
#include <cuda_runtime.h>
#include "cuda_intrinsics_shim.h"
#include "%(noise)s"
extern "C" float or_snoise3(float, float, float);
static inline bool same(float a, float b) {
    unsigned x, y; std::memcpy(&x, &a, 4); std::memcpy(&y, &b, 4);
    return x == y || (a != a && b != b);
}
extern "C" long long mh_snoise_pair(long long n, const float *v /* n x 4: x y za zb */, long long *first_bad) {
    const tb::PackedConsts k{1.0f, -1.0f, -0.0f};
    long long bad = 0; *first_bad = -1;
    for (long long i = 0; i < n; ++i) {
        float a, b;
        tb::snoise3_pair(k, v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3], a, b);
        if (!same(a, or_snoise3(v[4 * i], v[4 * i + 1], v[4 * i + 2])) || !same(b, or_snoise3(v[4 * i], v[4 * i + 1], v[4 * i + 3]))) {
            if (*first_bad < 0) *first_bad = i; ++bad;
        }
    }
    return bad;
}
'''


@pytest.fixture(scope="module")
def mh_packed(tmp_path_factory, oracle):
    import re
    d = tmp_path_factory.mktemp("mhp")
    src = open(os.path.join(ROOT, "tendrils_b200", "csrc", "tb_noise2.cuh")).read()
    a, b = src.index("__device__ __forceinline__ F2 pack2("), src.index("struct P2 {")
    assert src[a:b].count("asm(") == 3                                      # exactly the three primitives are swapped
    host = (src[:a] + PACKED_HOST_PRIMITIVES + src[b:]).replace("__device__", "")
    assert "asm(" not in host
    noise = d / "tb_noise2_host.cuh"
    noise.write_text(host)
    cpp = d / "noise_host.cpp"
    cpp.write_text(PACKED_HARNESS % {"noise": str(noise)})
    out = d / "libnoise_host.so"
    odir = os.path.join(ROOT, "oracle", "_build")
    subprocess.run(["g++", "-O2", "-std=c++17", "-march=x86-64-v3", "-ffp-contract=off", "-fno-fast-math", "-fPIC", "-shared",
                    "-Wno-unknown-pragmas", "-Wno-unused-function", "-I/usr/local/cuda/include",
                    "-I", os.path.join(ROOT, "tests", "host_harness"), "-o", str(out), str(cpp),
                    "-L", odir, "-ltendrils_oracle", f"-Wl,-rpath,{odir}"], check=True)
    L = C.CDLL(str(out))
    L.mh_snoise_pair.restype = C.c_longlong
    L.mh_snoise_pair.argtypes = [C.c_longlong, C.POINTER(C.c_float), C.POINTER(C.c_longlong)]
    return L


def test_packed_twin_noise_equals_scalar_reference(mh_packed):
    rng = np.random.default_rng(4)
    lattice = rng.integers(-60, 60, (300_000, 4)).astype(np.float64) + rng.choice([0.0, 1 / 3, 1 / 6, 0.5, 1e-7], (300_000, 4))
    workload = np.concatenate([rng.uniform(-3.5, 3.5, (2_000_000, 2)),                       # pos * noiseScale
                               rng.uniform(0, 60, (2_000_000, 1)) + rng.uniform(0, 1, (2_000_000, 1)),   # uv.x + noiseTime
                               rng.uniform(1234, 1300, (2_000_000, 1))], 1)                   # uv.y + noiseTime + 1234.5678
    # k_integrate only takes the packed path when every coordinate is below 2e6 in magnitude (`lattice_ok`: the lattice
    # integers then stay exact in binary32); anything wilder, NaN included, goes to the scalar noise tested above
    lim = np.nextafter(np.float32(2.0e6), np.float32(0))
    edge = np.clip(rng.uniform(-2.0e6, 2.0e6, (400_000, 4)), -lim, lim)
    a = np.concatenate([workload, rng.uniform(-300, 300, (800_000, 4)), lattice, rng.uniform(-1e5, 1e5, (200_000, 4)), edge,
                        np.array([[0, 0, 0, 0], [-0.0, 0.0, -0.0, 0.0], [lim, -lim, lim, -lim], [-lim, lim, -lim, lim],
                                  [1e-38, -1e-38, 1e-45, 0]])])
    run(mh_packed.mh_snoise_pair, a, 4)
