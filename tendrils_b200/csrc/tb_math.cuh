// tb_math.cuh -- device arithmetic of the Tendrils step for sm_100a.
//
// Contract (spec/PARITY.md): every GLSL operator of the reference shaders is ONE IEEE-754
// binary32 operation, evaluated in source order; no FMA contraction (this translation unit
// is compiled with -fmad=false, and uses __f*_rn intrinsics where the order matters most).
// sin/cos follow the TSIN-1 recipe so that the hash RNG of the spawners is reproducible.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace tb {

// ---- GLSL ES 1.00 built-ins, NaN behaviour of the spec's defining formulas ---------------
__device__ __forceinline__ float gmin(float x, float y) { return (y < x) ? y : x; }
__device__ __forceinline__ float gmax(float x, float y) { return (x < y) ? y : x; }
__device__ __forceinline__ float gstep(float edge, float x) { return (x < edge) ? 0.0f : 1.0f; }
__device__ __forceinline__ float gfract(float x) { return __fsub_rn(x, floorf(x)); }
__device__ __forceinline__ float gmod(float x, float y) {
    return __fsub_rn(x, __fmul_rn(y, floorf(__fdiv_rn(x, y))));
}
__device__ __forceinline__ float gmix(float x, float y, float a) {
    return __fadd_rn(__fmul_rn(x, __fsub_rn(1.0f, a)), __fmul_rn(y, a));
}
__device__ __forceinline__ float glength(float x, float y) {
    return __fsqrt_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)));
}
__device__ __forceinline__ bool is_finite(float x) { return fabsf(x) <= 3.402823466e+38f; }

// ---- TSIN-1: sin and cos by Cody-Waite reduction (pi/2 in three parts) + Cephes polynomials
__device__ __forceinline__ void sincos_t1(float x, float &s_out, float &c_out) {
    if (!(fabsf(x) <= 100000.0f)) {
        s_out = __int_as_float(0x7fc00000);
        c_out = s_out;
        return;
    }
    float kf = __fmul_rn(x, 0.636619772f);
    kf = __fsub_rn(__fadd_rn(kf, 12582912.0f), 12582912.0f);
    float r = __fsub_rn(x, __fmul_rn(kf, 1.5703125f));
    r = __fsub_rn(r, __fmul_rn(kf, 4.837512969970703125e-4f));
    r = __fsub_rn(r, __fmul_rn(kf, 7.54978995489188216e-8f));
    const int q = static_cast<int>(kf) & 3;
    const float z = __fmul_rn(r, r);
    float ps = __fmul_rn(-1.9515295891e-4f, z);
    ps = __fadd_rn(ps, 8.3321608736e-3f);
    ps = __fmul_rn(ps, z);
    ps = __fsub_rn(ps, 1.6666654611e-1f);
    ps = __fmul_rn(ps, z);
    ps = __fmul_rn(ps, r);
    const float s = __fadd_rn(ps, r);
    float pc = __fmul_rn(2.443315711809948e-5f, z);
    pc = __fsub_rn(pc, 1.388731625493765e-3f);
    pc = __fmul_rn(pc, z);
    pc = __fadd_rn(pc, 4.166664568298827e-2f);
    pc = __fmul_rn(pc, z);
    pc = __fmul_rn(pc, z);
    pc = __fsub_rn(pc, __fmul_rn(0.5f, z));
    const float c = __fadd_rn(pc, 1.0f);
    const float sv = (q & 1) ? c : s;
    const float cv = (q & 1) ? s : c;
    s_out = (q & 2) ? -sv : sv;
    c_out = ((q + 1) & 2) ? -cv : cv;
}

// glsl-random@0.0.5 -- spawn/ball/index.frag:12-15, spawn/pixels/frag/head.frag:30-31
__device__ __forceinline__ float grandom(float cx, float cy) {
    const float dt = __fadd_rn(__fmul_rn(cx, 12.9898f), __fmul_rn(cy, 78.233f));
    const float sn = gmod(dt, 3.14f);
    float s, c;
    sincos_t1(sn, s, c);
    return gfract(__fmul_rn(s, 43758.5453f));
}

// ---- glsl-noise simplex/3d (docs/js/index.js.map sourcesContent[79]; logic.frag:67-68) ----
__device__ __forceinline__ float mod289(float x) {
    return __fsub_rn(x, __fmul_rn(floorf(__fmul_rn(x, 1.0f / 289.0f)), 289.0f));
}
__device__ __forceinline__ float permute(float x) {
    return mod289(__fmul_rn(__fadd_rn(__fmul_rn(x, 34.0f), 1.0f), x));
}
__device__ __forceinline__ float dot3(float ax, float ay, float az, float bx, float by, float bz) {
    return __fadd_rn(__fadd_rn(__fmul_rn(ax, bx), __fmul_rn(ay, by)), __fmul_rn(az, bz));
}

// One simplex corner: gradient from the permuted hash p, falloff and projection onto (x,y,z).
__device__ __forceinline__ float snoise_corner(float p, float x, float y, float z) {
    const float n7 = 0.142857142857f;          // ns.z
    const float nsx = __fmul_rn(n7, 2.0f);     // ns.x = n_*D.w - D.x
    const float nsy = __fsub_rn(__fmul_rn(n7, 0.5f), 1.0f);
    const float j = __fsub_rn(p, __fmul_rn(49.0f, floorf(__fmul_rn(__fmul_rn(p, n7), n7))));
    const float xq = floorf(__fmul_rn(j, n7));
    const float yq = floorf(__fsub_rn(j, __fmul_rn(7.0f, xq)));
    const float gx0 = __fadd_rn(__fmul_rn(xq, nsx), nsy);
    const float gy0 = __fadd_rn(__fmul_rn(yq, nsx), nsy);
    const float h = __fsub_rn(__fsub_rn(1.0f, fabsf(gx0)), fabsf(gy0));
    const float sx = __fadd_rn(__fmul_rn(floorf(gx0), 2.0f), 1.0f);
    const float sy = __fadd_rn(__fmul_rn(floorf(gy0), 2.0f), 1.0f);
    const float sh = -gstep(h, 0.0f);
    float gx = __fadd_rn(gx0, __fmul_rn(sx, sh));
    float gy = __fadd_rn(gy0, __fmul_rn(sy, sh));
    float gz = h;
    const float nrm = __fsub_rn(1.79284291400159f, __fmul_rn(0.85373472095314f, dot3(gx, gy, gz, gx, gy, gz)));
    gx = __fmul_rn(gx, nrm);
    gy = __fmul_rn(gy, nrm);
    gz = __fmul_rn(gz, nrm);
    float m = gmax(__fsub_rn(0.6f, dot3(x, y, z, x, y, z)), 0.0f);
    m = __fmul_rn(m, m);
    m = __fmul_rn(m, m);
    return __fmul_rn(m, dot3(gx, gy, gz, x, y, z));
}

__device__ __forceinline__ float snoise3(float vx, float vy, float vz) {
    const float Cx = 1.0f / 6.0f, Cy = 1.0f / 3.0f;
    const float d = __fadd_rn(__fadd_rn(__fmul_rn(vx, Cy), __fmul_rn(vy, Cy)), __fmul_rn(vz, Cy));
    float ix = floorf(__fadd_rn(vx, d)), iy = floorf(__fadd_rn(vy, d)), iz = floorf(__fadd_rn(vz, d));
    const float e = __fadd_rn(__fadd_rn(__fmul_rn(ix, Cx), __fmul_rn(iy, Cx)), __fmul_rn(iz, Cx));
    const float x0 = __fadd_rn(__fsub_rn(vx, ix), e);
    const float y0 = __fadd_rn(__fsub_rn(vy, iy), e);
    const float z0 = __fadd_rn(__fsub_rn(vz, iz), e);

    const float gx = gstep(y0, x0), gy = gstep(z0, y0), gz = gstep(x0, z0);
    const float lx = __fsub_rn(1.0f, gx), ly = __fsub_rn(1.0f, gy), lz = __fsub_rn(1.0f, gz);
    const float i1x = gmin(gx, lz), i1y = gmin(gy, lx), i1z = gmin(gz, ly);
    const float i2x = gmax(gx, lz), i2y = gmax(gy, lx), i2z = gmax(gz, ly);

    ix = mod289(ix);
    iy = mod289(iy);
    iz = mod289(iz);

    float pa = permute(__fadd_rn(iz, 0.0f));
    float pb = permute(__fadd_rn(iz, i1z));
    float pc = permute(__fadd_rn(iz, i2z));
    float pd = permute(__fadd_rn(iz, 1.0f));
    pa = permute(__fadd_rn(__fadd_rn(pa, iy), 0.0f));
    pb = permute(__fadd_rn(__fadd_rn(pb, iy), i1y));
    pc = permute(__fadd_rn(__fadd_rn(pc, iy), i2y));
    pd = permute(__fadd_rn(__fadd_rn(pd, iy), 1.0f));
    pa = permute(__fadd_rn(__fadd_rn(pa, ix), 0.0f));
    pb = permute(__fadd_rn(__fadd_rn(pb, ix), i1x));
    pc = permute(__fadd_rn(__fadd_rn(pc, ix), i2x));
    pd = permute(__fadd_rn(__fadd_rn(pd, ix), 1.0f));

    const float t0 = snoise_corner(pa, x0, y0, z0);
    const float t1 = snoise_corner(pb, __fadd_rn(__fsub_rn(x0, i1x), Cx), __fadd_rn(__fsub_rn(y0, i1y), Cx),
                                   __fadd_rn(__fsub_rn(z0, i1z), Cx));
    const float t2 = snoise_corner(pc, __fadd_rn(__fsub_rn(x0, i2x), Cy), __fadd_rn(__fsub_rn(y0, i2y), Cy),
                                   __fadd_rn(__fsub_rn(z0, i2z), Cy));
    const float t3 = snoise_corner(pd, __fsub_rn(x0, 0.5f), __fsub_rn(y0, 0.5f), __fsub_rn(z0, 0.5f));
    return __fmul_rn(42.0f, __fadd_rn(__fadd_rn(__fadd_rn(t0, t1), t2), t3));
}

// NEAREST + CLAMP_TO_EDGE texel index; NaN selects texel 0 (spec/PARITY.md Q5).
__device__ __forceinline__ int texel_of(float u, int size) {
    const float f = floorf(__fmul_rn(u, static_cast<float>(size)));
    if (!(f > 0.0f)) return 0;
    if (f > static_cast<float>(size - 1)) return size - 1;
    return static_cast<int>(f);
}

__device__ __forceinline__ float vary(float base, float offset, float variance) {  // logic.frag:41-43
    return __fadd_rn(base, __fmul_rn(__fmul_rn(offset, variance), base));
}

}  // namespace tb
