// tb_noise2.cuh -- the two simplex-noise evaluations of logic.frag (src/logic.frag:67-68), done in
// lock step on Blackwell's packed FP32 pipe (FFMA2: two binary32 lanes per instruction).
//
// Lane 0 carries snoise(vec3(noisePos, uv.x + noiseTime)), lane 1 carries
// snoise(vec3(noisePos, uv.y + noiseTime + 1234.5678)).  The arithmetic contract is unchanged
// (spec/PARITY.md R1): every GLSL operator is one correctly rounded binary32 operation in source
// order.  Three things make that compatible with FFMA2:
//
//  * a*b is issued as fma(a, b, -0) and a+b as fma(a, 1, b): both are exact restatements (one
//    rounding, same value, same zero sign).  The -0 / 1 / -1 operands come from kernel arguments,
//    so ptxas can neither simplify them away nor contract a neighbouring mul+add pair into a
//    single-rounding FMA -- which it otherwise does for the packed forms even under explicit
//    .rn (observed with CUDA 12.9: mul.rn.f32x2 + add.rn.f32x2 -> FFMA2).
//  * a true FMA is used only where the product is exactly representable, so fusing cannot change
//    the result: permute's (x*34)+1 and x - floor(x/289)*289, j = p - 49*floor(..), j - 7*x_ --
//    all small integers (< 2^24) in binary32.
//  * floors whose argument is provably an integer are dropped (y_ = floor(j - 7*x_)), and
//    floor(b)*2+1 on b in (-1,1) is the sign of b; s*sh with sh in {-0,-1} is a select.  Each such
//    rewrite is value-identical for every finite input and NaN-in/NaN-out otherwise.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace tb {

struct F2 { unsigned long long v; };     // two binary32 lanes in one 64-bit register pair

struct PackedConsts {                    // opaque to the compiler: filled by the host with 1, -1, -0
    float one, neg_one, neg_zero;
};

__device__ __forceinline__ F2 pack2(float lo, float hi) {
    F2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r.v) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ F2 splat2(float x) { return pack2(x, x); }
__device__ __forceinline__ void unpack2(F2 a, float &lo, float &hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(a.v));
}
__device__ __forceinline__ F2 fma2(F2 a, F2 b, F2 c) {
    F2 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r.v) : "l"(a.v), "l"(b.v), "l"(c.v));
    return r;
}

struct P2 {                              // the packed operators, closed over the opaque constants
    F2 ONE, MONE, NZ;
    __device__ __forceinline__ explicit P2(const PackedConsts &k)
        : ONE(splat2(k.one)), MONE(splat2(k.neg_one)), NZ(splat2(k.neg_zero)) {}
    __device__ __forceinline__ F2 mul(F2 a, F2 b) const { return fma2(a, b, NZ); }      // a*b
    __device__ __forceinline__ F2 add(F2 a, F2 b) const { return fma2(a, ONE, b); }     // a+b
    __device__ __forceinline__ F2 sub(F2 a, F2 b) const { return fma2(b, MONE, a); }    // a-b
    __device__ __forceinline__ F2 mulc(F2 a, float c) const { return fma2(a, splat2(c), NZ); }
    __device__ __forceinline__ F2 addc(F2 a, float c) const { return fma2(a, ONE, splat2(c)); }
};

__device__ __forceinline__ F2 floor2(F2 a) {
    float lo, hi;
    unpack2(a, lo, hi);
    return pack2(floorf(lo), floorf(hi));
}

// mod289(x) = x - floor(x*(1/289))*289; the product floor(..)*289 is an exact integer
__device__ __forceinline__ F2 mod289_2(const P2 &p, F2 x) {
    const F2 q = floor2(p.mulc(x, 1.0f / 289.0f));
    return fma2(q, splat2(-289.0f), x);
}
// permute(x) = mod289(((x*34)+1)*x); x is an integer in [0, 579], so (x*34)+1 is exact
__device__ __forceinline__ F2 permute2(const P2 &p, F2 x) {
    const F2 t = fma2(x, splat2(34.0f), splat2(1.0f));
    return mod289_2(p, p.mul(t, x));
}

__device__ __forceinline__ F2 dot3_2(const P2 &p, F2 ax, F2 ay, F2 az, F2 bx, F2 by, F2 bz) {
    return p.add(p.add(p.mul(ax, bx), p.mul(ay, by)), p.mul(az, bz));
}

// one simplex corner for both lanes: m^4 * dot(gradient(perm), x)
__device__ __forceinline__ F2 corner2(const P2 &p, F2 perm, F2 x, F2 y, F2 z) {
    const float n7 = 0.142857142857f;                                  // ns.z
    const float nsx = __fmul_rn(n7, 2.0f);                             // ns.x
    const float nsy = __fsub_rn(__fmul_rn(n7, 0.5f), 1.0f);            // ns.y
    // j = p - 49*floor(p*ns.z*ns.z); x_ = floor(j*ns.z); y_ = floor(j - 7*x_) = j - 7*x_ (an integer)
    const F2 j = fma2(floor2(p.mulc(p.mulc(perm, n7), n7)), splat2(-49.0f), perm);
    const F2 xq = floor2(p.mulc(j, n7));
    const F2 yq = fma2(xq, splat2(-7.0f), j);
    const F2 gx0 = p.addc(p.mulc(xq, nsx), nsy);
    const F2 gy0 = p.addc(p.mulc(yq, nsx), nsy);
    float gxa, gxb, gya, gyb;
    unpack2(gx0, gxa, gxb);
    unpack2(gy0, gya, gyb);
    // h = 1 - |x| - |y|;  a0 = b0 + s0*sh with s0 = floor(b0)*2+1 = sign(b0) (b0 in (-1,1), never 0) and
    // sh = -step(h, 0) in {-0, -1}:  b0 when h > 0, else b0 - sign(b0)
    const float ha = __fsub_rn(__fsub_rn(1.0f, fabsf(gxa)), fabsf(gya));
    const float hb = __fsub_rn(__fsub_rn(1.0f, fabsf(gxb)), fabsf(gyb));
    const bool ina = 0.0f < ha, inb = 0.0f < hb;
    const float sxa = ina ? 0.0f : copysignf(1.0f, gxa), sya = ina ? 0.0f : copysignf(1.0f, gya);
    const float sxb = inb ? 0.0f : copysignf(1.0f, gxb), syb = inb ? 0.0f : copysignf(1.0f, gyb);
    F2 gx = p.sub(gx0, pack2(sxa, sxb));
    F2 gy = p.sub(gy0, pack2(sya, syb));
    F2 gz = pack2(ha, hb);
    // normalise: p *= 1.79284291400159 - 0.85373472095314*dot(p,p)
    const F2 nrm = p.sub(splat2(1.79284291400159f), p.mulc(dot3_2(p, gx, gy, gz, gx, gy, gz), 0.85373472095314f));
    gx = p.mul(gx, nrm);
    gy = p.mul(gy, nrm);
    gz = p.mul(gz, nrm);
    // m = max(0.6 - dot(x,x), 0)
    float ma, mb;
    unpack2(p.sub(splat2(0.6f), dot3_2(p, x, y, z, x, y, z)), ma, mb);
    ma = (ma < 0.0f) ? 0.0f : ma;
    mb = (mb < 0.0f) ? 0.0f : mb;
    F2 m = pack2(ma, mb);
    m = p.mul(m, m);
    m = p.mul(m, m);
    return p.mul(m, dot3_2(p, gx, gy, gz, x, y, z));
}

// (snoise(vx, vy, za), snoise(vx, vy, zb))
__device__ __forceinline__ void snoise3_pair(const PackedConsts &k, float vx, float vy, float za, float zb, float &out_a,
                                             float &out_b) {
    const P2 p(k);
    const float Cx = 1.0f / 6.0f, Cy = 1.0f / 3.0f;
    // first corner: i = floor(v + dot(v, C.yyy)); x0 = v - i + dot(i, C.xxx)
    const float dxy = __fadd_rn(__fmul_rn(vx, Cy), __fmul_rn(vy, Cy));       // shared by both lanes
    const F2 vz = pack2(za, zb), vxx = splat2(vx), vyy = splat2(vy);
    const F2 d = p.add(splat2(dxy), p.mulc(vz, Cy));
    F2 ix = floor2(p.add(vxx, d)), iy = floor2(p.add(vyy, d)), iz = floor2(p.add(vz, d));
    const F2 e = p.add(p.add(p.mulc(ix, Cx), p.mulc(iy, Cx)), p.mulc(iz, Cx));
    const F2 x0 = p.add(p.sub(vxx, ix), e), y0 = p.add(p.sub(vyy, iy), e), z0 = p.add(p.sub(vz, iz), e);

    // other corners: g = step(x0.yzx, x0.xyz); l = 1 - g; i1 = min(g, l.zxy); i2 = max(g, l.zxy)
    float x0a, x0b, y0a, y0b, z0a, z0b;
    unpack2(x0, x0a, x0b);
    unpack2(y0, y0a, y0b);
    unpack2(z0, z0a, z0b);
    const float gxa = (x0a < y0a) ? 0.0f : 1.0f, gya = (y0a < z0a) ? 0.0f : 1.0f, gza = (z0a < x0a) ? 0.0f : 1.0f;
    const float gxb = (x0b < y0b) ? 0.0f : 1.0f, gyb = (y0b < z0b) ? 0.0f : 1.0f, gzb = (z0b < x0b) ? 0.0f : 1.0f;
    const float lxa = __fsub_rn(1.0f, gxa), lya = __fsub_rn(1.0f, gya), lza = __fsub_rn(1.0f, gza);
    const float lxb = __fsub_rn(1.0f, gxb), lyb = __fsub_rn(1.0f, gyb), lzb = __fsub_rn(1.0f, gzb);
    // values are 0 or 1: fminf/fmaxf agree with GLSL min/max here
    const F2 i1x = pack2(fminf(gxa, lza), fminf(gxb, lzb)), i1y = pack2(fminf(gya, lxa), fminf(gyb, lxb)),
             i1z = pack2(fminf(gza, lya), fminf(gzb, lyb));
    const F2 i2x = pack2(fmaxf(gxa, lza), fmaxf(gxb, lzb)), i2y = pack2(fmaxf(gya, lxa), fmaxf(gyb, lxb)),
             i2z = pack2(fmaxf(gza, lya), fmaxf(gzb, lyb));

    // permutations.  After mod289 the integers are never -0, so the "+ 0.0" of the vec4(0, i1, i2, 1)
    // constructor is the identity and is dropped.
    ix = mod289_2(p, ix);
    iy = mod289_2(p, iy);
    iz = mod289_2(p, iz);
    F2 pa = permute2(p, iz);
    F2 pb = permute2(p, p.add(iz, i1z));
    F2 pc = permute2(p, p.add(iz, i2z));
    F2 pd = permute2(p, p.addc(iz, 1.0f));
    pa = permute2(p, p.add(pa, iy));
    pb = permute2(p, p.add(p.add(pb, iy), i1y));
    pc = permute2(p, p.add(p.add(pc, iy), i2y));
    pd = permute2(p, p.addc(p.add(pd, iy), 1.0f));
    pa = permute2(p, p.add(pa, ix));
    pb = permute2(p, p.add(p.add(pb, ix), i1x));
    pc = permute2(p, p.add(p.add(pc, ix), i2x));
    pd = permute2(p, p.addc(p.add(pd, ix), 1.0f));

    const F2 t0 = corner2(p, pa, x0, y0, z0);
    const F2 t1 = corner2(p, pb, p.addc(p.sub(x0, i1x), Cx), p.addc(p.sub(y0, i1y), Cx), p.addc(p.sub(z0, i1z), Cx));
    const F2 t2 = corner2(p, pc, p.addc(p.sub(x0, i2x), Cy), p.addc(p.sub(y0, i2y), Cy), p.addc(p.sub(z0, i2z), Cy));
    const F2 t3 = corner2(p, pd, p.addc(x0, -0.5f), p.addc(y0, -0.5f), p.addc(z0, -0.5f));
    unpack2(p.mulc(p.add(p.add(p.add(t0, t1), t2), t3), 42.0f), out_a, out_b);
}

}  // namespace tb
