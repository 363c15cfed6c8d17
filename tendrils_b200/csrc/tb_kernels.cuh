// tb_kernels.cuh -- sm_100a kernels of the Tendrils step: integrate, flow splat, spawners.
// Compiled with -fmad=false; see tb_math.cuh for the arithmetic contract.
#pragma once

#include "tb_math.cuh"
#include "tb_noise2.cuh"
#include "../../include/tendrils_b200.h"

namespace tb {

static constexpr float kInert = -1000000.0f;   // src/const/inert.glsl:1

// Line pair k of a column: which texel row and which buffer each of its 2 vertices samples
// (src/particles.js:171-190 seen through src/state/state-at-frame.glsl:12-22).
struct PairEntry {
    int32_t k;        // pair index within the column = position in draw order
    int32_t row_a;    // texel row of vertex 2k;   bit 31 set: samples the CURRENT buffer
    int32_t row_b;    // texel row of vertex 2k+1; bit 31 set: samples the CURRENT buffer
    int32_t pad;
};

struct IntegrateArgs {
    tb_state S;
    const float4 *__restrict__ in;
    float4 *__restrict__ out;
    const float4 *__restrict__ targets;
    const float4 *__restrict__ flow;
    float2 *__restrict__ wander;           // the two noise values per particle (split launch only)
    int PW, PH, W, H;
    int col0, cols;        // first global column, local columns
    float time, dt;
    int use_targets;       // 0: the target term is provably +-0 for every finite particle
    int use_noise;         // 0: the noise term is provably +-0 for every finite particle
    int packed_noise;      // 1: evaluate the two simplex noises on the packed FP32 pipe (tb_noise2.cuh)
    int pow2_res;          // 1: PW and PH are powers of two: x/res == x*(1/res) exactly
    float inv_resx, inv_resy, inv_n;
    PackedConsts pk;       // 1, -1, -0 (opaque to the compiler on purpose)
};

// [raster-begin]  (tests/test_raster_host.py compiles the text between these markers for the CPU)
// RASTER-1 (spec/PARITY.md): GL_LINES of width 1, centre-sampled along the major axis, half-open
// towards the second vertex, scissored to the grid.  emit(gx, gy, t) per fragment.
template <class Emit>
__device__ __forceinline__ void raster_line(float xa, float ya, float xb, float yb, int W, int H, Emit &&emit) {
    const float dx = __fsub_rn(xb, xa), dy = __fsub_rn(yb, ya);
    const float adx = fabsf(dx), ady = fabsf(dy);
    const bool xmajor = adx >= ady;
    // major/minor axis views
    const float ma = xmajor ? xa : ya, mb = xmajor ? xb : yb, dm = xmajor ? dx : dy;
    const float na = xmajor ? ya : xa, dn = xmajor ? dy : dx;
    const int M = xmajor ? W : H, N = xmajor ? H : W;
    if (!(fabsf(dm) > 0.0f)) return;
    // Candidate columns: a superset of those whose centre i + 0.5 lies within [min, max]; the membership
    // test below is the definition.  lo - 0.5 and hi - 0.5 are exact wherever they matter (0.5 <= v < 2^22).
    float flo = floorf(__fsub_rn(gmin(ma, mb), 0.5f)), fhi = floorf(__fsub_rn(gmax(ma, mb), 0.5f));
    if (flo < 0.0f) flo = 0.0f;
    if (fhi > static_cast<float>(M - 1)) fhi = static_cast<float>(M - 1);
    if (!(flo <= fhi)) return;
    const int ihi = static_cast<int>(fhi);
    for (int i = static_cast<int>(flo); i <= ihi; ++i) {
        const float ic = __fadd_rn(static_cast<float>(i), 0.5f);
        const bool in = (dm > 0.0f) ? (ma <= ic && ic < mb) : (mb < ic && ic <= ma);
        if (!in) continue;
        const float t = __fdiv_rn(__fsub_rn(ic, ma), dm);
        const float nn = __fadd_rn(na, __fmul_rn(t, dn));
        const float fj = floorf(nn);
        if (!(fj >= 0.0f && fj <= static_cast<float>(N - 1))) continue;
        const int j = static_cast<int>(fj);
        emit(xmajor ? i : j, xmajor ? j : i, t);
    }
}

// Loads the two vertices of pair (column, entry) and hands window coordinates + colours on.
// window coordinates of a vertex (PARITY V3) and the cull rules V1/V2 shared by every splat pass
__device__ __forceinline__ bool splat_vertex_ok(const float4 &s) {
    if (!(s.x != kInert || s.y != kInert)) return false;
    return is_finite(s.x) && is_finite(s.y) && is_finite(s.z) && is_finite(s.w);
}
__device__ __forceinline__ uint32_t count_fragments(const float4 &sa, const float4 &sb, float vsx, float vsy, int W, int H) {
    if (!splat_vertex_ok(sa) || !splat_vertex_ok(sb)) return 0u;
    const float hw = __fmul_rn(0.5f, static_cast<float>(W)), hh = __fmul_rn(0.5f, static_cast<float>(H));
    const float xa = __fadd_rn(__fmul_rn(__fmul_rn(sa.x, vsx), hw), hw), ya = __fadd_rn(__fmul_rn(__fmul_rn(sa.y, vsy), hh), hh);
    const float xb = __fadd_rn(__fmul_rn(__fmul_rn(sb.x, vsx), hw), hw), yb = __fadd_rn(__fmul_rn(__fmul_rn(sb.y, vsy), hh), hh);
    // Closed form when the line stays clear of the minor-axis borders: then every candidate column whose
    // centre passes the membership test yields a fragment, interior candidates always pass (their centres
    // lie strictly between the endpoints), and only the two end candidates need testing.  Same count as
    // the loop, without its per-fragment division.
    {
        const float dx = __fsub_rn(xb, xa), dy = __fsub_rn(yb, ya);
        const bool xmajor = fabsf(dx) >= fabsf(dy);
        const float ma = xmajor ? xa : ya, mb = xmajor ? xb : yb, dm = xmajor ? dx : dy;
        const float na = xmajor ? ya : xa, nb = xmajor ? yb : xb;
        const int M = xmajor ? W : H, N = xmajor ? H : W;
        if (!(fabsf(dm) > 0.0f)) return 0u;
        // (an endpoint so far out that its window coordinate overflowed to +-Inf makes every t NaN: the loop below
        // then emits nothing, and so must the count)
        if (gmin(na, nb) >= 1.0f && gmax(na, nb) <= static_cast<float>(N - 1) && is_finite(ma) && is_finite(mb)) {
            float flo = floorf(__fsub_rn(gmin(ma, mb), 0.5f)), fhi = floorf(__fsub_rn(gmax(ma, mb), 0.5f));
            if (flo < 0.0f) flo = 0.0f;
            if (fhi > static_cast<float>(M - 1)) fhi = static_cast<float>(M - 1);
            if (!(flo <= fhi)) return 0u;
            auto member = [&](float fi) {
                const float ic = __fadd_rn(fi, 0.5f);
                return (dm > 0.0f) ? (ma <= ic && ic < mb) : (mb < ic && ic <= ma);
            };
            uint32_t n = static_cast<uint32_t>(static_cast<int>(fhi) - static_cast<int>(flo)) + 1u;
            if (!member(flo)) --n;
            if (fhi != flo && !member(fhi)) --n;
            return n;
        }
    }
    uint32_t n = 0;
    raster_line(xa, ya, xb, yb, W, H, [&](int, int, float) { ++n; });
    return n;
}
// [raster-end]

#ifndef TB_INTEGRATE_MIN_BLOCKS
#define TB_INTEGRATE_MIN_BLOCKS 5
#endif

// logic.frag:45-101 comes in three launch shapes sharing one body:
//   kFused    the whole shader in one pass (16 B in, 16 B out, one 16 B L2 gather);
//   kNoise    only the two simplex noises (logic.frag:57-68) -> `wander` (8 B/particle).  They do not
//             depend on the flow grid, so tb_step runs this on a side stream while the PREVIOUS
//             step's flow splat (sort + fold, HBM-bound) still occupies the main stream;
//   kFinish   everything else, reading `wander` back (logic.frag:71-100).
// The arithmetic is the same expression tree in every shape: results are bit-identical.
enum IntegrateMode { kFused = 0, kNoise = 1, kFinish = 2 };

// Grid: x = 256-thread tiles along a texture column (y), y = local column.  No integer division.
template <int MODE>
__global__ void __launch_bounds__(256, TB_INTEGRATE_MIN_BLOCKS) k_integrate(const IntegrateArgs A) {
    const int y = blockIdx.x * 256 + threadIdx.x;
    if (y >= A.PH) return;
    const long long l = static_cast<long long>(blockIdx.y) * A.PH + y;
    const float4 st = __ldcs(A.in + l);
    float posx = st.x, posy = st.y, velx = st.z, vely = st.w;
    if (!(posx != kInert || posy != kInert)) {
        if (MODE != kNoise) __stcs(A.out + l, st);
        return;
    }
    const int x = A.col0 + static_cast<int>(blockIdx.y);
    const float resx = static_cast<float>(A.PW), resy = static_cast<float>(A.PH);
    const float fcx = __fadd_rn(static_cast<float>(x), 0.5f), fcy = __fadd_rn(static_cast<float>(y), 0.5f);
    float uvx, uvy, i;
    if (A.pow2_res) {      // dividing by a power of two is the same rounding as multiplying by its reciprocal
        uvx = __fmul_rn(fcx, A.inv_resx);
        uvy = __fmul_rn(fcy, A.inv_resy);
        i = __fmul_rn(__fadd_rn(fcx, __fmul_rn(fcy, resx)), A.inv_n);
    } else {
        uvx = __fdiv_rn(fcx, resx);
        uvy = __fdiv_rn(fcy, resy);
        i = __fdiv_rn(__fadd_rn(fcx, __fmul_rn(fcy, resx)), __fmul_rn(resx, resy));
    }
    const tb_state &S = A.S;

    const bool tame = fabsf(posx) < 1.0e6f && fabsf(posy) < 1.0e6f;
    float wx = 0.0f, wy = 0.0f;
    if (MODE == kFinish) {
        const float2 w = __ldcs(A.wander + l);
        wx = w.x;
        wy = w.y;
    } else if (A.use_noise || !tame) {
        const float ns = vary(S.noiseScale, i, S.varyNoiseScale);
        const float npx = __fmul_rn(posx, ns), npy = __fmul_rn(posy, ns);
        const float noiseTime = __fmul_rn(A.time, vary(S.noiseSpeed, i, S.varyNoiseSpeed));
        const float za = __fadd_rn(uvx, noiseTime), zb = __fadd_rn(__fadd_rn(uvy, noiseTime), 1234.5678f);
        // the packed path relies on lattice coordinates being exact integers below 2^24
        const bool lattice_ok = fabsf(npx) < 2.0e6f && fabsf(npy) < 2.0e6f && fabsf(za) < 2.0e6f && fabsf(zb) < 2.0e6f;
        if (A.packed_noise && lattice_ok) {
            snoise3_pair(A.pk, npx, npy, za, zb, wx, wy);
        } else {
            wx = snoise3(npx, npy, za);
            wy = snoise3(npx, npy, zb);
        }
    }
    if (MODE == kNoise) {
        __stcs(A.wander + l, make_float2(wx, wy));
        return;
    }

    // flowAtScreenPos (flow/flow-at-screen-pos.glsl:13-27) with levels = stride = 1
    const float spx = __fmul_rn(posx, S.viewSize[0]), spy = __fmul_rn(posy, S.viewSize[1]);
    const float fu = __fadd_rn(0.0f, __fdiv_rn(__fmul_rn(1.0f, __fsub_rn(spx, -1.0f)), 2.0f));
    const float fv = __fadd_rn(0.0f, __fdiv_rn(__fmul_rn(1.0f, __fsub_rn(spy, -1.0f)), 2.0f));
    const float4 fd = __ldg(A.flow + (static_cast<size_t>(texel_of(fv, A.H)) * A.W + texel_of(fu, A.W)));
    const float fac = gmax(0.0f, __fsub_rn(1.0f, __fmul_rn(__fsub_rn(A.time, fd.z), S.flowDecay)));
    float ffx = __fadd_rn(0.0f, __fmul_rn(__fmul_rn(fd.x, fac), 1.0f));
    float ffy = __fadd_rn(0.0f, __fmul_rn(__fmul_rn(fd.y, fac), 1.0f));
    ffx = __fdiv_rn(ffx, 1.0f);
    ffy = __fdiv_rn(ffy, 1.0f);

    const float vforce = vary(S.forceWeight, i, S.varyForce);
    const float vflow = vary(S.flowWeight, i, S.varyFlow);
    const float vnoise = vary(S.noiseWeight, i, S.varyNoise);
    float nvx = __fadd_rn(__fmul_rn(__fmul_rn(velx, S.damping), A.dt),
                          __fmul_rn(vforce, __fadd_rn(__fmul_rn(__fmul_rn(ffx, A.dt), vflow),
                                                      __fmul_rn(__fmul_rn(wx, A.dt), vnoise))));
    float nvy = __fadd_rn(__fmul_rn(__fmul_rn(vely, S.damping), A.dt),
                          __fmul_rn(vforce, __fadd_rn(__fmul_rn(__fmul_rn(ffy, A.dt), vflow),
                                                      __fmul_rn(__fmul_rn(wy, A.dt), vnoise))));
    if (A.use_targets || !tame) {
        const float4 tg = __ldcs(A.targets + l);
        const float vt = vary(S.target, i, S.varyTarget);
        nvx = __fadd_rn(nvx, __fmul_rn(__fsub_rn(tg.x, posx), vt));
        nvy = __fadd_rn(nvy, __fmul_rn(__fsub_rn(tg.y, posy), vt));
    }
    const float speed = glength(nvx, nvy);
    const float sc = __fdiv_rn(gmin(speed, S.speedLimit), speed);
    nvx = __fmul_rn(nvx, sc);
    nvy = __fmul_rn(nvy, sc);
    const float4 nst = make_float4(__fadd_rn(posx, nvx), __fadd_rn(posy, nvy), nvx, nvy);
    __stcs(A.out + l, nst);
}

constexpr int kMaxBandRanks = 16;     // ranks of a column-sharded run (one node)

// Full-grid alpha-over of an RGBA layer (L4 inputs drawn into the flow FBO).
__global__ void k_blend_layer(float4 *__restrict__ flow, const float4 *__restrict__ layer, int G) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= G) return;
    const float4 s = layer[t];
    float4 d = flow[t];
    const float a = s.w, om = __fsub_rn(1.0f, a);
    d.x = __fadd_rn(__fmul_rn(s.x, a), __fmul_rn(d.x, om));
    d.y = __fadd_rn(__fmul_rn(s.y, a), __fmul_rn(d.y, om));
    d.z = __fadd_rn(__fmul_rn(s.z, a), __fmul_rn(d.z, om));
    d.w = __fadd_rn(__fmul_rn(s.w, a), __fmul_rn(d.w, om));
    flow[t] = d;
}

// f1: optical flow of two RGBA8 frames, alpha-over blended into the flow grid
// (src/optical-flow/index.frag:55-81 drawn with the big triangle of src/screen/index.vert).
struct OpticalArgs {
    float4 *__restrict__ flow;
    const uchar4 *__restrict__ view;
    const uchar4 *__restrict__ last;
    int W, H, IW, IH;
    tb_optical_flow_params U;
};

__device__ __forceinline__ float of_gray(const uchar4 *__restrict__ tex, int w, int h, float u, float v) {
    const uchar4 t = __ldg(tex + (static_cast<size_t>(texel_of(v, h)) * w + texel_of(u, w)));
    const float r = __fdiv_rn(static_cast<float>(t.x), 255.0f), g = __fdiv_rn(static_cast<float>(t.y), 255.0f),
                b = __fdiv_rn(static_cast<float>(t.z), 255.0f);
    return dot3(r, g, b, 0.3f, 0.59f, 0.11f);                 // utils/gray-scale.glsl
}

__global__ void __launch_bounds__(256) k_optical_flow(const OpticalArgs A) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= A.W * A.H) return;
    const int gy = t / A.W, gx = t - gy * A.W;
    const tb_optical_flow_params &U = A.U;
    // the varying uv of the big triangle at the fragment centre (PARITY OF1)
    const float uvx = __fsub_rn(__fmul_rn(__fdiv_rn(__fadd_rn(static_cast<float>(gx), 0.5f), static_cast<float>(A.W)), 2.0f), 1.0f);
    const float uvy = __fsub_rn(__fmul_rn(__fdiv_rn(__fadd_rn(static_cast<float>(gy), 0.5f), static_cast<float>(A.H)), 2.0f), 1.0f);
    const float px = __fdiv_rn(__fmul_rn(uvx, U.scaleUV[0]), U.viewSize[0]), py = __fdiv_rn(__fmul_rn(uvy, U.scaleUV[1]), U.viewSize[1]);
    const float su = __fadd_rn(0.0f, __fdiv_rn(__fmul_rn(1.0f, __fsub_rn(px, -1.0f)), 2.0f));
    const float sv = __fadd_rn(0.0f, __fdiv_rn(__fmul_rn(1.0f, __fsub_rn(py, -1.0f)), 2.0f));
    const float up = __fadd_rn(su, U.offset), um = __fsub_rn(su, U.offset), u0p = __fadd_rn(su, 0.0f), u0m = __fsub_rn(su, 0.0f);
    const float vp = __fadd_rn(sv, U.offset), vm = __fsub_rn(sv, U.offset), v0p = __fadd_rn(sv, 0.0f), v0m = __fsub_rn(sv, 0.0f);
    const float gradX = __fadd_rn(__fsub_rn(of_gray(A.view, A.IW, A.IH, up, v0p), of_gray(A.view, A.IW, A.IH, um, v0m)),
                                  __fsub_rn(of_gray(A.last, A.IW, A.IH, up, v0p), of_gray(A.last, A.IW, A.IH, um, v0m)));
    const float gradY = __fadd_rn(__fsub_rn(of_gray(A.view, A.IW, A.IH, u0p, vp), of_gray(A.view, A.IW, A.IH, u0m, vm)),
                                  __fsub_rn(of_gray(A.last, A.IW, A.IH, u0p, vp), of_gray(A.last, A.IW, A.IH, u0m, vm)));
    const float gradMag = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(gradX, gradX), __fmul_rn(gradY, gradY)), U.lambda));
    const float diff = __fsub_rn(of_gray(A.view, A.IW, A.IH, su, sv), of_gray(A.last, A.IW, A.IH, su, sv));
    const float vx = __fmul_rn(__fmul_rn(diff, __fdiv_rn(gradX, gradMag)), U.speed);
    const float vy = __fmul_rn(__fmul_rn(diff, __fdiv_rn(gradY, gradMag)), U.speed);
    const float tt = __fdiv_rn(glength(vx, vy), U.speedLimit), ut = __fsub_rn(1.0f, tt);
    // bezier(vec3(0, 0, 1), t) (utils/bezier.glsl:9-13)
    const float bz = __fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(0.0f, ut), __fmul_rn(0.0f, tt)), ut),
                               __fmul_rn(__fadd_rn(__fmul_rn(0.0f, ut), __fmul_rn(1.0f, tt)), tt));
    const float ox = __fmul_rn(bz, vx), oy = __fmul_rn(bz, vy);
    const float a = gmin(__fdiv_rn(glength(ox, oy), U.speedLimit), 1.0f), om = __fsub_rn(1.0f, a);
    float4 d = A.flow[t];
    d.x = __fadd_rn(__fmul_rn(ox, a), __fmul_rn(d.x, om));
    d.y = __fadd_rn(__fmul_rn(oy, a), __fmul_rn(d.y, om));
    d.z = __fadd_rn(__fmul_rn(U.time, a), __fmul_rn(d.z, om));
    d.w = __fadd_rn(__fmul_rn(a, a), __fmul_rn(d.w, om));
    A.flow[t] = d;
}

// ------------------------------------------------------------------------------------------
// Spawners (a12-a15)
// ------------------------------------------------------------------------------------------
struct SpawnArgs {
    float4 *__restrict__ out;
    const float4 *__restrict__ state;   // `particles` sampler = buffers[1]
    const float4 *__restrict__ image;   // spawnData
    int PW, PH, IW, IH;
    int image_xmajor;                   // spawnData is a particle buffer (texel (x,y) at x*IH+y)
    long long p0, n;
    tb_pixel_spawner U;
    float time, flowDecay;
    float radius, speed;
    int apply, vignette, samples;       // pixel spawner composition
};

static constexpr float kTau = 6.28318530717958647692f;
enum { APPLY_COLOR = 0, APPLY_BRIGHTEST = 1, APPLY_IDENTITY = 2, APPLY_FLOW = 3 };

__global__ void k_spawn_init(float4 *__restrict__ out, long long n) {          // spawn/init/index.frag:5-10
    const long long l = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (l < n) out[l] = make_float4(kInert, kInert, 0.0f, 0.0f);
}

__global__ void k_spawn_ball(const SpawnArgs A) {                                // spawn/ball/index.frag:11-19
    const long long l = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (l >= A.n) return;
    const long long p = A.p0 + l;
    const int x = static_cast<int>(p / A.PH), y = static_cast<int>(p - static_cast<long long>(x) * A.PH);
    const float fx = __fadd_rn(static_cast<float>(x), 0.5f), fy = __fadd_rn(static_cast<float>(y), 0.5f);
    const float r0 = grandom(__fadd_rn(__fmul_rn(fx, 1.7654f), 2.3675f), __fadd_rn(__fmul_rn(fy, 1.7654f), 2.3675f));
    const float r1 = grandom(__fadd_rn(__fmul_rn(fx, 1.23494f), 0.36434f), __fadd_rn(__fmul_rn(fy, 1.23494f), 0.36434f));
    const float r2 = grandom(__fadd_rn(__fmul_rn(fx, 0.327789f), 3.498787f), __fadd_rn(__fmul_rn(fy, 0.327789f), 3.498787f));
    const float r3 = grandom(__fadd_rn(__fmul_rn(fx, 9.0374f), 0.2773f), __fadd_rn(__fmul_rn(fy, 9.0374f), 0.2773f));
    float s0, c0, s1, c1;
    sincos_t1(__fmul_rn(r0, kTau), s0, c0);
    sincos_t1(__fmul_rn(r2, kTau), s1, c1);
    A.out[l] = make_float4(__fmul_rn(__fmul_rn(c0, r1), A.radius), __fmul_rn(__fmul_rn(s0, r1), A.radius),
                           __fmul_rn(__fmul_rn(c1, r3), A.speed), __fmul_rn(__fmul_rn(s1, r3), A.speed));
}

// spawn/pixels/frag/head.frag:28-34
__device__ __forceinline__ float2 spawn_to_pos(const tb_pixel_spawner &U, float u, float v, float time) {
    const float tt = __fmul_rn(time, 0.001f);
    const float ox = gmix(-U.jitter[0], U.jitter[0],
                          grandom(__fadd_rn(__fsub_rn(u, 1.2345f), tt), __fadd_rn(__fsub_rn(v, 1.2345f), tt)));
    const float oy = gmix(-U.jitter[1], U.jitter[1],
                          grandom(__fadd_rn(__fadd_rn(u, 1.2345f), tt), __fadd_rn(__fadd_rn(v, 1.2345f), tt)));
    const float uu = __fadd_rn(u, ox), vv = __fadd_rn(v, oy);
    // uvToPos = glsl-map(uv, 0, 1, -1, 1)
    float qx = __fadd_rn(-1.0f, __fdiv_rn(__fmul_rn(2.0f, __fsub_rn(uu, 0.0f)), 1.0f));
    float qy = __fadd_rn(-1.0f, __fdiv_rn(__fmul_rn(2.0f, __fsub_rn(vv, 0.0f)), 1.0f));
    qx = __fmul_rn(__fmul_rn(qx, 1.0f), U.spawnSize[0]);
    qy = __fmul_rn(__fmul_rn(qy, -1.0f), U.spawnSize[1]);
    const float *m = U.spawnMatrix;
    return make_float2(__fadd_rn(__fadd_rn(__fmul_rn(m[0], qx), __fmul_rn(m[3], qy)), __fmul_rn(m[6], 1.0f)),
                       __fadd_rn(__fadd_rn(__fmul_rn(m[1], qx), __fmul_rn(m[4], qy)), __fmul_rn(m[7], 1.0f)));
}

// filter/vignette.glsl:5-24, spawn/pixels/vignette-head.glsl:4-6, utils/bezier.glsl:9-13
__device__ __forceinline__ float vignette(float u, float v) {
    const float amount = gmin(__fsub_rn(1.0f, __fdiv_rn(glength(__fsub_rn(u, 0.5f), __fsub_rn(v, 0.5f)), 0.6f)), 1.0f);
    const float t = amount, ut = __fsub_rn(1.0f, t);
    const float l = __fmul_rn(__fadd_rn(__fmul_rn(0.1f, ut), __fmul_rn(1.0f, t)), ut);
    const float r = __fmul_rn(__fadd_rn(__fmul_rn(1.0f, ut), __fmul_rn(1.0f, t)), t);
    return gmax(0.0f, __fadd_rn(l, r));
}

// libs/glsl-hsv/rgb-hsv.glsl:4-11
__device__ __forceinline__ float3 rgb2hsv(float r, float g, float b) {
    const float ky = -1.0f / 3.0f, kz = 2.0f / 3.0f, e = 1.0e-10f;
    float p0, p1, p2, p3;
    if (g < b) { p0 = b; p1 = g; p2 = -1.0f; p3 = kz; } else { p0 = g; p1 = b; p2 = 0.0f; p3 = ky; }
    float q0, q1, q2, q3;
    if (r < p0) { q0 = p0; q1 = p1; q2 = p3; q3 = r; } else { q0 = r; q1 = p1; q2 = p2; q3 = p0; }
    const float d = __fsub_rn(q0, gmin(q3, q1));
    const float h = fabsf(__fadd_rn(q2, __fdiv_rn(__fsub_rn(q3, q1), __fadd_rn(__fmul_rn(6.0f, d), e))));
    return make_float3(h, __fdiv_rn(d, __fadd_rn(q0, e)), q0);
}

__device__ __forceinline__ float4 fetch_image(const SpawnArgs &A, float u, float v) {
    const int tx = texel_of(u, A.IW), ty = texel_of(v, A.IH);
    const size_t idx = A.image_xmajor ? (static_cast<size_t>(tx) * A.IH + ty) : (static_cast<size_t>(ty) * A.IW + tx);
    return __ldg(A.image + idx);
}

// apply/<kind>.glsl, optionally composed with filter/pass/vignette.glsl (apply/compose-filter.glsl)
__device__ __forceinline__ float4 apply_pixel(const SpawnArgs &A, float u, float v, float2 pos, float4 px) {
    if (A.vignette) {
        const float w = vignette(u, v);
        px = make_float4(__fmul_rn(px.x, w), __fmul_rn(px.y, w), __fmul_rn(px.z, w), __fmul_rn(px.w, w));
    }
    float s, c;
    if (A.apply == APPLY_COLOR) {                  // apply/color.glsl:13-17
        const float3 hsv = rgb2hsv(px.x, px.y, px.z);
        sincos_t1(__fmul_rn(__fadd_rn(hsv.x, __fmul_rn(A.time, 0.00003f)), kTau), s, c);
        return make_float4(pos.x, pos.y, __fmul_rn(__fmul_rn(__fmul_rn(c, hsv.y), hsv.z), px.w),
                           __fmul_rn(__fmul_rn(__fmul_rn(s, hsv.y), hsv.z), px.w));
    }
    if (A.apply == APPLY_BRIGHTEST) {              // apply/brightest.glsl:12-16, glsl-luma
        const float dd = __fadd_rn(__fmul_rn(px.x, px.z), __fmul_rn(px.y, px.w));
        const float ang = __fmul_rn(gmod(grandom(__fmul_rn(u, dd), __fmul_rn(v, dd)), 1.0f), kTau);
        const float luma = dot3(px.x, px.y, px.z, 0.299f, 0.587f, 0.114f);
        sincos_t1(ang, s, c);
        return make_float4(pos.x, pos.y, __fmul_rn(__fmul_rn(c, luma), px.w), __fmul_rn(__fmul_rn(s, luma), px.w));
    }
    if (A.apply == APPLY_IDENTITY) return px;      // apply/identity.glsl
    const float fac = gmax(0.0f, __fsub_rn(1.0f, __fmul_rn(__fsub_rn(A.time, px.z), A.flowDecay)));
    return make_float4(pos.x, pos.y, __fmul_rn(px.x, fac), __fmul_rn(px.y, fac));   // apply/flow.glsl
}

// a14: spawn/pixels/index.frag -> frag/direct-main.frag:9-20
__global__ void k_spawn_direct(const SpawnArgs A) {
    const long long l = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (l >= A.n) return;
    const long long p = A.p0 + l;
    const int x = static_cast<int>(p / A.PH), y = static_cast<int>(p - static_cast<long long>(x) * A.PH);
    const float pw = static_cast<float>(A.PW), ph = static_cast<float>(A.PH);
    const float u = __fmul_rn(__fdiv_rn(__fadd_rn(static_cast<float>(x), 0.5f), pw), __fdiv_rn(pw, pw));
    const float v = __fmul_rn(__fdiv_rn(__fadd_rn(static_cast<float>(y), 0.5f), ph),
                              __fdiv_rn(static_cast<float>(2 * A.PH), ph));
    const float2 pos = spawn_to_pos(A.U, u, v, A.time);
    const float4 st = apply_pixel(A, u, v, pos, fetch_image(A, u, v));
    A.out[l] = make_float4(st.x, st.y, __fmul_rn(st.z, A.U.speed), __fmul_rn(st.w, A.U.speed));
}

// a15: spawn/pixels/*-sample.frag -> frag/best-sample-main.frag:21-46, test/particles.glsl:8-10
__global__ void k_spawn_sample(const SpawnArgs A) {
    const long long l = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (l >= A.n) return;
    const long long p = A.p0 + l;
    const int x = static_cast<int>(p / A.PH), y = static_cast<int>(p - static_cast<long long>(x) * A.PH);
    const float u = __fdiv_rn(__fadd_rn(static_cast<float>(x), 0.5f), static_cast<float>(A.PW));
    const float v = __fdiv_rn(__fadd_rn(static_cast<float>(y), 0.5f), static_cast<float>(A.PH));
    float4 st = A.state[l];
    const float k0 = __fadd_rn(1.2345f, __fmul_rn(A.time, 0.001f));
    const float b0 = __fadd_rn(__fadd_rn(st.x, u), k0), b1 = __fadd_rn(__fadd_rn(st.y, v), k0);
    const float b2 = __fadd_rn(__fadd_rn(st.z, u), k0), b3 = __fadd_rn(__fadd_rn(st.w, v), k0);
    for (int n = 0; n < A.samples; ++n) {
        const float fn = static_cast<float>(n);
        const float su = gmod(grandom(__fadd_rn(b0, fn), __fadd_rn(b1, fn)), 1.0f);
        const float sv = gmod(grandom(__fadd_rn(b2, fn), __fadd_rn(b3, fn)), 1.0f);
        const float2 pos = spawn_to_pos(A.U, su, sv, A.time);
        float4 o = apply_pixel(A, su, sv, pos, fetch_image(A, su, sv));
        o.z = __fmul_rn(o.z, A.U.speed);
        o.w = __fmul_rn(o.w, A.U.speed);
        const float tc = __fadd_rn(__fmul_rn(st.z, st.z), __fmul_rn(st.w, st.w));
        const float tn = __fadd_rn(__fmul_rn(o.z, o.z), __fmul_rn(o.w, o.w));
        if (!(tc > __fmul_rn(A.U.bias, tn))) st = o;
    }
    A.out[l] = st;
}

// sets *flag to 1 if any component of the buffer is non-finite
__global__ void k_check_finite(const float4 *__restrict__ buf, long long n, int *flag) {
    const long long l = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (l >= n) return;
    const float4 v = buf[l];
    if (!(is_finite(v.x) && is_finite(v.y) && is_finite(v.z) && is_finite(v.w))) *flag = 1;
}

}  // namespace tb
