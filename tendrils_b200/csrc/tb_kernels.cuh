// tb_kernels.cuh -- sm_100a kernels of the Tendrils step: integrate, flow splat, spawners.
// Compiled with -fmad=false; see tb_math.cuh for the arithmetic contract.
#pragma once

#include "tb_math.cuh"
#include "tb_noise2.cuh"
#include "../../include/tendrils_b200.h"

namespace tb {

static constexpr float kInert = -1000000.0f;   // src/const/inert.glsl:1

// One fragment of the flow splat: the interpolated colour's (vel.xy, alpha).  The colour's time
// channel is the uniform `time` at both vertices, hence constant along the line.  Fragments
// are generated in primitive (draw) order p = x*PH + k (src/particles.js:182-186) and stably
// sorted by texel, so the order within a texel's segment is the draw order.
#ifndef TB_FRAG_BYTES
#define TB_FRAG_BYTES 16
#endif
#if TB_FRAG_BYTES == 16
struct __align__(16) FragVal {   // padded to 16 B: one 128-bit load/store per fragment in emit, sort and fold
    float cx, cy;     // interpolated vel.xy
    float a;          // interpolated alpha
    float pad;
};
__device__ __forceinline__ FragVal load_frag(const FragVal *p) {
    const float4 v = __ldcs(reinterpret_cast<const float4 *>(p));
    FragVal f; f.cx = v.x; f.cy = v.y; f.a = v.z; f.pad = 0.f;
    return f;
}
__device__ __forceinline__ void store_frag(FragVal *p, float cx, float cy, float a) {
    *reinterpret_cast<float4 *>(p) = make_float4(cx, cy, a, 0.f);
}
#else
struct FragVal {
    float cx, cy;     // interpolated vel.xy
    float a;          // interpolated alpha
};
__device__ __forceinline__ FragVal load_frag(const FragVal *p) {
    FragVal f; f.cx = __ldcs(&p->cx); f.cy = __ldcs(&p->cy); f.a = __ldcs(&p->a);
    return f;
}
__device__ __forceinline__ void store_frag(FragVal *p, float cx, float cy, float a) { p->cx = cx; p->cy = cy; p->a = a; }
#endif

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
    // fused first pass of the flow splat: fragments of this particle's line, in draw order
    const int32_t *__restrict__ row_pair;   // per row: 0xffffffff, or pair index | kind << 30 (1: prev->cur, 2: cur->prev)
    uint32_t *__restrict__ prim_off;        // zeroed beforehand; null when the count is not fused
    int n_pairs;
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
    // Pass 1 of the flow splat, fused: previous and current state of this particle are both in registers.
    if (A.prim_off != nullptr) {
        const uint32_t rp = static_cast<uint32_t>(__ldg(A.row_pair + y));
        if (rp != 0xffffffffu) {
            const bool forward = (rp >> 30) == 1u;
            const uint32_t n = count_fragments(forward ? st : nst, forward ? nst : st, S.viewSize[0], S.viewSize[1], A.W, A.H);
            A.prim_off[static_cast<size_t>(blockIdx.y) * A.n_pairs + (rp & 0x3fffffffu)] = n;
        }
    }
}

// ------------------------------------------------------------------------------------------
// Flow splat (a7-a10).  RASTER-1 (spec/PARITY.md): GL_LINES of width 1, centre-sampled along
// the major axis, half-open towards the second vertex, scissored to the grid.
// ------------------------------------------------------------------------------------------
constexpr uint32_t kOpaqueBit = 0x80000000u;

struct SplatArgs {
    const float4 *__restrict__ cur;
    const float4 *__restrict__ prev;
    const PairEntry *__restrict__ pairs;
    int n_pairs;           // active (non-degenerate) pairs per column
    int PH;
    int cols;              // local columns
    int W, H;
    float vsx, vsy, speedLimit;
    uint32_t *prim_off;    // count pass: fragments per primitive; after the scan: first slot
    uint32_t *keys;        // texel index of every fragment, in draw order
    FragVal *vals;
    uint32_t cap;          // capacity of keys/vals
    const uint32_t *total; // device: total fragments of this collect (after the scan)
};

// Threads are numbered in draw order: tid = local column * n_pairs + index of the active pair.
template <class Body>
__device__ __forceinline__ void splat_pair(const SplatArgs &A, Body &&body) {
    const long long tid = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (tid >= static_cast<long long>(A.cols) * A.n_pairs) return;
    const int xl = static_cast<int>(tid / A.n_pairs);
    const PairEntry pe = A.pairs[tid - static_cast<long long>(xl) * A.n_pairs];
    const size_t base = static_cast<size_t>(xl) * A.PH;
    const float4 sa = __ldg(((pe.row_a < 0) ? A.cur : A.prev) + base + (pe.row_a & 0x7fffffff));
    const float4 sb = __ldg(((pe.row_b < 0) ? A.cur : A.prev) + base + (pe.row_b & 0x7fffffff));
    // inert vertices leave gl_Position unwritten: culled (V1); non-finite vertices: culled (V2)
    if (!(sa.x != kInert || sa.y != kInert)) return;
    if (!(sb.x != kInert || sb.y != kInert)) return;
    if (!(is_finite(sa.x) && is_finite(sa.y) && is_finite(sa.z) && is_finite(sa.w))) return;
    if (!(is_finite(sb.x) && is_finite(sb.y) && is_finite(sb.z) && is_finite(sb.w))) return;
    const float hw = __fmul_rn(0.5f, static_cast<float>(A.W)), hh = __fmul_rn(0.5f, static_cast<float>(A.H));
    const float xa = __fadd_rn(__fmul_rn(__fmul_rn(sa.x, A.vsx), hw), hw);
    const float ya = __fadd_rn(__fmul_rn(__fmul_rn(sa.y, A.vsy), hh), hh);
    const float xb = __fadd_rn(__fmul_rn(__fmul_rn(sb.x, A.vsx), hw), hw);
    const float yb = __fadd_rn(__fmul_rn(__fmul_rn(sb.y, A.vsy), hh), hh);
    body(tid, xa, ya, xb, yb, sa, sb);
}

// Pass 1: fragments per primitive (culled primitives count 0; prim_off is zeroed beforehand).
__global__ void __launch_bounds__(256) k_splat_count(const SplatArgs A) {
    splat_pair(A, [&](long long tid, float xa, float ya, float xb, float yb, const float4 &, const float4 &) {
        uint32_t n = 0;
        raster_line(xa, ya, xb, yb, A.W, A.H, [&](int, int, float) { ++n; });
        A.prim_off[tid] = n;
    });
}

// Pass 1 for a SUBSET of the pairs -- the few that cannot ride in k_integrate (TB_FUSE_PARTIAL): thread = (local
// column, odd pair).  Same count as k_splat_count (count_fragments == the enumeration, tests/test_raster_host.py).
__global__ void __launch_bounds__(256) k_splat_count_odd(const SplatArgs A, const int32_t *__restrict__ odd, int n_odd) {
    const long long tid = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (tid >= static_cast<long long>(A.cols) * n_odd) return;
    const int xl = static_cast<int>(tid / n_odd);
    const int pi = odd[tid - static_cast<long long>(xl) * n_odd];
    const PairEntry pe = A.pairs[pi];
    const size_t base = static_cast<size_t>(xl) * A.PH;
    const float4 sa = __ldg(((pe.row_a < 0) ? A.cur : A.prev) + base + (pe.row_a & 0x7fffffff));
    const float4 sb = __ldg(((pe.row_b < 0) ? A.cur : A.prev) + base + (pe.row_b & 0x7fffffff));
    A.prim_off[static_cast<size_t>(xl) * A.n_pairs + pi] = count_fragments(sa, sb, A.vsx, A.vsy, A.W, A.H);
}

// Pass 2 (after the exclusive scan of prim_off): write (texel, colour) of every fragment at its
// slot.  No atomics: the slot order IS the draw order.
__global__ void __launch_bounds__(256) k_splat_emit(const SplatArgs A) {
    if (*A.total > A.cap) return;            // host re-runs the collect with a larger buffer
    splat_pair(A, [&](long long tid, float xa, float ya, float xb, float yb, const float4 &sa, const float4 &sb) {
        // flow(vel, speedLimit): src/flow/apply/state.glsl:5-16
        const float aa = gmin(__fdiv_rn(glength(sa.z, sa.w), A.speedLimit), 1.0f);
        const float ab = gmin(__fdiv_rn(glength(sb.z, sb.w), A.speedLimit), 1.0f);
        uint32_t slot = A.prim_off[tid];
        raster_line(xa, ya, xb, yb, A.W, A.H, [&](int gx, int gy, float t) {
            const float fcx = __fadd_rn(sa.z, __fmul_rn(t, __fsub_rn(sb.z, sa.z)));
            const float fcy = __fadd_rn(sa.w, __fmul_rn(t, __fsub_rn(sb.w, sa.w)));
            const float fa = __fadd_rn(aa, __fmul_rn(t, __fsub_rn(ab, aa)));
            // bit 31 flags a fragment whose alpha is exactly 1; it rides along the sort (which only
            // looks at the texel bits) and lets the fold skip everything such a fragment overwrites
            A.keys[slot] = (static_cast<uint32_t>(gy) * static_cast<uint32_t>(A.W) + static_cast<uint32_t>(gx)) |
                           (fa == 1.0f ? kOpaqueBit : 0u);
            store_frag(A.vals + slot, fcx, fcy, fa);
            ++slot;
        });
    });
}

// Pass 4 (after the stable radix sort by texel): for every texel that has fragments,
// seg[2t+1] = end of its segment and seg[2t] = where its fold starts: the segment's first
// fragment, or the LAST fragment with alpha == 1 if there is one.  Such a fragment makes all
// earlier ones irrelevant: dst = c*1 + dst*0 = c + (+-0) for every finite dst, and NaN for a
// non-finite dst whether or not the earlier fragments were applied (Inf and NaN never become
// finite under this blend) -- so folding from it onto the ORIGINAL texel is exact.
// seg is zero-filled beforehand (empty segments stay [0,0)).
__global__ void __launch_bounds__(256) k_splat_bounds(const uint32_t *__restrict__ keys, uint32_t n, uint32_t *__restrict__ seg) {
    // four consecutive keys per thread (one 128-bit load) plus the two neighbours
    const uint32_t i0 = (blockIdx.x * blockDim.x + threadIdx.x) * 4u;
    if (i0 >= n) return;
    uint32_t k[6];
    k[0] = i0 ? keys[i0 - 1] : 0u;
    if (i0 + 4 <= n && (reinterpret_cast<uintptr_t>(keys + i0) & 15u) == 0) {   // pieces of an exchange may start unaligned
        const uint4 v = *reinterpret_cast<const uint4 *>(keys + i0);
        k[1] = v.x; k[2] = v.y; k[3] = v.z; k[4] = v.w;
    } else {
        for (uint32_t j = 0; j < 4; ++j) k[1 + j] = (i0 + j < n) ? keys[i0 + j] : 0u;
    }
    k[5] = (i0 + 4 < n) ? keys[i0 + 4] : 0u;
#pragma unroll
    for (uint32_t j = 0; j < 4; ++j) {
        const uint32_t i = i0 + j;
        if (i >= n) break;
        const uint32_t raw = k[1 + j], t = raw & ~kOpaqueBit;
        const bool first = (i == 0) || ((k[j] & ~kOpaqueBit) != t);
        if (first || (raw & kOpaqueBit)) atomicMax(&seg[2 * t], i);
        if (i == n - 1 || (k[2 + j] & ~kOpaqueBit) != t) seg[2 * t + 1] = i + 1;
    }
}

// Pass 5: ordered alpha-over fold: blendFunc(SRC_ALPHA, ONE_MINUS_SRC_ALPHA) on all four
// channels, in primitive order (src/index.js:267-268):  dst = src*a + dst*(1-a).
//
// One warp owns 32 consecutive texels, whose sorted segments are one contiguous range of the
// fragment array.  The warp streams that range through shared memory in chunks: all lanes load
// fragments coalesced and pre-multiply the order-independent part (src*a per channel, 1-a); then
// each lane folds the part of ITS texel's segment that lies in the chunk, reading shared memory.
// The only serial work left is the two dependent roundings per fragment the blend demands, so a
// texel that holds a whole chunk (a hot spot) runs at the latency of that chain, not of DRAM.
// Texels whose segment is longer than kFoldHot fragments ("hot": dense filaments, the centre of
// a ball spawn) would make their whole warp wait; the lane-per-texel kernel hands them to
// k_splat_fold_hot through a worklist, where a full warp streams ONE texel's segment.
constexpr int kFoldChunk = 256;           // fragments per warp per pass through shared memory
constexpr int kFoldWarps = 4;             // warps per CTA
constexpr int kFoldPer = kFoldChunk / 32; // fragments per lane per chunk
constexpr uint32_t kFoldHot = 96;         // segment length above which a texel is "hot" (swept: 24..2048)

struct __align__(16) FoldTerm { float tx, ty, tz, tw; };   // src*a for the four channels

__device__ __forceinline__ uint32_t warp_min(uint32_t v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = min(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// the serial half of the blend for fragments [i, e) of a chunk held in shared memory
__device__ __forceinline__ void fold_terms(float4 &d, const FoldTerm *term, const float *om, uint32_t i, uint32_t e) {
    // software-pipelined by 8: the shared-memory reads do not depend on d, so only the multiply-add
    // chain (two dependent roundings per fragment and channel) is serial
    for (; i + 8 <= e; i += 8) {
        FoldTerm tm[8];
        float m[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) { tm[j] = term[i + j]; m[j] = om[i + j]; }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            d.x = __fadd_rn(tm[j].tx, __fmul_rn(d.x, m[j]));
            d.y = __fadd_rn(tm[j].ty, __fmul_rn(d.y, m[j]));
            d.z = __fadd_rn(tm[j].tz, __fmul_rn(d.z, m[j]));
            d.w = __fadd_rn(tm[j].tw, __fmul_rn(d.w, m[j]));
        }
    }
    for (; i < e; ++i) {
        const FoldTerm t1 = term[i];
        const float m1 = om[i];
        d.x = __fadd_rn(t1.tx, __fmul_rn(d.x, m1));
        d.y = __fadd_rn(t1.ty, __fmul_rn(d.y, m1));
        d.z = __fadd_rn(t1.tz, __fmul_rn(d.z, m1));
        d.w = __fadd_rn(t1.tw, __fmul_rn(d.w, m1));
    }
}

// the order-independent half, all lanes in parallel: src*a per channel and 1-a, into shared memory
__device__ __forceinline__ void stage_terms(FoldTerm *term, float *om, const float (&rcx)[kFoldPer], const float (&rcy)[kFoldPer],
                                            const float (&ra)[kFoldPer], uint32_t c0, uint32_t c1, int lane, float time) {
#pragma unroll
    for (int j = 0; j < kFoldPer; ++j) {
        if (c0 + j * 32 + lane < c1) {
            FoldTerm tm;
            tm.tx = __fmul_rn(rcx[j], ra[j]);
            tm.ty = __fmul_rn(rcy[j], ra[j]);
            tm.tz = __fmul_rn(time, ra[j]);
            tm.tw = __fmul_rn(ra[j], ra[j]);
            term[j * 32 + lane] = tm;
            om[j * 32 + lane] = __fsub_rn(1.0f, ra[j]);
        }
    }
}

// Where a fold launch reads and writes texels.  Single GPU: src = dst = the flow grid, the whole grid.
// Sharded ring (tb_splat_fold_ring): one launch per grid chunk; src is this rank's inbox (rank 0: its own
// grid), dst the NEXT rank's inbox mapped over NVLink (last rank: its own grid, dst2 = rank 0's grid), and
// because src != dst every texel of the chunk is written, touched or not.
struct FoldIO {
    const float4 *src;
    float4 *dst;
    float4 *dst2;
    int t_begin, t_end;
    int copy_all;
    // which 32-texel tiles this launch owns: tile_first + k*tile_stride (all of them by default; the band
    // fold of a sharded run deals the tiles round-robin to the ranks)
    int tile_first = 0, tile_stride = 1;
};
__device__ __forceinline__ int fold_tile(int tile_first, int tile_stride, int k) { return tile_first + tile_stride * k; }

__global__ void __launch_bounds__(kFoldWarps * 32) k_splat_fold(const FoldIO io, const uint2 *__restrict__ seg,
                                                                 const FragVal *__restrict__ vals, float time,
                                                                 uint32_t *__restrict__ hot_count, uint32_t *__restrict__ hot_list,
                                                                 uint32_t hot_threshold) {
    __shared__ FoldTerm s_term[kFoldWarps][kFoldChunk];
    __shared__ float s_om[kFoldWarps][kFoldChunk];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int t = io.t_begin + fold_tile(io.tile_first, io.tile_stride, blockIdx.x * kFoldWarps + warp) * 32 + lane;
    uint2 se = make_uint2(0u, 0u);
    if (t < io.t_end) se = seg[t];
    bool has = se.y > se.x;
    bool hot = false;
    if (has && se.y - se.x > hot_threshold) {           // hot texel: a whole warp will fold it
        hot_list[atomicAdd(hot_count, 1u)] = static_cast<uint32_t>(t);
        has = false;
        hot = true;
    }
    if (io.copy_all && t < io.t_end && !has && !hot) {  // untouched texel: carry it over to the next buffer
        const float4 v = io.src[t];
        io.dst[t] = v;
        if (io.dst2) io.dst2[t] = v;
    }
    // Segments are stored in texel order, so the warp's fragments lie in [lo, hi); fragments
    // overwritten by an opaque one (k_splat_bounds) and hot texels leave gaps between the lanes' ranges.
    const uint32_t lo = warp_min(has ? se.x : 0xffffffffu);
    if (lo == 0xffffffffu) return;                      // nothing to fold in these 32 texels
    const uint32_t hi = ~warp_min(has ? ~se.y : 0xffffffffu);
    // first chunk at or after `from` that some lane needs (chunks are aligned to lo)
    auto next_chunk = [&](uint32_t from) -> uint32_t {
        const uint32_t need = warp_min((has && se.y > from) ? max(se.x, from) : 0xffffffffu);
        return need == 0xffffffffu ? hi : lo + (need - lo) / kFoldChunk * kFoldChunk;
    };
    float rcx[kFoldPer], rcy[kFoldPer], ra[kFoldPer];
    auto prefetch = [&](uint32_t c0) {                  // global -> registers, coalesced over the warp
#pragma unroll
        for (int j = 0; j < kFoldPer; ++j) {
            const uint32_t i = c0 + j * 32 + lane;
            if (i < hi) { const FragVal f = load_frag(vals + i); rcx[j] = f.cx; rcy[j] = f.cy; ra[j] = f.a; }
        }
    };
    float4 d = has ? io.src[t] : make_float4(0.f, 0.f, 0.f, 0.f);
    FoldTerm *term = s_term[warp];
    float *om = s_om[warp];
    uint32_t c0 = next_chunk(lo);
    prefetch(c0);
    while (c0 < hi) {
        const uint32_t c1 = min(c0 + static_cast<uint32_t>(kFoldChunk), hi);
        stage_terms(term, om, rcx, rcy, ra, c0, c1, lane, time);
        __syncwarp();
        const uint32_t cn = next_chunk(c1);
        if (cn < hi) prefetch(cn);                      // in flight while this chunk is folded
        if (has) {
            const uint32_t b_abs = max(se.x, c0), e_abs = min(se.y, c1);
            if (b_abs < e_abs) fold_terms(d, term, om, b_abs - c0, e_abs - c0);
        }
        __syncwarp();
        c0 = cn;
    }
    if (has) {
        io.dst[t] = d;
        if (io.dst2) io.dst2[t] = d;
    }
}

// Long segments ("hot" texels: dense filaments, the centre of a ball spawn, a whole shard of a sharded run):
// a warp runs kHotChains blend chains side by side, one texel per chain, and takes the next texel from the
// worklist whenever a chain finishes (persistent, dynamically balanced).  Per pass every chain advances by up
// to kHotStep fragments: all 32 lanes pull the chains' next fragments into shared memory with cp.async
// (coalesced, no registers), pre-multiply the order-independent half in place (src*a, 1-a), then lanes
// 0..kHotChains-1 each run the serial half of their texel: two dependent roundings per fragment and channel.
// (Measured alternatives, profiles/r01_fold_variants.txt: one chain per warp is issue bound, 32 chains per warp
// with 32-fragment steps pays the staging latency too often on the longest segments.)
constexpr int kHotChains = 8;
constexpr int kHotStep = 128;
constexpr int kHotWarps = 2;

__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gmem_src) {
    const uint32_t dst = static_cast<uint32_t>(__cvta_generic_to_shared(smem_dst));
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async4(void *smem_dst, const void *gmem_src) {
    const uint32_t dst = static_cast<uint32_t>(__cvta_generic_to_shared(smem_dst));
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 0;" ::: "memory");
}
// [cp-async-end]  (the CPU tests swap the helpers above for plain copies)
// one fragment into the first three floats of a FoldTerm slot (tx = cx, ty = cy, tz = a)
__device__ __forceinline__ void stage_frag(FoldTerm *slot, const FragVal *f) {
#if TB_FRAG_BYTES == 16
    cp_async16(slot, f);
#else
    cp_async4(&slot->tx, &f->cx);
    cp_async4(&slot->ty, &f->cy);
    cp_async4(&slot->tz, &f->a);
#endif
}

__global__ void __launch_bounds__(kHotWarps * 32) k_splat_fold_hot(const FoldIO io, const uint2 *__restrict__ seg,
                                                                   const FragVal *__restrict__ vals, float time,
                                                                   const uint32_t *__restrict__ hot_count,
                                                                   const uint32_t *__restrict__ hot_list,
                                                                   uint32_t *__restrict__ cursor) {
    __shared__ __align__(16) FoldTerm s_term[kHotWarps][kHotChains][kHotStep];
    __shared__ float s_om[kHotWarps][kHotChains][kHotStep];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t n_hot = *hot_count;
    // chain state, held by lane k for chain k
    uint32_t pos = 0, end = 0, tex = 0;
    float4 d = make_float4(0.f, 0.f, 0.f, 0.f);
    bool active = false;
    auto grab = [&]() {                                  // next hot texel for this chain, if any
        active = false;
        const uint32_t w = atomicAdd(cursor, 1u);
        if (w < n_hot) {
            tex = hot_list[w];
            const uint2 se = seg[tex];
            pos = se.x; end = se.y;
            d = io.src[tex];
            active = true;
        }
    };
    if (lane < kHotChains) grab();
    while (__any_sync(0xffffffffu, active)) {
        // 1. pull the next fragments of every chain into shared memory
#pragma unroll
        for (int k = 0; k < kHotChains; ++k) {
            const uint32_t p = __shfl_sync(0xffffffffu, pos, k), e = __shfl_sync(0xffffffffu, end, k);
            const bool on = __shfl_sync(0xffffffffu, active ? 1 : 0, k) != 0;
            if (on) {
#pragma unroll
                for (int j = 0; j < kHotStep / 32; ++j) {
                    const uint32_t i = p + j * 32 + lane;
                    if (i < e) stage_frag(&s_term[warp][k][j * 32 + lane], vals + i);
                }
            }
        }
        cp_async_wait_all();
        __syncwarp();
        // 2. pre-multiply in place: (cx, cy, a, -) -> (cx*a, cy*a, time*a, a*a), 1-a
#pragma unroll
        for (int k = 0; k < kHotChains; ++k) {
            const uint32_t p = __shfl_sync(0xffffffffu, pos, k), e = __shfl_sync(0xffffffffu, end, k);
            const bool on = __shfl_sync(0xffffffffu, active ? 1 : 0, k) != 0;
            if (on) {
#pragma unroll
                for (int j = 0; j < kHotStep / 32; ++j) {
                    const int q = j * 32 + lane;
                    if (p + q < e) {
                        const FoldTerm raw = s_term[warp][k][q];          // raw fragment: tx = cx, ty = cy, tz = a
                        const float a = raw.tz;
                        FoldTerm tm;
                        tm.tx = __fmul_rn(raw.tx, a);
                        tm.ty = __fmul_rn(raw.ty, a);
                        tm.tz = __fmul_rn(time, a);
                        tm.tw = __fmul_rn(a, a);
                        s_term[warp][k][q] = tm;
                        s_om[warp][k][q] = __fsub_rn(1.0f, a);
                    }
                }
            }
        }
        __syncwarp();
        // 3. the serial half: lane k folds chain k
        if (lane < kHotChains && active) {
            const uint32_t n = min(static_cast<uint32_t>(kHotStep), end - pos);
            fold_terms(d, s_term[warp][lane], s_om[warp][lane], 0u, n);
            pos += n;
            if (pos == end) {
                io.dst[tex] = d;
                if (io.dst2) io.dst2[tex] = d;
                grab();
            }
        }
        __syncwarp();
    }
}

// first index of the sorted key array whose texel is >= each band's first texel (one thread per band edge)
__global__ void k_band_offsets(const uint32_t *__restrict__ keys, uint32_t n, int band_texels, int n_bands,
                               uint32_t *__restrict__ out) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b > n_bands) return;
    const uint64_t want = static_cast<uint64_t>(b) * static_cast<uint64_t>(band_texels);
    uint32_t lo = 0, hi = n;
    while (lo < hi) {
        const uint32_t mid = lo + (hi - lo) / 2;
        if (static_cast<uint64_t>(keys[mid] & ~kOpaqueBit) < want) lo = mid + 1; else hi = mid;
    }
    out[b] = lo;
}

// Gates of the sharded ring fold: flags carry the step number ("epoch") and live in memory the
// neighbouring rank maps over NVLink.  A kernel boundary orders the peer stores of the preceding fold
// launch before the signal; the fence makes them visible system-wide first.
__global__ void k_ring_wait(const uint32_t *flag, uint32_t epoch) {
    const volatile uint32_t *f = flag;
    while (*f < epoch) __nanosleep(200);
    __threadfence_system();
}
__global__ void k_ring_signal(uint32_t *peer_flag, uint32_t epoch) {
    __threadfence_system();
    *reinterpret_cast<volatile uint32_t *>(peer_flag) = epoch;
}

// Band fold of a sharded run (tb_splat_fold_bands): every rank maps every other rank's segment table, merged
// fragment buffer, offset table, flow grid and flags over NVLink (CUDA IPC).  The 32-texel tiles of the grid are
// dealt round-robin to the ranks.  Per step:
//   1. an owner reads the lengths of its texels' segments from every source rank's segment table (small),
//      scans them texel-major -- per texel the sources side by side in rank order = column order = draw order --
//      and writes each source the offsets its segments get in the owner's merged array;
//   2. every source PUSHES its sorted fragments to the owners of their texels (posted NVLink writes: measured
//      5-8x faster here than pulling the same bytes with remote loads, profiles/r01_multi_gpu.txt);
//   3. the local fold kernels run unchanged on the merged array, and the finished tiles are stored into every
//      rank's grid.
// All-rank barriers fence the phases: sorted / offsets known / fragments landed / grid complete.
constexpr int kMaxBandRanks = 16;
constexpr int kBandPhases = 4;
struct BandPeers {
    float4 *flow[kMaxBandRanks];
    uint32_t *flags[kMaxBandRanks];
    int n, me;
};
struct BandSources { const uint2 *seg[kMaxBandRanks]; };
struct BandSinks {
    FragVal *merged[kMaxBandRanks];
    uint32_t *dst[kMaxBandRanks];
};

// Lengths of the (local texel, source) segments, texel-major so that their exclusive scan lays a texel's
// sources side by side: warp = (tile k of mine, source j), lane = texel of the tile.
__global__ void __launch_bounds__(256) k_bands_lengths(const BandSources S, int n_src, int me, int mine, int G,
                                                        uint32_t *__restrict__ len) {
    const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (w >= mine * n_src) return;
    const int k = w / n_src, j = w - k * n_src;
    const int t = fold_tile(me, n_src, k) * 32 + lane;
    uint2 se = make_uint2(0u, 0u);
    if (t < G) se = S.seg[j][t];
    len[static_cast<size_t>(k * 32 + lane) * n_src + j] = se.y - se.x;
}

// off = exclusive scan of len (one extra element: the total).  Tell source j where its segment of each of my
// texels goes (its table dst, indexed by texel), and write the merged segment table of my texels.
__global__ void __launch_bounds__(256) k_bands_offsets(const BandSinks D, int n_src, int me, int mine, int G,
                                                        const uint32_t *__restrict__ off, uint32_t cap,
                                                        uint32_t *__restrict__ seg_m, int *__restrict__ overflow) {
    const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (w >= mine * n_src) return;
    const int k = w / n_src, j = w - k * n_src;
    const int t = fold_tile(me, n_src, k) * 32 + lane;
    if (t >= G) return;
    const size_t idx = static_cast<size_t>(k * 32 + lane) * n_src + j;
    const uint32_t o = off[idx];
    D.dst[j][t] = o;
    if (j == 0) {
        const uint32_t e = off[idx + n_src];
        seg_m[2 * t] = o;
        seg_m[2 * t + 1] = e <= cap ? e : o;              // overflow: fold nothing, the host raises
        if (e > cap) *overflow = 1;
    }
}

// Every source pushes its sorted fragments into the merged arrays of the owners of their texels.
__global__ void __launch_bounds__(256) k_bands_push(const uint32_t *__restrict__ keys, const FragVal *__restrict__ vals, uint32_t n,
                                                     const uint2 *__restrict__ seg, const uint32_t *__restrict__ dst,
                                                     const BandSinks D, int n_src, uint32_t cap) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t t = keys[i] & ~kOpaqueBit;
    const uint32_t b = seg[t].x;
    if (i < b) return;                                    // overwritten by a later opaque fragment of this source
    const uint32_t d = dst[t] + (i - b);
    if (d < cap) D.merged[(t >> 5) % static_cast<uint32_t>(n_src)][d] = vals[i];
}

// all-rank barrier over peer memory: thread j tells rank j "I reached `epoch`" and waits for rank j to say so
__global__ void k_bands_barrier(const BandPeers P, uint32_t *my_flags, int phase, uint32_t epoch) {
    const int j = threadIdx.x;
    if (j >= P.n) return;
    __threadfence_system();                               // everything this rank stored before the barrier
    *reinterpret_cast<volatile uint32_t *>(P.flags[j] + phase * kMaxBandRanks + P.me) = epoch;
    const volatile uint32_t *f = my_flags + phase * kMaxBandRanks + j;
    const long long t0 = clock64();
    while (*f < epoch) {
        __nanosleep(100);
        if (clock64() - t0 > (1ll << 37)) __trap();       // ~1 min: a rank died; fail instead of hanging the GPU
    }
    __threadfence_system();
}

// copy this rank's finished tiles into every other rank's grid
__global__ void __launch_bounds__(256) k_bands_publish(const float4 *__restrict__ flow, const BandPeers P, int G) {
    const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    const long long t = (static_cast<long long>(P.me) + static_cast<long long>(P.n) * w) * 32 + lane;
    if (t >= G) return;
    const float4 v = flow[t];
    for (int j = 0; j < P.n; ++j)
        if (j != P.me) P.flow[j][t] = v;
}

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
