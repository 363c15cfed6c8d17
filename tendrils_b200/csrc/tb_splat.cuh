// tb_splat.cuh -- the ordered flow splat (a7-a10) as a binned pipeline, sm_100a.
//
// The reference draws every particle as a GL_LINES segment prev -> cur into the flow FBO under
// blendFunc(SRC_ALPHA, ONE_MINUS_SRC_ALPHA), in primitive order p = x*PH + k
// (src/index.js:267-268,278-303, src/particles.js:147-158,182-186).  The blend is not
// commutative, so per texel the fragments must be applied in draw order.  This file does that
// without a global sort.  The grid is cut into STRIPS (16 x 8 texels at 1024^2); a strip is one BIN
// of the fragment array, or -- where the previous draw found it crowded -- 2, 4, ... 256 bins, one
// per range of its texels (the split map; any map gives the same result, it only balances the work):
//
//   k_splat_hist     per slab of consecutive primitives: fragments per bin                    (count)
//   k_splat_rows     per bin: exclusive scan over the slabs                                    (scan)
//   k_splat_plan     bin totals -> bin offsets, the fold work list, capacity check, and the
//                    split map of the NEXT draw                                                (scan)
//   k_splat_scatter  rasterise; every fragment goes straight to its slot of its bin, bins filled
//                    in draw order: the warps of a CTA take 32 consecutive primitives each and
//                    claim bin slots from shared-memory cursors one warp after the other (a token
//                    passed over named barriers), in slot order inside the warp                (emit)
//   k_splat_fold     one warp per bin: the bin's texels live in shared memory, the bin streams
//                    past (cp.async.bulk + mbarrier) in batches of 32 fragments, blended in order
//
// A fragment is written once (16 B) and read once; nothing is sorted, nothing returns to the host.
// Compiled with -fmad=false; see tb_math.cuh for the arithmetic contract.
#pragma once

#include "tb_kernels.cuh"

namespace tb {

// ------------------------------------------------------------------------------------------
// Geometry of the binning: strips of (1 << sxl) x (1 << syl) texels, at most kMaxStripTexels each.
// ------------------------------------------------------------------------------------------
constexpr int kMaxStrips = 8192;
constexpr int kMaxBins = 11264;           // bins per grid: k_splat_scatter keeps one cursor per bin in shared memory
constexpr int kMaxStripTexels = 512;      // texels of a strip = the most one warp of k_splat_fold holds (128 up to 1024^2 grids)
constexpr int kHistSegs = 16;             // the slabs are scanned in this many segments (k_splat_rows)
constexpr uint32_t kKeyLocalMask = 0x000fffffu;

struct StripGeom {
    int W, H;
    int sxl, syl;            // log2 of the strip width / height
    int strips_x, strips_y;
    int T;                   // strips_x * strips_y <= kMaxStrips
};

__host__ __device__ __forceinline__ int strip_of(const StripGeom &g, int gx, int gy) { return (gy >> g.syl) * g.strips_x + (gx >> g.sxl); }
__host__ __device__ __forceinline__ uint32_t local_of(const StripGeom &g, int gx, int gy) {
    return (static_cast<uint32_t>(gy & ((1 << g.syl) - 1)) << g.sxl) | static_cast<uint32_t>(gx & ((1 << g.sxl) - 1));
}

// The split map: map[strip] = first bin | log2(bins of the strip) << 24; bin_info[bin] = strip | sub << 16 | log2(bins) << 24.
// Sub-bin `sub` of a strip split 2^ls ways holds local texel indices [sub * (S >> ls), (sub + 1) * (S >> ls)).
struct BinMap {
    const uint32_t *__restrict__ map;        // [T]
    const uint32_t *__restrict__ n_bins;     // device scalar, <= kMaxBins
    int lS;                                  // log2 of the texels per strip
};
__device__ __forceinline__ uint32_t bin_of(const BinMap &M, uint32_t strip, uint32_t local) {
    const uint32_t m = __ldg(M.map + strip);
    return (m & 0xffffffu) + (local >> (M.lS - (m >> 24)));
}

// One fragment in a bin: the interpolated colour's (vel.xy, alpha) -- the time channel is the uniform
// `time` -- and where it goes inside its strip.
struct __align__(16) Frag {
    float cx, cy, a;
    uint32_t key;            // texel index inside the strip
};

// What the plan kernel leaves for the host (read lazily, never waited for on the hot path).
struct PlanOut {
    unsigned long long total;     // fragments of this draw (sharded run: of this rank)
    unsigned long long needed;    // fragments the bin array of this context has to hold
    uint32_t overflow;            // 1: they do not fit the bin array -- scatter and fold did nothing
    uint32_t n_items;
};

// ------------------------------------------------------------------------------------------
// Per-primitive setup shared by the count and the emit pass.
// ------------------------------------------------------------------------------------------
struct PrimGeom {
    float ma, mb, na, nb;    // major / minor window coordinates of the two vertices
    int c0;                  // closed form: first column (major axis) that yields a fragment
    uint32_t flags;          // bit 0: x is the major axis; bit 1: closed form applies
};

// [prim-begin]  (tests/test_raster_host.py compiles the text between these markers for the CPU)
// The rare lines that touch the minor-axis borders of the grid are enumerated; kept out of line so that the kernels'
// unrolled fragment rounds stay small (instruction cache).
__device__ __noinline__ uint32_t prim_count_slow(float xa, float ya, float xb, float yb, int W, int H) {
    uint32_t n = 0;
    raster_line(xa, ya, xb, yb, W, H, [&](int, int, float) { ++n; });
    return n;
}
__device__ __noinline__ void prim_fragment_slow(const PrimGeom &P, uint32_t j, int W, int H, int &gx, int &gy, float &t) {
    const bool xmajor = (P.flags & 1u) != 0u;
    uint32_t k = 0;
    gx = gy = 0; t = 0.0f;
    raster_line(xmajor ? P.ma : P.na, xmajor ? P.na : P.ma, xmajor ? P.mb : P.nb, xmajor ? P.nb : P.mb, W, H,
                [&](int x, int y, float tt) { if (k == j) { gx = x; gy = y; t = tt; } ++k; });
}
// Number of fragments of the line sa -> sb and what is needed to enumerate them.  Same decisions as
// count_fragments()/raster_line() (RASTER-1): when the line stays clear of the minor-axis borders, the
// fragments are exactly the member columns c0, c0+1, ..., c0+n-1 of the major axis.
__device__ __forceinline__ uint32_t prim_setup(const float4 &sa, const float4 &sb, float vsx, float vsy, int W, int H, PrimGeom &P) {
    P.flags = 0u; P.c0 = 0; P.ma = P.mb = P.na = P.nb = 0.0f;
    if (!splat_vertex_ok(sa) || !splat_vertex_ok(sb)) return 0u;
    const float hw = __fmul_rn(0.5f, static_cast<float>(W)), hh = __fmul_rn(0.5f, static_cast<float>(H));
    const float xa = __fadd_rn(__fmul_rn(__fmul_rn(sa.x, vsx), hw), hw), ya = __fadd_rn(__fmul_rn(__fmul_rn(sa.y, vsy), hh), hh);
    const float xb = __fadd_rn(__fmul_rn(__fmul_rn(sb.x, vsx), hw), hw), yb = __fadd_rn(__fmul_rn(__fmul_rn(sb.y, vsy), hh), hh);
    const float dx = __fsub_rn(xb, xa), dy = __fsub_rn(yb, ya);
    const bool xmajor = fabsf(dx) >= fabsf(dy);
    const float ma = xmajor ? xa : ya, mb = xmajor ? xb : yb, dm = xmajor ? dx : dy;
    const float na = xmajor ? ya : xa, nb = xmajor ? yb : xb;
    const int M = xmajor ? W : H, N = xmajor ? H : W;
    P.ma = ma; P.mb = mb; P.na = na; P.nb = nb;
    P.flags = xmajor ? 1u : 0u;
    if (!(fabsf(dm) > 0.0f)) return 0u;
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
        int c0 = static_cast<int>(flo);
        if (!member(flo)) { --n; ++c0; }
        if (fhi != flo && !member(fhi)) --n;
        P.c0 = c0;
        P.flags |= 2u;
        return n;
    }
    return prim_count_slow(xa, ya, xb, yb, W, H);
}

// Fragment j (0 <= j < n) of a primitive: its texel and the interpolation parameter.  The order of a line's
// fragments among themselves is immaterial (they hit distinct texels); closed form: ascending columns.
__device__ __forceinline__ void prim_fragment(const PrimGeom &P, uint32_t j, int W, int H, int &gx, int &gy, float &t) {
    const bool xmajor = (P.flags & 1u) != 0u;
    if (P.flags & 2u) {
        const int i = P.c0 + static_cast<int>(j);
        const float ic = __fadd_rn(static_cast<float>(i), 0.5f);
        t = __fdiv_rn(__fsub_rn(ic, P.ma), __fsub_rn(P.mb, P.ma));
        const float nn = __fadd_rn(P.na, __fmul_rn(t, __fsub_rn(P.nb, P.na)));
        const int jj = static_cast<int>(floorf(nn));
        gx = xmajor ? i : jj;
        gy = xmajor ? jj : i;
        return;
    }
    prim_fragment_slow(P, j, W, H, gx, gy, t);
}
// [prim-end]

// The two vertices of primitive (local column xl, active pair pi): src/state/state-at-frame.glsl:12-22 through the D6 table.
struct PrimSource {
    const float4 *__restrict__ cur;
    const float4 *__restrict__ prev;
    const PairEntry *__restrict__ pairs;
    int n_pairs, PH;
    long long n_prims;
};
__device__ __forceinline__ void load_prim(const PrimSource &S, long long p, float4 &sa, float4 &sb) {
    const uint32_t xl = static_cast<uint32_t>(p) / static_cast<uint32_t>(S.n_pairs);            // n_prims < 2^31
    const int4 pv = __ldg(reinterpret_cast<const int4 *>(S.pairs + (static_cast<uint32_t>(p) - xl * static_cast<uint32_t>(S.n_pairs))));
    const size_t base = static_cast<size_t>(xl) * S.PH;
    sa = __ldg(((pv.y < 0) ? S.cur : S.prev) + base + (pv.y & 0x7fffffff));
    sb = __ldg(((pv.z < 0) ? S.cur : S.prev) + base + (pv.z & 0x7fffffff));
}

__device__ __forceinline__ uint32_t warp_incl_scan(uint32_t v, int lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t u = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += u;
    }
    return v;
}

// ------------------------------------------------------------------------------------------
// Pass 0 (optional; sharded runs): opaque pruning.  A fragment with alpha == 1 overwrites its texel: dst = c*1 + dst*0.  For
// every texel, `last[texel]` = 1 + the (global) index of the LAST primitive that lays such a fragment on it; every fragment
// of an earlier primitive on that texel is dead -- it cannot influence the result -- and is neither counted, emitted nor
// folded.  Exactness (spec/PARITY.md B3): c + dst*0 = c needs a finite dst (guaranteed when every skipped fragment is
// finite with 0 <= a <= 1: the kernel raises `flags` bit 0 on a primitive that could break that, and the draw then prunes
// nothing) and a colour that is not -0 (such primitives do not mark).  Only primitives with alpha == 1 at BOTH vertices
// mark: then a = 1 + t*0 = 1 on every fragment.  Marking less than possible is always exact.
// ------------------------------------------------------------------------------------------
struct OpaqueArgs {
    PrimSource src;
    StripGeom g;
    float vsx, vsy, speedLimit, time;
    long long prim_base;                       // global index of this context's first primitive
    uint32_t *__restrict__ last;               // [W*H], zeroed beforehand
    uint32_t *flags;                           // bit 0: pruning would not be exact for this draw
};

__global__ void __launch_bounds__(256) k_splat_opaque(const OpaqueArgs A) {
    const long long stride = static_cast<long long>(gridDim.x) * 256;
    const bool time_ok = __float_as_uint(A.time) != 0x80000000u && A.speedLimit > 0.0f;
    for (long long pb = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x; pb < A.src.n_prims; pb += 4 * stride) {
        float4 sa[4], sb[4];
#pragma unroll
        for (int u = 0; u < 4; ++u)
            if (pb + u * stride < A.src.n_prims) load_prim(A.src, pb + u * stride, sa[u], sb[u]);
#pragma unroll 1
        for (int u = 0; u < 4; ++u) {
            const long long p = pb + u * stride;
            if (p >= A.src.n_prims) break;
            const float4 a = sa[u], b = sb[u];
            if (!splat_vertex_ok(a) || !splat_vertex_ok(b)) continue;
            const bool tame = fabsf(a.z) < 1.0e30f && fabsf(a.w) < 1.0e30f && fabsf(b.z) < 1.0e30f && fabsf(b.w) < 1.0e30f;
            if (!tame || !time_ok) { atomicOr(A.flags, 1u); continue; }
            const float aa = gmin(__fdiv_rn(glength(a.z, a.w), A.speedLimit), 1.0f);
            const float ab = gmin(__fdiv_rn(glength(b.z, b.w), A.speedLimit), 1.0f);
            if (!(aa == 1.0f && ab == 1.0f)) continue;
            if (__float_as_uint(a.z) == 0x80000000u || __float_as_uint(a.w) == 0x80000000u) continue;     // the colour could be -0
            PrimGeom P;
            const uint32_t n = prim_setup(a, b, A.vsx, A.vsy, A.g.W, A.g.H, P);
            const uint32_t mark = static_cast<uint32_t>(A.prim_base + p) + 1u;
            for (uint32_t j = 0; j < n; ++j) {
                int gx, gy; float t;
                prim_fragment(P, j, A.g.W, A.g.H, gx, gy, t);
                atomicMax(A.last + static_cast<size_t>(gy) * A.g.W + gx, mark);
            }
        }
    }
}

// is the fragment of (global) primitive p on texel (gx, gy) overwritten by a later primitive?
struct Prune {
    const uint32_t *__restrict__ last;         // null: no pruning
    const uint32_t *flags;
    long long prim_base;
};
__device__ __forceinline__ bool prune_on(const Prune &Q) { return Q.last != nullptr && (*Q.flags & 1u) == 0u; }
__device__ __forceinline__ bool fragment_dead(const uint32_t *__restrict__ last, uint32_t mark, int W, int gx, int gy) {
    return __ldg(last + static_cast<size_t>(gy) * W + gx) > mark;
}

// ------------------------------------------------------------------------------------------
// Pass 1: fragments per (slab, bin).  A slab is a fixed range of consecutive primitives; CTAs take
// slabs from a ticket counter.  slab_hist[slab * kMaxBins + bin]; seg_total[seg * kMaxBins + bin] accumulates
// the slabs of a segment (zeroed by k_splat_rows of the previous draw).
// ------------------------------------------------------------------------------------------
constexpr int kHistThreads = 512;

struct HistArgs {
    PrimSource src;
    StripGeom g;
    BinMap bm;
    Prune prune;
    float vsx, vsy;
    int slab_prims, n_slabs, slabs_per_seg;
    uint32_t *__restrict__ slab_hist;
    uint32_t *__restrict__ seg_total;
    uint32_t *ticket;
};

// Texel of fragment j of a closed-form primitive without the exact division where the answer cannot depend on it: the
// major coordinate is exact, the minor one is first estimated with a reciprocal; only an estimate within `eps` of a
// texel boundary is redone exactly.
__device__ __forceinline__ void prim_fragment_texel(const PrimGeom &P, uint32_t j, float inv_dm, float eps, int &gx, int &gy) {
    const int i = P.c0 + static_cast<int>(j);
    const float ic = __fadd_rn(static_cast<float>(i), 0.5f);
    const float dn = __fsub_rn(P.nb, P.na);
    const float est = __fadd_rn(P.na, __fmul_rn(__fmul_rn(__fsub_rn(ic, P.ma), inv_dm), dn));
    const float fl = floorf(est);
    int jj = static_cast<int>(fl);
    if (__fsub_rn(est, fl) < eps || __fsub_rn(__fadd_rn(fl, 1.0f), est) < eps) {
        const float t = __fdiv_rn(__fsub_rn(ic, P.ma), __fsub_rn(P.mb, P.ma));
        jj = static_cast<int>(floorf(__fadd_rn(P.na, __fmul_rn(t, dn))));
    }
    gx = (P.flags & 1u) ? i : jj;
    gy = (P.flags & 1u) ? jj : i;
}

constexpr int kHistWarps = kHistThreads / 32;
constexpr int kHistSlots = 128;                            // slots per pass of a warp (four rounds of 32)
constexpr size_t kHistSmemBytes = static_cast<size_t>(kMaxBins) * 4 + static_cast<size_t>(kHistWarps) * (8 * 32 + kHistSlots) * 4;

// Each warp takes 32 consecutive lines at a time, one per lane.  A line whose bounding box sits inside one unsplit strip adds its
// fragment count at once.  The fragments of the other lines are numbered by a warp scan and enumerated 32 at a time (like the
// emit pass, but in any order and without claims): all lanes busy whatever the lines' lengths.
__global__ void __launch_bounds__(kHistThreads) k_splat_hist(const HistArgs A) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint32_t *hist = reinterpret_cast<uint32_t *>(smem_raw);              // [kMaxBins]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    float *rec = reinterpret_cast<float *>(hist + kMaxBins) + warp * (8 * 32 + kHistSlots);    // [8][32]
    uint32_t *owner = reinterpret_cast<uint32_t *>(rec + 8 * 32);                             // [kHistSlots]
    __shared__ int s_slab;
    const int B = static_cast<int>(*A.bm.n_bins);
    const uint32_t *dead = prune_on(A.prune) ? A.prune.last : nullptr;
    for (int t = tid; t < B; t += kHistThreads) hist[t] = 0u;
    for (;;) {
        __syncthreads();
        if (tid == 0) s_slab = static_cast<int>(atomicAdd(A.ticket, 1u));
        __syncthreads();
        const int slab = s_slab;
        if (slab >= A.n_slabs) break;
        const long long p0 = static_cast<long long>(slab) * A.slab_prims;
        const long long p1 = (p0 + A.slab_prims < A.src.n_prims) ? p0 + A.slab_prims : A.src.n_prims;
        // the vertices of the warp's next 32 lines are loaded while these are worked on
        float4 sa_nx = make_float4(0.f, 0.f, 0.f, 0.f), sb_nx = sa_nx;
        if (p0 + tid < p1) load_prim(A.src, p0 + tid, sa_nx, sb_nx);
        for (long long w0 = p0; w0 < p1; w0 += kHistThreads) {
            const long long p = w0 + tid;
            const float4 sa = sa_nx, sb = sb_nx;
            if (p + kHistThreads < p1) load_prim(A.src, p + kHistThreads, sa_nx, sb_nx);
            uint32_t n = 0;                                            // fragments of this lane's line still to be enumerated
            if (p < p1) {
                PrimGeom P;
                n = prim_setup(sa, sb, A.vsx, A.vsy, A.g.W, A.g.H, P);
                const uint32_t mark = static_cast<uint32_t>(A.prune.prim_base + p) + 1u;
                if (n != 0u && !(P.flags & 2u)) {
                    // the rare line that touches the minor-axis borders: enumerated by its own lane
                    const bool xmajor = (P.flags & 1u) != 0u;
                    raster_line(xmajor ? P.ma : P.na, xmajor ? P.na : P.ma, xmajor ? P.mb : P.nb, xmajor ? P.nb : P.mb, A.g.W, A.g.H,
                                [&](int gx, int gy, float) {
                                    if (dead && fragment_dead(dead, mark, A.g.W, gx, gy)) return;
                                    atomicAdd(&hist[bin_of(A.bm, static_cast<uint32_t>(strip_of(A.g, gx, gy)), local_of(A.g, gx, gy))], 1u);
                                });
                    n = 0u;
                } else if (n != 0u) {
                    // closed form.  All fragments lie in the box spanned by the two vertices (the minor coordinate up to a few
                    // ulps): when that box, widened by 1/16 texel, sits in one unsplit strip, so do they.
                    const bool xmajor = (P.flags & 1u) != 0u;
                    const float mlo = gmin(P.ma, P.mb), mhi = gmax(P.ma, P.mb), nlo = __fsub_rn(gmin(P.na, P.nb), 0.0625f), nhi = __fadd_rn(gmax(P.na, P.nb), 0.0625f);
                    const int M = xmajor ? A.g.W : A.g.H, N = xmajor ? A.g.H : A.g.W;
                    int m0 = static_cast<int>(floorf(mlo)), m1 = static_cast<int>(floorf(mhi));
                    int q0 = static_cast<int>(floorf(nlo)), q1 = static_cast<int>(floorf(nhi));
                    m0 = m0 < 0 ? 0 : m0; m1 = m1 > M - 1 ? M - 1 : m1;
                    q0 = q0 < 0 ? 0 : q0; q1 = q1 > N - 1 ? N - 1 : q1;
                    const int s0 = xmajor ? strip_of(A.g, m0, q0) : strip_of(A.g, q0, m0);
                    const int s1 = xmajor ? strip_of(A.g, m1, q1) : strip_of(A.g, q1, m1);
                    const uint32_t m = __ldg(A.bm.map + s0);
                    if (s0 == s1 && (m >> 24) == 0u && !dead) {
                        atomicAdd(&hist[m & 0xffffffu], n);
                        n = 0u;
                    } else {
                        rec[0 * 32 + lane] = P.ma; rec[1 * 32 + lane] = P.mb;
                        rec[2 * 32 + lane] = P.na; rec[3 * 32 + lane] = P.nb;
                        rec[4 * 32 + lane] = __int_as_float(P.c0);
                        rec[5 * 32 + lane] = __uint_as_float(P.flags);
                        rec[6 * 32 + lane] = __fdiv_rn(1.0f, __fsub_rn(P.mb, P.ma));
                        rec[7 * 32 + lane] = __fmul_rn(__fadd_rn(__fadd_rn(fabsf(P.na), fabsf(__fsub_rn(P.nb, P.na))), 1.0f), 1.9073486328125e-06f);   // 2^-19
                    }
                }
            }
            // ---- the fragments still to be enumerated: slots by a warp scan, 32 per round
            const uint32_t inc = warp_incl_scan(n, lane);
            const uint32_t my_off = inc - n;
            const uint32_t n_warp = __shfl_sync(0xffffffffu, inc, 31);
            for (uint32_t s_lo = 0; s_lo < n_warp; s_lo += kHistSlots) {
                const uint32_t cnt = (n_warp - s_lo < static_cast<uint32_t>(kHistSlots)) ? n_warp - s_lo : static_cast<uint32_t>(kHistSlots);
                {
                    const uint32_t b = my_off > s_lo ? my_off : s_lo;
                    const uint32_t e = (my_off + n < s_lo + cnt) ? my_off + n : s_lo + cnt;
                    for (uint32_t s = b; s < e; ++s) owner[s - s_lo] = (static_cast<uint32_t>(lane) << 20) | (s - my_off);
                }
                __syncwarp();
                for (uint32_t s = lane; s < cnt; s += 32) {
                    const uint32_t o = owner[s];
                    const int q = static_cast<int>(o >> 20);
                    PrimGeom P;
                    P.ma = rec[0 * 32 + q]; P.mb = rec[1 * 32 + q];
                    P.na = rec[2 * 32 + q]; P.nb = rec[3 * 32 + q];
                    P.c0 = __float_as_int(rec[4 * 32 + q]);
                    P.flags = __float_as_uint(rec[5 * 32 + q]);
                    int gx, gy;
                    prim_fragment_texel(P, o & 0xfffffu, rec[6 * 32 + q], rec[7 * 32 + q], gx, gy);
                    if (dead && fragment_dead(dead, static_cast<uint32_t>(A.prune.prim_base + w0 + warp * 32 + q) + 1u, A.g.W, gx, gy)) continue;
                    atomicAdd(&hist[bin_of(A.bm, static_cast<uint32_t>(strip_of(A.g, gx, gy)), local_of(A.g, gx, gy))], 1u);
                }
                __syncwarp();
            }
        }
        __syncthreads();
        uint32_t *row = A.slab_hist + static_cast<size_t>(slab) * kMaxBins;
        uint32_t *seg = A.seg_total + static_cast<size_t>(slab / A.slabs_per_seg) * kMaxBins;
        for (int t = tid; t < B; t += kHistThreads) {
            const uint32_t h = hist[t];
            row[t] = h;
            if (h) { atomicAdd(&seg[t], h); hist[t] = 0u; }
        }
    }
}

// ------------------------------------------------------------------------------------------
// Pass 2a: per bin the exclusive scan of its slab counts (in place), segment by segment: thread (bin, segment)
// starts from the totals of the earlier segments.  Segment 0 also writes the bin's total.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_splat_rows(uint32_t *__restrict__ slab_hist, const uint32_t *__restrict__ seg_total,
                                                    uint32_t *__restrict__ seg_total_next, const uint32_t *__restrict__ n_bins, int n_slabs,
                                                    int slabs_per_seg, uint32_t *__restrict__ bin_total, uint32_t *__restrict__ too_many) {
    const int t = blockIdx.x * 256 + threadIdx.x;
    const int seg = blockIdx.y;
    seg_total_next[static_cast<size_t>(seg) * kMaxBins + t] = 0u;   // the next draw accumulates into the other buffer (grid covers kMaxBins)
    if (t >= static_cast<int>(*n_bins)) return;
    unsigned long long base = 0ull, all = 0ull;
#pragma unroll
    for (int s = 0; s < kHistSegs; ++s) {
        const uint32_t v = seg_total[static_cast<size_t>(s) * kMaxBins + t];
        if (s < seg) base += v;
        all += v;
    }
    if (seg == 0) {
        bin_total[t] = static_cast<uint32_t>(all);
        if (all > 0xffffffffull) *too_many = 1u;            // a single bin beyond 2^32 fragments: reported as overflow
    }
    const int s0 = seg * slabs_per_seg, s1 = (s0 + slabs_per_seg < n_slabs) ? s0 + slabs_per_seg : n_slabs;
    uint32_t run = static_cast<uint32_t>(base);
    if (s1 - s0 <= 64) {
        // the whole segment in flight at once: one DRAM round trip instead of one per batch
        uint32_t v[64];
#pragma unroll
        for (int k = 0; k < 64; ++k) v[k] = (s0 + k < s1) ? __ldcs(slab_hist + static_cast<size_t>(s0 + k) * kMaxBins + t) : 0u;
#pragma unroll
        for (int k = 0; k < 64; ++k) {
            if (s0 + k < s1) slab_hist[static_cast<size_t>(s0 + k) * kMaxBins + t] = run;
            run += v[k];
        }
        return;
    }
    for (int s = s0; s < s1; s += 8) {
        uint32_t v[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] = (s + k < s1) ? __ldcs(slab_hist + static_cast<size_t>(s + k) * kMaxBins + t) : 0u;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            if (s + k < s1) slab_hist[static_cast<size_t>(s + k) * kMaxBins + t] = run;
            run += v[k];
        }
    }
}

// ------------------------------------------------------------------------------------------
// Pass 2b (single GPU): bin offsets, the fold work list (longest bins first), the capacity check -- and the split map
// of the next draw, from this draw's fragments per strip.  One CTA.
// ------------------------------------------------------------------------------------------
constexpr int kPlanThreads = 1024;
constexpr int kPlanBins = kMaxBins / kPlanThreads;       // bins per thread
constexpr int kPlanStrips = kMaxStrips / kPlanThreads;   // strips per thread

// Segments (spec/PARITY.md B4).  A bin too long for one warp is folded by 2 .. 16 warps, a range of the bin's fragments
// each; k_splat_mend joins the ranges.  Per such bin: log2 of its segments, where its segment results and record counts live.
constexpr uint32_t kSegMaxLog = 4;
struct SegPlan {
    uint32_t seg_at;                           // fragments above which a (split) bin is folded in segments; 0: never
    uint32_t scaled;                           // 1: ... and above 1/1024 of the draw (a draw that large does not wait for one bin)
    uint32_t seg_len;                          // fragments per segment aimed at
    uint32_t out_cap;                          // float4 entries of the segment results
    uint4 *desc;                               // [kMaxBins] bin, log2(segments), first result entry, first count entry
    uint32_t *of_bin;                          // [kMaxBins] bin -> entry of desc
};
// fragments per segment: the bin cut into 2^lp ranges, whole bulk-copy windows each
__host__ __device__ __forceinline__ uint32_t seg_span(uint32_t n, uint32_t lp) { return ((n + (64u << lp) - 1u) >> (lp + 6u)) << 6u; }
// Is a bin of n fragments over R of the strip's S texels folded in segments, and in how many (log2)?  `s_seg`: shared counters
// [0] such bins, [1] result entries, [2] count entries.  Both bracketing chains of a texel live in the warp's texel array: R <= S/2.
__device__ __forceinline__ uint32_t plan_segments(uint32_t bin, uint32_t n, uint32_t R, uint32_t S, unsigned long long total, const SegPlan &G,
                                                  uint32_t *s_seg) {
    if (G.seg_at == 0u || 2u * R > S) return 0u;
    const unsigned long long at = (G.scaled && total >> 10 > G.seg_at) ? total >> 10 : G.seg_at;
    if (n <= at) return 0u;
    const unsigned long long len = total / (2ull * kMaxBins) > G.seg_len ? total / (2ull * kMaxBins) : G.seg_len;
    uint32_t lp = 1u;
    while (lp < kSegMaxLog && (n >> lp) > len) ++lp;
    const uint32_t slot = atomicAdd(&s_seg[1], R << lp);
    if (slot + (R << lp) > G.out_cap) return 0u;
    const uint32_t at_desc = atomicAdd(&s_seg[0], 1u), cnt = atomicAdd(&s_seg[2], 1u << lp);
    G.desc[at_desc] = make_uint4(bin, lp, slot, cnt);
    G.of_bin[bin] = at_desc;
    return lp;
}

struct PlanArgs {
    int T, lS;                                 // strips, log2 of the texels per strip
    BinMap bm;                                 // this draw's map
    const uint32_t *__restrict__ bin_info;     // [n_bins] strip | sub << 16 | log2(bins of the strip) << 24
    const uint32_t *__restrict__ bin_total;    // [n_bins]
    uint32_t *__restrict__ bin_off;            // [n_bins + 1]
    uint32_t *__restrict__ items;              // [16 kMaxBins] fold work items, longest first: bin | part << 16 | log2(parts) << 24 | segments << 31
    uint32_t cap;                              // capacity of the bin array (fragments)
    uint32_t split_at;                         // the most fragments a bin should hold (upper end of the target of plan_next_map)
    uint32_t share_at;                         // a bin with more fragments than this is folded by 2 warps, 2x: 4, 4x: 8
    SegPlan seg;                               // bins long enough to be folded in segments (k_splat_mend)
    uint32_t *too_many;                        // set by k_splat_rows; reset here
    uint32_t *tickets;                         // [0] hist (reset for the next draw), [1] scatter, [2] fold, [3] items, [5] mended bins, [6] mend
    uint32_t *map_next;                        // [T]   the next draw's map ...
    uint32_t *bin_info_next;                   // [kMaxBins]
    uint32_t *n_bins_next;                     // ... and its bin count
    PlanOut *out;
};

// block-wide exclusive scan of one value per thread (64-bit); returns the exclusive prefix, *total = sum
__device__ __forceinline__ unsigned long long block_excl_scan64(unsigned long long v, unsigned long long *warp_sums /* [32] */,
                                                               unsigned long long *total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    unsigned long long inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned long long u = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += u;
    }
    if (lane == 31) warp_sums[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        unsigned long long w = (lane < nw) ? warp_sums[lane] : 0ull;
        unsigned long long winc = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned long long u = __shfl_up_sync(0xffffffffu, winc, o);
            if (lane >= o) winc += u;
        }
        warp_sums[lane] = winc - w;
        if (lane == 31) *total = winc;
    }
    __syncthreads();
    const unsigned long long r = warp_sums[warp] + inc - v;
    __syncthreads();
    return r;
}

// The next draw's split map from this draw's fragments per bin: every strip gets as many bins (a power of two) as it takes to
// bring its bins under a target -- 1/4096 of the draw, between 256 and split_at fragments -- and the target is raised until the
// map fits kMaxBins.  Called by every thread of the (single) plan CTA.
__device__ __forceinline__ void plan_next_map(int T, int lS, const uint32_t *__restrict__ map, const uint32_t *__restrict__ bin_total,
                                              unsigned long long total, uint32_t split_at, uint32_t *map_next, uint32_t *bin_info_next, uint32_t *n_bins_next,
                                              unsigned long long *s_warp, unsigned long long *s_total) {
    const int u0 = threadIdx.x * kPlanStrips;
    uint32_t ns[kPlanStrips];
#pragma unroll
    for (int k = 0; k < kPlanStrips; ++k) {
        ns[k] = 0u;
        if (u0 + k < T) {
            const uint32_t m = map[u0 + k], first = m & 0xffffffu, cnt = 1u << (m >> 24);
            for (uint32_t b = 0; b < cnt; ++b) ns[k] += bin_total[first + b];
        }
    }
    // as many bins (a power of two, at most one per texel and at most 256: `sub` has 8 bits in bin_info) as it takes to bring a
    // strip's bins down to `at` fragments
    const uint32_t max_ls = lS < 8 ? static_cast<uint32_t>(lS) : 8u;
    auto want = [&](uint32_t frags, unsigned long long at) -> uint32_t {
        uint32_t ls = 0;
        while (ls < max_ls && (static_cast<unsigned long long>(frags) >> ls) > at) ++ls;
        return ls;
    };
    // A small draw (few fragments for this many SMs) splits earlier, so that the fold still has a few thousand bins to hand
    // out; then raise the threshold until the bins fit: the most crowded strips are the ones that stay split.
    unsigned long long at = total / 4096ull;
    at = at < 256ull ? 256ull : (at > split_at ? split_at : at);
    unsigned long long base = 0ull;
    for (int guard = 0; guard < 200; ++guard) {
        unsigned long long bins = 0ull;
#pragma unroll
        for (int k = 0; k < kPlanStrips; ++k)
            if (u0 + k < T) bins += 1ull << want(ns[k], at);
        base = block_excl_scan64(bins, s_warp, s_total);
        if (*s_total <= static_cast<unsigned long long>(kMaxBins)) break;
        at += (at >> 2) + 1ull;
    }
    const unsigned long long cap_ls = at;
    if (threadIdx.x == 0) *n_bins_next = static_cast<uint32_t>(*s_total);
#pragma unroll
    for (int k = 0; k < kPlanStrips; ++k) {
        if (u0 + k < T) {
            const uint32_t ls = want(ns[k], cap_ls);
            map_next[u0 + k] = static_cast<uint32_t>(base) | (ls << 24);
            for (uint32_t sub = 0; sub < (1u << ls); ++sub)
                bin_info_next[static_cast<uint32_t>(base) + sub] = static_cast<uint32_t>(u0 + k) | (sub << 16) | (ls << 24);
            base += 1ull << ls;
        }
    }
}

// log2 of the warps that share the fold of a bin of n fragments over R texels (every one streams the whole bin)
__device__ __forceinline__ uint32_t fold_lparts(uint32_t n, uint32_t R, uint32_t share_at) {
    uint32_t lp = 0;
    while (lp < 3u && (R >> (lp + 1)) >= 1u && (n >> lp) > share_at) ++lp;
    return lp;
}

__global__ void __launch_bounds__(kPlanThreads) k_splat_plan(const PlanArgs A) {
    __shared__ unsigned long long s_warp[32];
    __shared__ unsigned long long s_total;
    __shared__ uint32_t s_bucket[33];
    __shared__ uint32_t s_ok;
    __shared__ uint32_t s_seg[3];
    const int B = static_cast<int>(*A.bm.n_bins);
    const int t0 = threadIdx.x * kPlanBins;
    if (threadIdx.x < 3) s_seg[threadIdx.x] = 0u;
    uint32_t n[kPlanBins];
    unsigned long long mine = 0ull;
#pragma unroll
    for (int k = 0; k < kPlanBins; ++k) {
        n[k] = (t0 + k < B) ? A.bin_total[t0 + k] : 0u;
        mine += n[k];
    }
    const unsigned long long ex = block_excl_scan64(mine, s_warp, &s_total);
    const unsigned long long total = s_total;
    if (threadIdx.x == 0) {
        const bool ok = s_total <= static_cast<unsigned long long>(A.cap) && *A.too_many == 0u;
        s_ok = ok ? 1u : 0u;
        A.out->total = s_total;
        A.out->needed = s_total;
        A.out->overflow = ok ? 0u : 1u;
        A.tickets[0] = 0u; A.tickets[1] = 0u; A.tickets[2] = 0u;
        *A.too_many = 0u;
    }
    if (threadIdx.x < 33) s_bucket[threadIdx.x] = 0u;
    __syncthreads();
    const bool ok = s_ok != 0u;
    unsigned long long run = ex;
    const uint32_t share_at = static_cast<uint32_t>(total / 2048ull < 512ull ? 512ull : (total / 2048ull > A.share_at ? A.share_at : total / 2048ull));
    uint32_t rank[kPlanBins], lparts[kPlanBins];
    int bucket[kPlanBins];
#pragma unroll
    for (int k = 0; k < kPlanBins; ++k) {
        if (t0 + k < B) A.bin_off[t0 + k] = ok ? static_cast<uint32_t>(run) : 0u;
        run += n[k];
        // work list, longest bins first (bucketed by log2 of the length): the tail of the fold is short bins
        bucket[k] = __clz(n[k] | 1u);
        lparts[k] = 0u;
        if (ok && n[k]) {
            const uint32_t R = (1u << A.lS) >> (A.bin_info[t0 + k] >> 24);
            const uint32_t lseg = plan_segments(static_cast<uint32_t>(t0 + k), n[k], R, 1u << A.lS, total, A.seg, s_seg);
            lparts[k] = lseg ? (lseg | 0x80u) : fold_lparts(n[k], R, share_at);
        }
        rank[k] = (ok && n[k]) ? atomicAdd(&s_bucket[bucket[k]], 1u << (lparts[k] & 0x7fu)) : 0u;
    }
    if (threadIdx.x == 0) A.bin_off[B] = ok ? static_cast<uint32_t>(s_total) : 0u;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t r = 0;
        for (int b = 0; b < 32; ++b) { const uint32_t c = s_bucket[b]; s_bucket[b] = r; r += c; }
        A.out->n_items = r;
        A.tickets[3] = r;
        A.tickets[5] = s_seg[0];
        A.tickets[6] = 0u;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < kPlanBins; ++k)
        if (ok && n[k])
            for (uint32_t part = 0; part < (1u << (lparts[k] & 0x7fu)); ++part)
                A.items[s_bucket[bucket[k]] + rank[k] + part] = static_cast<uint32_t>(t0 + k) | (part << 16) | (lparts[k] << 24);

    plan_next_map(A.T, A.lS, A.bm.map, A.bin_total, total, A.split_at, A.map_next, A.bin_info_next, A.n_bins_next, s_warp, &s_total);
}

// the identity map (one bin per strip): first draw, and after a resize
__global__ void k_splat_map_identity(int T, uint32_t *map, uint32_t *bin_info, uint32_t *n_bins) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < T) { map[t] = static_cast<uint32_t>(t); bin_info[t] = static_cast<uint32_t>(t); }
    if (t == 0) *n_bins = static_cast<uint32_t>(T);
}

// ------------------------------------------------------------------------------------------
// Pass 3: emit.  Every fragment is computed once and stored once, at its final place: its bin, draw order.
//
// A CTA takes a slab (ticket) and walks it in windows of kEmitThreads primitives; warp w takes primitives
// [32w, 32w + 32) of the window, one per lane: load the two vertices, set the line up, count its fragments.  The warp's
// fragments are numbered in draw order by a warp scan ("slots") and handled in passes of at most kEmitSlots: the lanes
// expand their slot ranges into a table, then the lanes take the slots 32 at a time -- so that (round, lane) order is
// draw order -- compute the fragment and find the lanes of the round that hit the same bin (__match_any_sync).
// The slot in the bin comes from a shared-memory cursor per bin (atomicAdd by the group's first lane): the warps
// advance the cursors strictly one after the other, warp 0, 1, ... 7, 0, ... -- a token handed on over named barriers
// -- which is the draw order.  Only the claims are serial; computing and storing overlap between the warps.
// ------------------------------------------------------------------------------------------
constexpr int kEmitThreads = 256;
constexpr int kEmitWarps = kEmitThreads / 32;
constexpr int kEmitRounds = 5;                            // rounds of 32 slots per pass (a warp holds the token while it works
                                                          // on a second pass: the slots should cover 32 lines almost always)
constexpr int kEmitSlots = 32 * kEmitRounds;

struct ScatterArgs {
    PrimSource src;
    StripGeom g;
    BinMap bm;
    Prune prune;
    float vsx, vsy, speedLimit, time;
    int slab_prims, n_slabs;
    const uint32_t *__restrict__ slab_hist;     // scanned: fragments of this bin in earlier slabs
    const uint32_t *__restrict__ bin_off;       // [n_bins + 1] (sharded run: where this rank's fragments start in the owner's bin)
    const PlanOut *plan;
    uint32_t *ticket;
    Frag *bins[kMaxBandRanks];                  // the bin array of every rank (single GPU: [0]); bin b lives on rank b % n_ranks
    int n_ranks;
};

constexpr size_t kScatterSmemBytes = static_cast<size_t>(kMaxBins) * 4                                   // cursors
                                     + static_cast<size_t>(kEmitWarps) * (12 * 32 + kEmitSlots) * 4;   // per warp: primitive records (SoA), slot table

// The claim loop of k_splat_scatter relies on a converged warp issuing its (unrolled) shared-memory atomics in program order.
// In the SASS every round's ATOMS sits between BSSY.RECONVERGENT / BSYNC.RECONVERGENT, i.e. the warp reconverges before the
// next round's atomic (cuobjdump -sass; tests/test_host.py keeps an eye on it).  tests/test_pipeline_host.py runs the lanes as
// free OS threads and defines this as a warp barrier.
#ifndef TB_LOCKSTEP_FENCE
#define TB_LOCKSTEP_FENCE()
#endif

// [bar-begin]  (the CPU tests swap the helpers between these markers for host equivalents)
__device__ __forceinline__ void named_bar_sync(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }
__device__ __forceinline__ void named_bar_arrive(int id, int count) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory"); }
// [bar-end]

__global__ void __launch_bounds__(kEmitThreads, 3) k_splat_scatter(const ScatterArgs A) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    if (A.plan->overflow) return;
    const int B = static_cast<int>(*A.bm.n_bins);
    uint32_t *cur = reinterpret_cast<uint32_t *>(smem_raw);                    // [kMaxBins]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    float *rec = reinterpret_cast<float *>(cur + kMaxBins) + warp * (12 * 32 + kEmitSlots);    // [12][32]
    uint32_t *owner = reinterpret_cast<uint32_t *>(rec + 12 * 32);             // [kEmitSlots]
    __shared__ int s_slab;
    const uint32_t lt_mask = (1u << lane) - 1u;
    const uint32_t *dead = prune_on(A.prune) ? A.prune.last : nullptr;

    for (;;) {
        __syncthreads();
        if (tid == 0) s_slab = static_cast<int>(atomicAdd(A.ticket, 1u));
        __syncthreads();
        const int slab = s_slab;
        if (slab >= A.n_slabs) break;
        {
            const uint32_t *row = A.slab_hist + static_cast<size_t>(slab) * kMaxBins;
            for (int t = tid; t < B; t += kEmitThreads) cur[t] = A.bin_off[t] + row[t];
        }
        __syncthreads();
        const long long p0 = static_cast<long long>(slab) * A.slab_prims;
        const long long p1 = (p0 + A.slab_prims < A.src.n_prims) ? p0 + A.slab_prims : A.src.n_prims;
        // the vertices of the next window are loaded while this one is worked on
        float4 sa_nx = make_float4(0.f, 0.f, 0.f, 0.f), sb_nx = sa_nx;
        if (p0 + tid < p1) load_prim(A.src, p0 + tid, sa_nx, sb_nx);
        for (long long w0 = p0; w0 < p1; w0 += kEmitThreads) {
            const bool first_window = w0 == p0, last_window = w0 + kEmitThreads >= p1;
            // ---- one primitive per lane
            const long long p = w0 + tid;
            const float4 sa = sa_nx, sb = sb_nx;
            if (p + kEmitThreads < p1) load_prim(A.src, p + kEmitThreads, sa_nx, sb_nx);
            uint32_t n = 0;
            if (p < p1) {
                PrimGeom P;
                n = prim_setup(sa, sb, A.vsx, A.vsy, A.g.W, A.g.H, P);
                if (n) {
                    // flow(vel, speedLimit): src/flow/apply/state.glsl:5-16
                    const float aa = gmin(__fdiv_rn(glength(sa.z, sa.w), A.speedLimit), 1.0f);
                    const float ab = gmin(__fdiv_rn(glength(sb.z, sb.w), A.speedLimit), 1.0f);
                    rec[0 * 32 + lane] = P.ma; rec[1 * 32 + lane] = P.mb;
                    rec[2 * 32 + lane] = P.na; rec[3 * 32 + lane] = P.nb;
                    rec[4 * 32 + lane] = __int_as_float(P.c0);
                    rec[5 * 32 + lane] = __uint_as_float(P.flags);
                    rec[6 * 32 + lane] = sa.z; rec[7 * 32 + lane] = sb.z;
                    rec[8 * 32 + lane] = sa.w; rec[9 * 32 + lane] = sb.w;
                    rec[10 * 32 + lane] = aa; rec[11 * 32 + lane] = ab;
                }
            }
            // ---- slots: exclusive scan of the counts over the warp
            const uint32_t inc = warp_incl_scan(n, lane);
            const uint32_t my_off = inc - n;
            const uint32_t n_warp = __shfl_sync(0xffffffffu, inc, 31);
            bool have_token = false;
            // the token: warp w claims after warp w-1 of the same window, warp 0 after warp 7 of the previous window
            auto acquire = [&]() {
                if (warp != 0 || !first_window) named_bar_sync(1 + warp, 64);
                have_token = true;
            };
            auto release = [&]() {
                __threadfence_block();
                const int next = (warp + 1) % kEmitWarps;
                if (next != 0 || !last_window) named_bar_arrive(1 + next, 64);
            };
            for (uint32_t s_lo = 0; s_lo < n_warp; s_lo += kEmitSlots) {
                const uint32_t cnt = (n_warp - s_lo < static_cast<uint32_t>(kEmitSlots)) ? n_warp - s_lo : static_cast<uint32_t>(kEmitSlots);
                const bool last_pass = s_lo + kEmitSlots >= n_warp;
                // ---- the lanes expand their slots of this pass
                {
                    const uint32_t b = my_off > s_lo ? my_off : s_lo;
                    const uint32_t e = (my_off + n < s_lo + cnt) ? my_off + n : s_lo + cnt;
                    for (uint32_t s = b; s < e; ++s) owner[s - s_lo] = (static_cast<uint32_t>(lane) << 20) | (s - my_off);
                }
                __syncwarp();
                // ---- fragments, 32 slots per round
                float fcx[kEmitRounds], fcy[kEmitRounds], fa[kEmitRounds];
                uint32_t fbin[kEmitRounds], fpeers[kEmitRounds];                             // fbin: bin | texel index << 16
#pragma unroll
                for (int r = 0; r < kEmitRounds; ++r) {
                    fbin[r] = 0xffffffffu;
                    fpeers[r] = 0u;
                    if (static_cast<uint32_t>(r * 32) >= cnt) continue;                       // warp-uniform
                    const uint32_t s = r * 32 + lane;
                    if (s < cnt) {
                        const uint32_t o = owner[s];
                        const int q = static_cast<int>(o >> 20);
                        PrimGeom P;
                        P.ma = rec[0 * 32 + q]; P.mb = rec[1 * 32 + q];
                        P.na = rec[2 * 32 + q]; P.nb = rec[3 * 32 + q];
                        P.c0 = __float_as_int(rec[4 * 32 + q]);
                        P.flags = __float_as_uint(rec[5 * 32 + q]);
                        int gx, gy; float t;
                        prim_fragment(P, o & 0xfffffu, A.g.W, A.g.H, gx, gy, t);
                        const float za = rec[6 * 32 + q], zb = rec[7 * 32 + q];
                        const float wa = rec[8 * 32 + q], wb = rec[9 * 32 + q];
                        const float aa = rec[10 * 32 + q], ab = rec[11 * 32 + q];
                        fcx[r] = __fadd_rn(za, __fmul_rn(t, __fsub_rn(zb, za)));
                        fcy[r] = __fadd_rn(wa, __fmul_rn(t, __fsub_rn(wb, wa)));
                        fa[r] = __fadd_rn(aa, __fmul_rn(t, __fsub_rn(ab, aa)));
                        const uint32_t loc = local_of(A.g, gx, gy);
                        fbin[r] = bin_of(A.bm, static_cast<uint32_t>(strip_of(A.g, gx, gy)), loc) | (loc << 16);
                        // a fragment a later primitive overwrites takes no slot (k_splat_hist did not count it either)
                        if (dead && fragment_dead(dead, static_cast<uint32_t>(A.prune.prim_base + w0 + warp * 32 + q) + 1u, A.g.W, gx, gy)) fbin[r] = 0xffffffffu;
                    }
                    fpeers[r] = __match_any_sync(0xffffffffu, fbin[r] & 0xffffu);
                }
                // ---- claim the bin slots, in draw order.  This is the serial section of the CTA: the rounds' atomics are issued back
                // to back (one warp's shared-memory atomics execute in program order, which is the round order) and the token moves
                // on before the results are even looked at.
                if (!have_token) acquire();
                uint32_t fdst[kEmitRounds];
#pragma unroll
                for (int r = 0; r < kEmitRounds; ++r) {
                    fdst[r] = 0u;
                    if (static_cast<uint32_t>(r * 32) >= cnt) continue;                       // warp-uniform
                    if (fbin[r] != 0xffffffffu && lane == __ffs(fpeers[r]) - 1)
                        fdst[r] = atomicAdd(&cur[fbin[r] & 0xffffu], static_cast<uint32_t>(__popc(fpeers[r])));
                    TB_LOCKSTEP_FENCE();
                }
                if (last_pass) release();
                // ---- store
#pragma unroll
                for (int r = 0; r < kEmitRounds; ++r) {
                    if (static_cast<uint32_t>(r * 32) >= cnt) continue;                       // warp-uniform
                    fdst[r] = __shfl_sync(0xffffffffu, fdst[r], __ffs(fpeers[r]) - 1) + static_cast<uint32_t>(__popc(fpeers[r] & lt_mask));
                    if (fbin[r] == 0xffffffffu) continue;
                    Frag *bin = A.bins[A.n_ranks > 1 ? (fbin[r] & 0xffffu) % static_cast<uint32_t>(A.n_ranks) : 0u];
                    *reinterpret_cast<float4 *>(bin + fdst[r]) = make_float4(fcx[r], fcy[r], fa[r], __uint_as_float(fbin[r] >> 16));
                }
                __syncwarp();                                                   // the slot table is reused by the next pass
            }
            if (n_warp == 0u) { acquire(); release(); }
        }
    }
}

// ------------------------------------------------------------------------------------------
// Pass 4: fold.  blendFunc(SRC_ALPHA, ONE_MINUS_SRC_ALPHA) on all four channels in primitive order
// (src/index.js:267-268): dst = src*a + dst*(1-a), two roundings per channel.
//
// One warp per bin (work list, longest first, ticket counter).  The bin's texels (a strip, or 16 / 4 / 1 texels of a
// split strip) live in the warp's shared memory; the bin arrives through a double-buffered shared-memory window
// (cp.async.bulk + mbarrier, 1 KiB per copy) and is blended 32 fragments at a time: every lane forms the
// order-independent half of its fragment (src*a, 1-a), __match_any_sync finds the lanes that hit the same texel, and
// the batch is applied in as many rounds as the fullest texel has fragments: round 0 blends the first fragment of every
// texel onto the texel's value, round r blends the r-th onto what its predecessor lane produced (one shuffle), the last
// lane of every texel stores.  Lane order = draw order is kept per texel, different texels go in parallel.  A batch
// that piles up on few texels is chained by the first lane of each texel instead, operands staged in shared memory; there
// a fragment with alpha == 1, which overwrites the texel, lets the chain start at the last such fragment.  Two batches
// are in flight: the match of the second hides behind the rounds of the first.
// ------------------------------------------------------------------------------------------
constexpr int kFoldThreads = 256;
constexpr int kFoldNWarps = kFoldThreads / 32;
constexpr int kFoldStage = 64;                           // fragments per bulk copy
constexpr uint32_t kFoldRounds = 6;                      // batches whose fullest texel has more fragments are chained

struct FoldArgs {
    StripGeom g;
    float time;
    const Frag *__restrict__ bins;
    const uint32_t *__restrict__ bin_off;    // [n_bins + 1] where a bin starts (sharded run: the bins this rank owns, in its own array)
    const uint32_t *__restrict__ bin_count;  // null: bin b ends where b + 1 starts
    const uint32_t *__restrict__ bin_info;   // [n_bins] strip | sub << 16 | log2(bins of the strip) << 24
    const uint32_t *__restrict__ items;      // bins to fold
    const uint32_t *n_items;                 // device: number of work items
    uint32_t *ticket;
    float4 *flow[kMaxBandRanks];             // [0] the grid that is read; every entry below n_flow is written
    int n_flow;
    // bins folded in segments (plan_segments)
    const uint4 *__restrict__ seg_desc;      // bin, log2(segments), first result entry, first count entry
    const uint32_t *__restrict__ seg_of_bin;
    float4 *seg_out;                         // per segment the bin's texels: what the segment leaves behind (or kSegOpen / kSegRedo in .x)
    uint32_t *seg_cnt;                       // per segment: records in `replay`
    Frag *replay;                            // laid out like `bins`: a segment's records start where the segment starts
    const uint32_t *n_seg;                   // device: bins folded in segments
    uint32_t *seg_ticket;
};

constexpr int kFoldRing = 64;                            // compaction ring of a warp that shares a long bin
struct __align__(16) FoldWarp {             // shared memory of one warp, followed by the bin's texels: float4[texels per strip]
    Frag stage[2][kFoldStage];
    Frag ring[kFoldRing];
    float4 term[32];                         // chained batches: src*a per lane
    float om[32];                            //                  1 - a per lane
    unsigned long long bar[2];
};
__host__ __device__ inline size_t fold_warp_bytes(int strip_texels) { return sizeof(FoldWarp) + static_cast<size_t>(strip_texels) * sizeof(float4); }
__host__ __device__ inline size_t fold_smem_bytes(int strip_texels) { return fold_warp_bytes(strip_texels) * kFoldNWarps; }

// [bulk-begin]  (the CPU tests swap the helpers between these markers for plain copies)
__device__ __forceinline__ uint32_t smem_addr(const void *p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
// one thread: expect `bytes` on the barrier and start the bulk copy global -> shared that completes it
__device__ __forceinline__ void bulk_load(void *smem_dst, const void *gmem_src, uint32_t bytes, unsigned long long *bar) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_addr(smem_dst)), "l"(gmem_src), "r"(bytes), "r"(smem_addr(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_addr(bar)), "r"(parity) : "memory");
}
// [bulk-end]

// The order-independent half of one batch: lane l holds fragment l (in draw order) if `active`.
struct FoldPrep {
    uint32_t rel, peers, most, cut;          // cut: lanes whose fragment overwrites the texel and may start a chain
    float tx, ty, tz, tw, om;
};
__device__ __forceinline__ FoldPrep fold_prep(const Frag &f, bool active, uint32_t lo, float time, int lane) {
    FoldPrep p;
    p.rel = active ? (f.key & kKeyLocalMask) - lo : (0xffffff00u | static_cast<uint32_t>(lane));
    const float a = f.a;
    p.tx = __fmul_rn(f.cx, a); p.ty = __fmul_rn(f.cy, a); p.tz = __fmul_rn(time, a); p.tw = __fmul_rn(a, a);
    p.om = __fsub_rn(1.0f, a);
    p.peers = __match_any_sync(0xffffffffu, p.rel);
    p.most = __reduce_max_sync(0xffffffffu, static_cast<uint32_t>(__popc(p.peers)));
    p.cut = 0u;
    if (p.most > kFoldRounds) {
        // alpha == 1: dst = c*1 + dst*0 = c + (+-0) = c for every finite dst unless c is -0, and NaN for a non-finite dst
        // whatever was blended before (a finite, in-range fragment never makes a non-finite texel finite or a finite one
        // non-finite): starting the chain at such a fragment, on the texel's value as it was, is exact provided every
        // skipped fragment is finite with 0 <= a <= 1.
        const bool tame = is_finite(f.cx) && is_finite(f.cy) && a >= 0.0f && a <= 1.0f;
        const bool opaque = tame && a == 1.0f && __float_as_uint(f.cx) != 0x80000000u && __float_as_uint(f.cy) != 0x80000000u &&
                            __float_as_uint(time) != 0x80000000u;
        const uint32_t opq = __ballot_sync(0xffffffffu, active && opaque) & p.peers;
        const uint32_t wild = __ballot_sync(0xffffffffu, active && !tame) & p.peers;
        if (opq) {
            const uint32_t below = (1u << (31 - __clz(opq))) - 1u;      // lanes before the group's last opaque fragment
            if ((wild & below) == 0u) p.cut = below;
        }
    }
    return p;
}

__device__ __forceinline__ void fold_apply(FoldWarp &W, float4 *tex, const FoldPrep &p, bool active, int lane) {
    const uint32_t before = p.peers & ((1u << lane) - 1u);
    const uint32_t rank = static_cast<uint32_t>(__popc(before));
    if (p.most <= kFoldRounds) {
        float4 d = make_float4(0.f, 0.f, 0.f, 0.f);
        if (active) d = tex[p.rel];
        d.x = __fadd_rn(p.tx, __fmul_rn(d.x, p.om)); d.y = __fadd_rn(p.ty, __fmul_rn(d.y, p.om));
        d.z = __fadd_rn(p.tz, __fmul_rn(d.z, p.om)); d.w = __fadd_rn(p.tw, __fmul_rn(d.w, p.om));
        const int pred = before ? 31 - __clz(before) : lane;            // the lane with the previous fragment of my texel
        for (uint32_t r = 1; r < p.most; ++r) {
            const float px = __shfl_sync(0xffffffffu, d.x, pred), py = __shfl_sync(0xffffffffu, d.y, pred);
            const float pz = __shfl_sync(0xffffffffu, d.z, pred), pw = __shfl_sync(0xffffffffu, d.w, pred);
            if (rank == r) {
                d.x = __fadd_rn(p.tx, __fmul_rn(px, p.om)); d.y = __fadd_rn(p.ty, __fmul_rn(py, p.om));
                d.z = __fadd_rn(p.tz, __fmul_rn(pz, p.om)); d.w = __fadd_rn(p.tw, __fmul_rn(pw, p.om));
            }
        }
        if (active && (p.peers >> lane) == 1u) tex[p.rel] = d;        // the last fragment of the texel in this batch
    } else {
        W.term[lane] = make_float4(p.tx, p.ty, p.tz, p.tw);
        W.om[lane] = p.om;
        __syncwarp();
        // The chain of a crowded texel is serial: two dependent roundings per fragment and channel.  The first FOUR lanes of the
        // texel's group run it, one colour channel each (a group of fewer than four fragments: its first lane, all channels),
        // over the group's fragments in lane order, operands read from shared memory four steps ahead of the chain.
        const uint32_t members = p.peers & ~p.cut;
        const bool wide = __popc(p.peers) >= 4;
        if (active && wide && rank < 4u) {
            const float *term = reinterpret_cast<const float *>(W.term) + rank;          // channel `rank` of lane j at term[4 j]
            float *slot = reinterpret_cast<float *>(tex + p.rel) + rank;
            float d = *slot;
            uint32_t rem = members;
            while (rem) {
                int j[4];
                bool on[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    on[k] = rem != 0u;
                    j[k] = on[k] ? __ffs(rem) - 1 : 0;
                    rem &= rem - 1u;
                }
                float v[4], m[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) { v[k] = term[4 * j[k]]; m[k] = W.om[j[k]]; }
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    if (on[k]) d = __fadd_rn(v[k], __fmul_rn(d, m[k]));
            }
            *slot = d;
        } else if (active && !wide && rank == 0u) {
            float4 d = tex[p.rel];
            uint32_t rem = members;
            while (rem) {
                const int j0 = __ffs(rem) - 1;
                rem &= rem - 1u;
                const float4 v0 = W.term[j0];
                const float m0 = W.om[j0];
                d.x = __fadd_rn(v0.x, __fmul_rn(d.x, m0)); d.y = __fadd_rn(v0.y, __fmul_rn(d.y, m0));
                d.z = __fadd_rn(v0.z, __fmul_rn(d.z, m0)); d.w = __fadd_rn(v0.w, __fmul_rn(d.w, m0));
            }
            tex[p.rel] = d;
        }
    }
    __syncwarp();
}

// ---- segments (spec/PARITY.md B4) -----------------------------------------------------------------------------------
// One step of a texel's chain, x -> fl(c + fl(x * m)), is monotone in x for finite c, m (each rounding is), so a run of steps
// is.  Segment j >= 1 of a long bin does not know what the earlier segments leave in its texels, so it runs every texel's
// chain twice, from -FLT_MAX and from +FLT_MAX.  Once the two agree bit for bit (and are finite) the texel has SETTLED: every
// finite start value lies between them, so the true chain has the same value from there on, whatever came before.  Hot texels
// settle after a few hundred fragments (the start's weight shrinks by 1 - a per fragment); until they do -- and for texels
// that never do (few fragments, tiny alphas) -- the segment RECORDS the texel's fragments, in order, and k_splat_mend
// replays the records onto the true value.  Nothing is approximated: a settled value is the value the sequential blend has.
constexpr uint32_t kSegOpen = 0x7fc00001u;               // result .x of a texel that did not settle: replay its records
constexpr uint32_t kSegRedo = 0x7fc00002u;               // settled, but not finite in the end: fold the bin again, serially
constexpr float kSegLo = -3.402823466e+38f, kSegHi = 3.402823466e+38f;

__device__ __forceinline__ bool finite4(const float4 &v) { return is_finite(v.x) && is_finite(v.y) && is_finite(v.z) && is_finite(v.w); }
__device__ __forceinline__ bool same_bits4(const float4 &a, const float4 &b) {
    return __float_as_uint(a.x) == __float_as_uint(b.x) && __float_as_uint(a.y) == __float_as_uint(b.y) &&
           __float_as_uint(a.z) == __float_as_uint(b.z) && __float_as_uint(a.w) == __float_as_uint(b.w);
}
__device__ __forceinline__ bool seg_bit(const uint32_t *mask, uint32_t i) { return ((mask[i >> 5] >> (i & 31u)) & 1u) != 0u; }

__device__ __forceinline__ void seg_begin(float4 *lo_tex, float4 *hi_tex, uint32_t *settled, uint32_t R, int lane) {
    for (uint32_t l = lane; l < R; l += 32) {
        lo_tex[l] = make_float4(kSegLo, kSegLo, kSegLo, kSegLo);
        hi_tex[l] = make_float4(kSegHi, kSegHi, kSegHi, kSegHi);
    }
    for (uint32_t l = lane; l < (R + 31u) / 32u; l += 32) settled[l] = 0u;
    __syncwarp();
}
// one batch of a segment: record the fragments of texels that have not settled, blend onto both chains, see who settles
__device__ __forceinline__ void seg_batch(FoldWarp &W, float4 *lo_tex, float4 *hi_tex, uint32_t *settled, const Frag &f, bool active, uint32_t lo,
                                          float time, int lane, Frag *rec, uint32_t &n_rec) {
    const FoldPrep p = fold_prep(f, active, lo, time, lane);
    const bool open = active && !seg_bit(settled, p.rel);
    const uint32_t ob = __ballot_sync(0xffffffffu, open);
    if (open) rec[n_rec + static_cast<uint32_t>(__popc(ob & ((1u << lane) - 1u)))] = f;
    n_rec += static_cast<uint32_t>(__popc(ob));
    fold_apply(W, lo_tex, p, active, lane);
    fold_apply(W, hi_tex, p, active, lane);
    if (open && (p.peers >> lane) == 1u) {                            // the texel's last fragment in this batch
        const float4 l = lo_tex[p.rel], h = hi_tex[p.rel];
        if (same_bits4(l, h) && finite4(l)) atomicOr(&settled[p.rel >> 5], 1u << (p.rel & 31u));
    }
    __syncwarp();
}
// what the segment leaves behind, per texel: the settled value, or "replay my records", or "cannot tell"
__device__ __forceinline__ void seg_end(const float4 *lo_tex, const float4 *hi_tex, const uint32_t *settled, uint32_t R, float4 *out, int lane) {
    for (uint32_t l = lane; l < R; l += 32) {
        float4 e = lo_tex[l];
        if (!seg_bit(settled, l)) e.x = __uint_as_float(kSegOpen);
        else if (!(same_bits4(e, hi_tex[l]) && finite4(e))) e.x = __uint_as_float(kSegRedo);
        out[l] = e;
    }
}
// k_splat_mend, before segment j's records are replayed onto the true values T: which texels the segment settles (mask).
// True: the bin has to be folded again serially (a settled value holds for a finite start only).
__device__ __forceinline__ bool mend_begin(const float4 *T, const float4 *E, uint32_t R, uint32_t *mask, int lane) {
    bool redo = false;
    for (uint32_t l0 = 0; l0 < R; l0 += 32) {
        const uint32_t l = l0 + static_cast<uint32_t>(lane);
        bool settles = false;
        if (l < R) {
            const uint32_t ex = __float_as_uint(E[l].x);
            if (ex == kSegRedo) redo = true;
            else if (ex != kSegOpen) { settles = true; if (!finite4(T[l])) redo = true; }
        }
        const uint32_t m = __ballot_sync(0xffffffffu, settles);
        if (lane == 0) mask[l0 >> 5] = m;
    }
    __syncwarp();
    return __any_sync(0xffffffffu, redo);
}
__device__ __forceinline__ void mend_batch(FoldWarp &W, float4 *T, const uint32_t *mask, const Frag &f, bool active, uint32_t lo, float time, int lane) {
    const bool act = active && !seg_bit(mask, (f.key & kKeyLocalMask) - lo);           // records made before the texel settled: moot
    fold_apply(W, T, fold_prep(f, act, lo, time, lane), act, lane);
}
__device__ __forceinline__ void mend_end(float4 *T, const float4 *E, uint32_t R, const uint32_t *mask, int lane) {
    for (uint32_t l = lane; l < R; l += 32)
        if (seg_bit(mask, l)) T[l] = E[l];
    __syncwarp();
}

// [fold-host-end]  (tests/test_fold_host.py runs everything from fold_prep to here on an emulated warp)

// the warp's window over fragments [src, src + n): batch(fragment of this lane, active) for every 32 of them, in order
template <class Batch>
__device__ __forceinline__ void fold_stream(FoldWarp &W, const Frag *src, uint32_t n, int lane, uint32_t &phase0, uint32_t &phase1, Batch &&batch) {
    const uint32_t n_stage = (n + kFoldStage - 1) / kFoldStage;
    auto issue = [&](uint32_t j) {
        const uint32_t c = (n - j * kFoldStage < static_cast<uint32_t>(kFoldStage)) ? n - j * kFoldStage : static_cast<uint32_t>(kFoldStage);
        bulk_load(W.stage[j & 1u], src + static_cast<size_t>(j) * kFoldStage, c * static_cast<uint32_t>(sizeof(Frag)), &W.bar[j & 1u]);
    };
    if (lane == 0) {
        if (n_stage > 0) issue(0);
        if (n_stage > 1) issue(1);
    }
    for (uint32_t j = 0; j < n_stage; ++j) {
        const uint32_t s = j & 1u;
        const uint32_t c = (n - j * kFoldStage < static_cast<uint32_t>(kFoldStage)) ? n - j * kFoldStage : static_cast<uint32_t>(kFoldStage);
        if (s == 0u) { mbar_wait(&W.bar[0], phase0); phase0 ^= 1u; } else { mbar_wait(&W.bar[1], phase1); phase1 ^= 1u; }
        Frag f0{0.f, 0.f, 0.f, 0u}, f1{0.f, 0.f, 0.f, 0u};
        const bool a0 = static_cast<uint32_t>(lane) < c, a1 = static_cast<uint32_t>(32 + lane) < c;
        if (a0) f0 = W.stage[s][lane];
        if (a1) f1 = W.stage[s][32 + lane];
        batch(f0, a0);
        if (c > 32u) batch(f1, a1);
        __syncwarp();
        if (lane == 0 && j + 2 < n_stage) issue(j + 2);
    }
}

__device__ __forceinline__ void fold_load_texels(const FoldArgs &A, float4 *tex, uint32_t lo, uint32_t R, int gx0, int gy0, int lane) {
    for (uint32_t l = lane; l < R; l += 32) {
        const uint32_t loc = lo + l;
        const int gx = gx0 + static_cast<int>(loc & ((1u << A.g.sxl) - 1u)), gy = gy0 + static_cast<int>(loc >> A.g.sxl);
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (gx < A.g.W && gy < A.g.H) v = A.flow[0][static_cast<size_t>(gy) * A.g.W + gx];
        tex[l] = v;
    }
    __syncwarp();
}
__device__ __forceinline__ void fold_store_texels(const FoldArgs &A, const float4 *tex, uint32_t lo, uint32_t R, int gx0, int gy0, int lane) {
    for (uint32_t l = lane; l < R; l += 32) {
        const uint32_t loc = lo + l;
        const int gx = gx0 + static_cast<int>(loc & ((1u << A.g.sxl) - 1u)), gy = gy0 + static_cast<int>(loc >> A.g.sxl);
        if (gx < A.g.W && gy < A.g.H) {
            const float4 v = tex[l];
            const size_t at = static_cast<size_t>(gy) * A.g.W + gx;
            for (int r = 0; r < A.n_flow; ++r) A.flow[r][at] = v;
        }
    }
    __syncwarp();
}

// One segment of a long bin (work item with bit 31): fragments [part * span, + span) of the bin, all of its texels.
__device__ __forceinline__ void fold_segment(const FoldArgs &A, FoldWarp &W, float4 *tex, uint32_t bin_id, uint32_t part, uint32_t lp, uint32_t st,
                                             uint32_t lo, uint32_t R, int lane, uint32_t &phase0, uint32_t &phase1) {
    const uint32_t begin = A.bin_off[bin_id], n = A.bin_count ? A.bin_count[bin_id] : A.bin_off[bin_id + 1] - begin;
    const uint4 d = A.seg_desc[A.seg_of_bin[bin_id]];
    const uint32_t span = seg_span(n, lp);
    const uint32_t b0 = part * span < n ? part * span : n, b1 = b0 + span < n ? b0 + span : n;
    float4 *out = A.seg_out + d.z + static_cast<size_t>(part) * R;
    if (part == 0u) {                                      // the first segment starts from the grid: its result is the truth so far
        const int gx0 = (static_cast<int>(st) % A.g.strips_x) << A.g.sxl, gy0 = (static_cast<int>(st) / A.g.strips_x) << A.g.syl;
        fold_load_texels(A, tex, lo, R, gx0, gy0, lane);
        fold_stream(W, A.bins + begin, b1, lane, phase0, phase1,
                    [&](const Frag &f, bool act) { fold_apply(W, tex, fold_prep(f, act, lo, A.time, lane), act, lane); });
        for (uint32_t l = lane; l < R; l += 32) out[l] = tex[l];
    } else {
        float4 *hi_tex = tex + R;
        uint32_t *settled = reinterpret_cast<uint32_t *>(W.ring);
        seg_begin(tex, hi_tex, settled, R, lane);
        uint32_t n_rec = 0;
        Frag *rec = A.replay + begin + b0;
        fold_stream(W, A.bins + begin + b0, b1 - b0, lane, phase0, phase1,
                    [&](const Frag &f, bool act) { seg_batch(W, tex, hi_tex, settled, f, act, lo, A.time, lane, rec, n_rec); });
        seg_end(tex, hi_tex, settled, R, out, lane);
        if (lane == 0) A.seg_cnt[d.w + part] = n_rec;
    }
    __syncwarp();
}

__global__ void __launch_bounds__(kFoldThreads, 4) k_splat_fold(const FoldArgs A) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t S = 1u << (A.g.sxl + A.g.syl);
    FoldWarp &W = *reinterpret_cast<FoldWarp *>(smem_raw + fold_warp_bytes(static_cast<int>(S)) * warp);
    float4 *tex = reinterpret_cast<float4 *>(&W + 1);
    const uint32_t n_items = *A.n_items;
    if (lane == 0) {
        mbar_init(&W.bar[0], 1u);
        mbar_init(&W.bar[1], 1u);
        mbar_fence_init();
    }
    __syncwarp();
    uint32_t phase0 = 0u, phase1 = 0u;
    for (;;) {
        uint32_t item = 0;
        if (lane == 0) item = atomicAdd(A.ticket, 1u);
        item = __shfl_sync(0xffffffffu, item, 0);
        if (item >= n_items) break;
        const uint32_t icode = A.items[item];                        // bin | part << 16 | log2(parts) << 24
        const uint32_t bin_id = icode & 0xffffu, part = (icode >> 16) & 0xffu, lparts = (icode >> 24) & 0x7fu;
        const uint32_t code = A.bin_info[bin_id];
        const uint32_t st = code & 0xffffu, sub = (code >> 16) & 0xffu, ls = code >> 24;
        if (icode >> 31) {
            fold_segment(A, W, tex, bin_id, part, lparts, st, sub * (S >> ls), S >> ls, lane, phase0, phase1);
            continue;
        }
        // the bin's texels are local indices [sub * (S >> ls), + S >> ls); a long bin is shared by 2^lparts warps, each of which
        // streams the whole bin and keeps the fragments of its part of the texels
        const uint32_t R = (S >> ls) >> lparts, lo = sub * (S >> ls) + part * R;
        const bool whole = lparts == 0u;
        const uint32_t begin = A.bin_off[bin_id], n = A.bin_count ? A.bin_count[bin_id] : A.bin_off[bin_id + 1] - begin;
        const uint32_t n_stage = (n + kFoldStage - 1) / kFoldStage;
        const Frag *bin = A.bins + begin;
        auto issue = [&](uint32_t j) {
            const uint32_t c = (n - j * kFoldStage < static_cast<uint32_t>(kFoldStage)) ? n - j * kFoldStage : static_cast<uint32_t>(kFoldStage);
            bulk_load(W.stage[j & 1u], bin + static_cast<size_t>(j) * kFoldStage, c * static_cast<uint32_t>(sizeof(Frag)), &W.bar[j & 1u]);
        };
        if (lane == 0) {
            if (n_stage > 0) issue(0);
            if (n_stage > 1) issue(1);
        }
        const int gx0 = (static_cast<int>(st) % A.g.strips_x) << A.g.sxl, gy0 = (static_cast<int>(st) / A.g.strips_x) << A.g.syl;
        for (uint32_t l = lane; l < R; l += 32) {
            const uint32_t loc = lo + l;
            const int gx = gx0 + static_cast<int>(loc & ((1u << A.g.sxl) - 1u)), gy = gy0 + static_cast<int>(loc >> A.g.sxl);
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (gx < A.g.W && gy < A.g.H) v = A.flow[0][static_cast<size_t>(gy) * A.g.W + gx];
            tex[l] = v;
        }
        __syncwarp();
        uint32_t head = 0, queued = 0;
        for (uint32_t j = 0; j < n_stage; ++j) {
            const uint32_t s = j & 1u;
            const uint32_t c = (n - j * kFoldStage < static_cast<uint32_t>(kFoldStage)) ? n - j * kFoldStage : static_cast<uint32_t>(kFoldStage);
            if (s == 0u) { mbar_wait(&W.bar[0], phase0); phase0 ^= 1u; } else { mbar_wait(&W.bar[1], phase1); phase1 ^= 1u; }
            Frag f0{0.f, 0.f, 0.f, 0u}, f1{0.f, 0.f, 0.f, 0u};
            const bool a0 = static_cast<uint32_t>(lane) < c, a1 = static_cast<uint32_t>(32 + lane) < c;
            if (a0) f0 = W.stage[s][lane];
            if (a1) f1 = W.stage[s][32 + lane];
            if (whole) {
                const FoldPrep p0 = fold_prep(f0, a0, lo, A.time, lane);
                if (c > 32u) {
                    const FoldPrep p1 = fold_prep(f1, a1, lo, A.time, lane);
                    fold_apply(W, tex, p0, a0, lane);
                    fold_apply(W, tex, p1, a1, lane);
                } else {
                    fold_apply(W, tex, p0, a0, lane);
                }
            } else {
                // keep my texels' fragments, in order, in the ring; blend whenever 32 are queued
                const bool m0 = a0 && ((f0.key & kKeyLocalMask) - lo) < R, m1 = a1 && ((f1.key & kKeyLocalMask) - lo) < R;
                const uint32_t b0 = __ballot_sync(0xffffffffu, m0), b1 = __ballot_sync(0xffffffffu, m1);
                const uint32_t lt_mask = (1u << lane) - 1u;
                if (m0) W.ring[(head + queued + static_cast<uint32_t>(__popc(b0 & lt_mask))) & (kFoldRing - 1)] = f0;
                queued += static_cast<uint32_t>(__popc(b0));
                if (queued >= 32u) {
                    __syncwarp();
                    const Frag g = W.ring[(head + lane) & (kFoldRing - 1)];
                    fold_apply(W, tex, fold_prep(g, true, lo, A.time, lane), true, lane);
                    head = (head + 32u) & (kFoldRing - 1);
                    queued -= 32u;
                }
                if (m1) W.ring[(head + queued + static_cast<uint32_t>(__popc(b1 & lt_mask))) & (kFoldRing - 1)] = f1;
                queued += static_cast<uint32_t>(__popc(b1));
                if (queued >= 32u) {
                    __syncwarp();
                    const Frag g = W.ring[(head + lane) & (kFoldRing - 1)];
                    fold_apply(W, tex, fold_prep(g, true, lo, A.time, lane), true, lane);
                    head = (head + 32u) & (kFoldRing - 1);
                    queued -= 32u;
                }
                __syncwarp();
            }
            // (every lane is done with this window)
            if (lane == 0 && j + 2 < n_stage) issue(j + 2);
        }
        if (queued) {
            __syncwarp();
            Frag g{0.f, 0.f, 0.f, 0u};
            const bool act = static_cast<uint32_t>(lane) < queued;
            if (act) g = W.ring[(head + lane) & (kFoldRing - 1)];
            fold_apply(W, tex, fold_prep(g, act, lo, A.time, lane), act, lane);
        }
        for (uint32_t l = lane; l < R; l += 32) {
            const uint32_t loc = lo + l;
            const int gx = gx0 + static_cast<int>(loc & ((1u << A.g.sxl) - 1u)), gy = gy0 + static_cast<int>(loc >> A.g.sxl);
            if (gx < A.g.W && gy < A.g.H) {
                const float4 v = tex[l];
                const size_t at = static_cast<size_t>(gy) * A.g.W + gx;
                for (int r = 0; r < A.n_flow; ++r) A.flow[r][at] = v;
            }
        }
        __syncwarp();
    }
}

// Pass 5: join the segments of the long bins.  One warp per bin: the first segment's result is the truth; every later
// segment either settles a texel (take its value) or left the texel's fragments on record (replay them, in order).
__global__ void __launch_bounds__(kFoldThreads, 4) k_splat_mend(const FoldArgs A) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t S = 1u << (A.g.sxl + A.g.syl);
    FoldWarp &W = *reinterpret_cast<FoldWarp *>(smem_raw + fold_warp_bytes(static_cast<int>(S)) * warp);
    float4 *T = reinterpret_cast<float4 *>(&W + 1);
    uint32_t *mask = reinterpret_cast<uint32_t *>(W.ring);
    const uint32_t n_seg = *A.n_seg;
    if (n_seg == 0u) return;
    if (lane == 0) {
        mbar_init(&W.bar[0], 1u);
        mbar_init(&W.bar[1], 1u);
        mbar_fence_init();
    }
    __syncwarp();
    uint32_t phase0 = 0u, phase1 = 0u;
    for (;;) {
        uint32_t item = 0;
        if (lane == 0) item = atomicAdd(A.seg_ticket, 1u);
        item = __shfl_sync(0xffffffffu, item, 0);
        if (item >= n_seg) break;
        const uint4 d = A.seg_desc[item];
        const uint32_t bin_id = d.x, lp = d.y;
        const uint32_t code = A.bin_info[bin_id];
        const uint32_t st = code & 0xffffu, sub = (code >> 16) & 0xffu, ls = code >> 24;
        const uint32_t R = S >> ls, lo = sub * R;
        const uint32_t begin = A.bin_off[bin_id], n = A.bin_count ? A.bin_count[bin_id] : A.bin_off[bin_id + 1] - begin;
        const uint32_t span = seg_span(n, lp);
        const int gx0 = (static_cast<int>(st) % A.g.strips_x) << A.g.sxl, gy0 = (static_cast<int>(st) / A.g.strips_x) << A.g.syl;
        for (uint32_t l = lane; l < R; l += 32) T[l] = A.seg_out[d.z + l];
        __syncwarp();
        bool redo = false;
        for (uint32_t part = 1; part < (1u << lp) && !redo; ++part) {
            const float4 *E = A.seg_out + d.z + static_cast<size_t>(part) * R;
            redo = mend_begin(T, E, R, mask, lane);
            if (redo) break;
            const uint32_t b0 = part * span < n ? part * span : n;
            fold_stream(W, A.replay + begin + b0, A.seg_cnt[d.w + part], lane, phase0, phase1,
                        [&](const Frag &f, bool act) { mend_batch(W, T, mask, f, act, lo, A.time, lane); });
            mend_end(T, E, R, mask, lane);
        }
        if (redo) {                                        // a non-finite texel, or a settled chain that overflowed: the plain way
            fold_load_texels(A, T, lo, R, gx0, gy0, lane);
            fold_stream(W, A.bins + begin, n, lane, phase0, phase1,
                        [&](const Frag &f, bool act) { fold_apply(W, T, fold_prep(f, act, lo, A.time, lane), act, lane); });
        }
        fold_store_texels(A, T, lo, R, gx0, gy0, lane);
    }
}

}  // namespace tb
