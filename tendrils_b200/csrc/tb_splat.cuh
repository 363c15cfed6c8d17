// tb_splat.cuh -- the ordered flow splat (a7-a10) as a tile-binned pipeline, sm_100a.
//
// The reference draws every particle as a GL_LINES segment prev -> cur into the flow FBO under
// blendFunc(SRC_ALPHA, ONE_MINUS_SRC_ALPHA), in primitive order p = x*PH + k
// (src/index.js:267-268,278-303, src/particles.js:147-158,182-186).  The blend is not
// commutative, so per texel the fragments must be applied in draw order.  This file does that
// without a global sort:
//
//   k_splat_hist     per slab of consecutive primitives: fragments per grid TILE            (count)
//   k_splat_rows     per tile: exclusive scan over the slabs, tile totals                    (scan)
//   k_splat_plan     tile totals -> bin offsets, the fold work list, capacity check          (scan)
//   k_splat_scatter  re-rasterise; every fragment goes straight to its slot of its tile's bin,
//                    bins filled in draw order (a stable multi-split: warp match + per-warp
//                    counters, no global atomics on the data path)                           (emit)
//   k_splat_fold     per tile: stream the bin through shared memory (cp.async.bulk + mbarrier,
//                    double buffered), blend in order onto the tile held in shared memory,
//                    write the tile back                                                     (fold)
//
// A fragment is written once (16 B) and read once; nothing is sorted, nothing returns to the host.
// Compiled with -fmad=false; see tb_math.cuh for the arithmetic contract.
#pragma once

#include "tb_kernels.cuh"

namespace tb {

// ------------------------------------------------------------------------------------------
// Geometry of the binning: the grid is cut into tiles of (1 << txl) x (1 << tyl) texels.
// ------------------------------------------------------------------------------------------
constexpr int kMaxTiles = 2048;           // bins per grid (k_splat_scatter keeps per-warp counters per tile in shared memory)
constexpr int kFoldTexels = 1024;         // texels one fold work item holds in shared memory
constexpr uint32_t kKeyLocalMask = 0x000fffffu;
constexpr uint32_t kKeyTame = 0x40000000u;     // cx, cy finite and 0 <= a <= 1
constexpr uint32_t kKeyOpaque = 0x80000000u;   // a == 1 and no colour channel is -0: the fragment overwrites its texel

struct TileGeom {
    int W, H;
    int txl, tyl;            // log2 of the tile width / height
    int tiles_x, tiles_y;
    int T;                   // tiles_x * tiles_y <= kMaxTiles
};

__host__ __device__ __forceinline__ int tile_of(const TileGeom &g, int gx, int gy) { return (gy >> g.tyl) * g.tiles_x + (gx >> g.txl); }
__host__ __device__ __forceinline__ uint32_t local_of(const TileGeom &g, int gx, int gy) {
    return (static_cast<uint32_t>(gy & ((1 << g.tyl) - 1)) << g.txl) | static_cast<uint32_t>(gx & ((1 << g.txl) - 1));
}

// One fragment in a bin: the interpolated colour's (vel.xy, alpha) -- the time channel is the uniform
// `time` -- and where it goes inside its tile.
struct __align__(16) Frag {
    float cx, cy, a;
    uint32_t key;            // local texel index | kKeyTame | kKeyOpaque
};

// One unit of fold work: texels [lo, hi) (local indices) of tile `tile`, fed by bin [begin, end).
struct FoldItem {
    uint32_t tile, lo, hi, pad;
    uint32_t begin, end;     // fragment range in the bin array
    uint32_t pad2[2];
};

// What the plan kernel leaves for the host (read lazily, never waited for on the hot path).
struct PlanOut {
    unsigned long long total;     // fragments of this draw
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
    uint32_t n = 0;
    raster_line(xa, ya, xb, yb, W, H, [&](int, int, float) { ++n; });
    return n;
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
    uint32_t k = 0;
    gx = gy = 0; t = 0.0f;
    raster_line(xmajor ? P.ma : P.na, xmajor ? P.na : P.ma, xmajor ? P.mb : P.nb, xmajor ? P.nb : P.mb, W, H,
                [&](int x, int y, float tt) { if (k == j) { gx = x; gy = y; t = tt; } ++k; });
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
    const int xl = static_cast<int>(p / S.n_pairs);
    const int4 pv = __ldg(reinterpret_cast<const int4 *>(S.pairs + (p - static_cast<long long>(xl) * S.n_pairs)));
    PairEntry pe;
    pe.k = pv.x; pe.row_a = pv.y; pe.row_b = pv.z; pe.pad = pv.w;
    const size_t base = static_cast<size_t>(xl) * S.PH;
    sa = __ldg(((pe.row_a < 0) ? S.cur : S.prev) + base + (pe.row_a & 0x7fffffff));
    sb = __ldg(((pe.row_b < 0) ? S.cur : S.prev) + base + (pe.row_b & 0x7fffffff));
}

// ------------------------------------------------------------------------------------------
// Pass 1: fragments per (tile, slab).  A slab is a fixed range of consecutive primitives; CTAs take
// slabs from a ticket counter.  hist[tile * n_slabs + slab].
// ------------------------------------------------------------------------------------------
constexpr int kHistThreads = 256;

struct HistArgs {
    PrimSource src;
    TileGeom g;
    float vsx, vsy;
    int slab_prims, n_slabs;
    uint32_t *__restrict__ slab_hist;
    uint32_t *ticket;
};

__global__ void __launch_bounds__(kHistThreads) k_splat_hist(const HistArgs A) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint32_t *hist = reinterpret_cast<uint32_t *>(smem_raw);              // [T]
    __shared__ int s_slab;
    const int T = A.g.T;
    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) s_slab = static_cast<int>(atomicAdd(A.ticket, 1u));
        for (int t = threadIdx.x; t < T; t += kHistThreads) hist[t] = 0u;
        __syncthreads();
        const int slab = s_slab;
        if (slab >= A.n_slabs) break;
        const long long p0 = static_cast<long long>(slab) * A.slab_prims;
        const long long p1 = (p0 + A.slab_prims < A.src.n_prims) ? p0 + A.slab_prims : A.src.n_prims;
        for (long long p = p0 + threadIdx.x; p < p1; p += kHistThreads) {
            float4 sa, sb;
            load_prim(A.src, p, sa, sb);
            PrimGeom P;
            const uint32_t n = prim_setup(sa, sb, A.vsx, A.vsy, A.g.W, A.g.H, P);
            // runs of fragments in one tile are added at once
            int run_tile = -1;
            uint32_t run = 0;
            if (n != 0u && !(P.flags & 2u)) {
                const bool xmajor = (P.flags & 1u) != 0u;
                raster_line(xmajor ? P.ma : P.na, xmajor ? P.na : P.ma, xmajor ? P.mb : P.nb, xmajor ? P.nb : P.mb, A.g.W, A.g.H,
                            [&](int gx, int gy, float) {
                                const int tl = tile_of(A.g, gx, gy);
                                if (tl != run_tile) { if (run) atomicAdd(&hist[run_tile], run); run_tile = tl; run = 0; }
                                ++run;
                            });
            } else {
                for (uint32_t j = 0; j < n; ++j) {
                    int gx, gy; float t;
                    prim_fragment(P, j, A.g.W, A.g.H, gx, gy, t);
                    const int tl = tile_of(A.g, gx, gy);
                    if (tl != run_tile) { if (run) atomicAdd(&hist[run_tile], run); run_tile = tl; run = 0; }
                    ++run;
                }
            }
            if (run) atomicAdd(&hist[run_tile], run);
        }
        __syncthreads();
        for (int t = threadIdx.x; t < T; t += kHistThreads) A.slab_hist[static_cast<size_t>(t) * A.n_slabs + slab] = hist[t];
    }
}

// ------------------------------------------------------------------------------------------
// Pass 2a: per tile (one warp each) the exclusive scan of its row of slab counts, in place, and the tile's total.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t warp_incl_scan(uint32_t v, int lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t u = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += u;
    }
    return v;
}

__global__ void __launch_bounds__(256) k_splat_rows(uint32_t *__restrict__ slab_hist, int T, int n_slabs,
                                                    uint32_t *__restrict__ tile_total, uint32_t *__restrict__ too_many) {
    const int lane = threadIdx.x & 31;
    const int t = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (t >= T) return;
    uint32_t *row = slab_hist + static_cast<size_t>(t) * n_slabs;
    unsigned long long carry = 0ull;
    for (int s0 = 0; s0 < n_slabs; s0 += 32) {
        const int s = s0 + lane;
        const uint32_t v = (s < n_slabs) ? row[s] : 0u;
        const uint32_t inc = warp_incl_scan(v, lane);
        if (s < n_slabs) row[s] = static_cast<uint32_t>(carry) + inc - v;
        carry += __shfl_sync(0xffffffffu, inc, 31);
    }
    if (lane == 0) {
        tile_total[t] = static_cast<uint32_t>(carry);
        if (carry > 0xffffffffull) *too_many = 1u;          // a single bin beyond 2^32 fragments: reported as overflow
    }
}

// ------------------------------------------------------------------------------------------
// Pass 2b (single GPU): bin offsets, the fold work list (largest bins first), the capacity check.  One CTA.
// ------------------------------------------------------------------------------------------
constexpr int kPlanThreads = 1024;

struct PlanArgs {
    TileGeom g;
    const uint32_t *__restrict__ tile_total;   // [T]
    uint32_t *__restrict__ bin_off;            // [T + 1]
    FoldItem *__restrict__ items;              // [max_items]
    int max_items;
    uint32_t cap;                              // capacity of the bin array (fragments)
    uint32_t hot_bin;                          // bins longer than this are split into sub-items by texel rows
    uint32_t *too_many;                        // set by k_splat_rows; reset here
    uint32_t *tickets;                         // [0] hist (reset for the next draw), [1] scatter, [2] fold, [3] items
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

// sub-items a bin of `n` fragments over `L` local texels is cut into (powers of two; each scans the whole bin)
__device__ __forceinline__ uint32_t fold_splits(uint32_t n, uint32_t L, uint32_t hot_bin) {
    uint32_t s = (L + kFoldTexels - 1) / kFoldTexels;
    if (s == 0) s = 1;
    uint32_t per = L / s;
    while (n > hot_bin && per > 32u && s < 64u) { s *= 4u; per /= 4u; n /= 4u; }
    return s;
}

__global__ void __launch_bounds__(kPlanThreads) k_splat_plan(const PlanArgs A) {
    __shared__ unsigned long long s_warp[32];
    __shared__ unsigned long long s_total;
    __shared__ uint32_t s_bucket[33];
    __shared__ uint32_t s_ok;
    const int T = A.g.T;
    const uint32_t L = 1u << (A.g.txl + A.g.tyl);
    // two tiles per thread (T <= 2048)
    const int t0 = threadIdx.x * 2, t1 = t0 + 1;
    const uint32_t n0 = (t0 < T) ? A.tile_total[t0] : 0u, n1 = (t1 < T) ? A.tile_total[t1] : 0u;
    const unsigned long long ex = block_excl_scan64(static_cast<unsigned long long>(n0) + n1, s_warp, &s_total);
    if (threadIdx.x == 0) {
        const bool ok = s_total <= static_cast<unsigned long long>(A.cap) && *A.too_many == 0u;
        s_ok = ok ? 1u : 0u;
        A.out->total = s_total;
        A.out->overflow = ok ? 0u : 1u;
        A.tickets[0] = 0u; A.tickets[1] = 0u; A.tickets[2] = 0u;
        *A.too_many = 0u;
    }
    if (threadIdx.x < 33) s_bucket[threadIdx.x] = 0u;
    __syncthreads();
    const bool ok = s_ok != 0u;
    if (t0 < T) A.bin_off[t0] = ok ? static_cast<uint32_t>(ex) : 0u;
    if (t1 < T) A.bin_off[t1] = ok ? static_cast<uint32_t>(ex + n0) : 0u;
    if (threadIdx.x == 0) A.bin_off[T] = ok ? static_cast<uint32_t>(s_total) : 0u;
    // work list, longest bins first (bucketed by log2 of the length): the tail of the fold is short items
    const uint32_t sp0 = (ok && n0) ? fold_splits(n0, L, A.hot_bin) : 0u, sp1 = (ok && n1) ? fold_splits(n1, L, A.hot_bin) : 0u;
    const int b0 = n0 ? 31 - __clz(n0 / (sp0 ? sp0 : 1u) | 1u) : 0, b1 = n1 ? 31 - __clz(n1 / (sp1 ? sp1 : 1u) | 1u) : 0;
    uint32_t r0 = 0, r1 = 0;
    if (sp0) r0 = atomicAdd(&s_bucket[31 - b0], sp0);
    if (sp1) r1 = atomicAdd(&s_bucket[31 - b1], sp1);
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t run = 0;
        for (int b = 0; b < 32; ++b) { const uint32_t c = s_bucket[b]; s_bucket[b] = run; run += c; }
        s_bucket[32] = run;
        const uint32_t n_items = run <= static_cast<uint32_t>(A.max_items) ? run : 0u;
        A.out->n_items = n_items;
        A.tickets[3] = n_items;
        if (run > static_cast<uint32_t>(A.max_items)) A.out->overflow = 1u;      // cannot happen with max_items = 64 * T
    }
    __syncthreads();
    if (s_bucket[32] > static_cast<uint32_t>(A.max_items)) return;
    auto put = [&](int t, uint32_t n, uint32_t sp, int b, uint32_t r, uint32_t begin) {
        const uint32_t per = L / sp;
        for (uint32_t k = 0; k < sp; ++k) {
            FoldItem it{};
            it.tile = static_cast<uint32_t>(t);
            it.lo = k * per; it.hi = (k + 1) * per;
            it.begin = begin; it.end = begin + n;
            A.items[s_bucket[31 - b] + r + k] = it;
        }
    };
    if (sp0) put(t0, n0, sp0, b0, r0, static_cast<uint32_t>(ex));
    if (sp1) put(t1, n1, sp1, b1, r1, static_cast<uint32_t>(ex + n0));
}

// ------------------------------------------------------------------------------------------
// Pass 3: emit.  Every fragment is computed once and stored once, at its final place: bin of its tile, draw order.
//
// A CTA takes a slab (ticket), and walks it in windows of kEmitThreads primitives (one per thread: load the two
// vertices, set the line up, count its fragments).  The window's fragments are numbered in draw order by a block
// scan ("slots") and handled in passes of at most kEmitSlots: the primitives expand their slot range into a table,
// then the threads take the slots warp-striped -- warp w the w-th stretch, 32 consecutive slots per round -- so
// that (warp, round, lane) order is slot order.  Rank of a fragment among those of its tile:
//   within the round       __match_any_sync on the tile,
//   within the warp        a per-warp counter per tile, advanced by the round's leader,
//   across the warps       exclusive scan of those counters over the warps (one thread per tile pair),
//   across passes/windows  a per-CTA cursor per tile, which starts at bin_off[tile] + (this slab's prefix).
// ------------------------------------------------------------------------------------------
constexpr int kEmitThreads = 512;
constexpr int kEmitWarps = kEmitThreads / 32;
constexpr int kEmitSlots = 2048;                          // slots per pass
constexpr int kEmitRounds = kEmitSlots / kEmitThreads;    // rounds of 32 slots per warp and pass

struct ScatterArgs {
    PrimSource src;
    TileGeom g;
    float vsx, vsy, speedLimit, time;
    int slab_prims, n_slabs;
    const uint32_t *__restrict__ slab_hist;     // scanned rows: fragments of this tile in earlier slabs
    const uint32_t *__restrict__ bin_off;       // [T + 1] (sharded run: where this rank's fragments start in the owner's bin)
    const PlanOut *plan;
    uint32_t *ticket;
    Frag *bins[kMaxBandRanks];                  // the bin array of every rank (single GPU: [0])
    const uint8_t *__restrict__ tile_owner;     // null: everything goes to bins[0]
};

__host__ __device__ inline size_t scatter_smem_bytes(int T) {
    const size_t Tp = static_cast<size_t>((T + 1) / 2);
    return static_cast<size_t>(kEmitThreads) * 12 * 4        // primitive records (SoA)
           + static_cast<size_t>(kEmitSlots) * 4             // slot -> (primitive, fragment) table
           + static_cast<size_t>(kEmitWarps) * Tp * 4        // per-warp counters, two tiles per word
           + Tp * 4                                          // per-pass totals
           + static_cast<size_t>(T) * 4                      // cursors
           + 64 * 4;                                         // scan scratch
}

__global__ void __launch_bounds__(kEmitThreads) k_splat_scatter(const ScatterArgs A) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    if (A.plan->overflow) return;
    const int T = A.g.T, Tp = (T + 1) / 2;
    float *rec = reinterpret_cast<float *>(smem_raw);                          // [12][kEmitThreads]
    uint32_t *owner = reinterpret_cast<uint32_t *>(rec + 12 * kEmitThreads);   // [kEmitSlots]
    uint32_t *wh = owner + kEmitSlots;                                         // [kEmitWarps][Tp]
    uint32_t *tot = wh + kEmitWarps * Tp;                                      // [Tp]
    uint32_t *cur = tot + Tp;                                                  // [T]
    uint32_t *scratch = cur + T;                                               // [64]
    __shared__ int s_slab;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t lt_mask = (1u << lane) - 1u;
    const bool opaque_ok = __float_as_uint(A.time) != 0x80000000u;            // time*1 must not be -0 for the cut to be exact

    for (;;) {
        __syncthreads();
        if (tid == 0) s_slab = static_cast<int>(atomicAdd(A.ticket, 1u));
        __syncthreads();
        const int slab = s_slab;
        if (slab >= A.n_slabs) break;
        for (int t = tid; t < T; t += kEmitThreads) cur[t] = A.bin_off[t] + A.slab_hist[static_cast<size_t>(t) * A.n_slabs + slab];
        for (int i = tid; i < kEmitWarps * Tp; i += kEmitThreads) wh[i] = 0u;
        const long long p0 = static_cast<long long>(slab) * A.slab_prims;
        const long long p1 = (p0 + A.slab_prims < A.src.n_prims) ? p0 + A.slab_prims : A.src.n_prims;
        for (long long w0 = p0; w0 < p1; w0 += kEmitThreads) {
            // ---- window setup: one primitive per thread
            const long long p = w0 + tid;
            uint32_t n = 0;
            if (p < p1) {
                float4 sa, sb;
                load_prim(A.src, p, sa, sb);
                PrimGeom P;
                n = prim_setup(sa, sb, A.vsx, A.vsy, A.g.W, A.g.H, P);
                if (n) {
                    // flow(vel, speedLimit): src/flow/apply/state.glsl:5-16
                    const float aa = gmin(__fdiv_rn(glength(sa.z, sa.w), A.speedLimit), 1.0f);
                    const float ab = gmin(__fdiv_rn(glength(sb.z, sb.w), A.speedLimit), 1.0f);
                    rec[0 * kEmitThreads + tid] = P.ma; rec[1 * kEmitThreads + tid] = P.mb;
                    rec[2 * kEmitThreads + tid] = P.na; rec[3 * kEmitThreads + tid] = P.nb;
                    rec[4 * kEmitThreads + tid] = __int_as_float(P.c0);
                    rec[5 * kEmitThreads + tid] = __uint_as_float(P.flags);
                    rec[6 * kEmitThreads + tid] = sa.z; rec[7 * kEmitThreads + tid] = sb.z;
                    rec[8 * kEmitThreads + tid] = sa.w; rec[9 * kEmitThreads + tid] = sb.w;
                    rec[10 * kEmitThreads + tid] = aa; rec[11 * kEmitThreads + tid] = ab;
                }
            }
            // ---- slots: exclusive scan of the counts over the window
            uint32_t inc = warp_incl_scan(n, lane);
            __syncthreads();                                                    // (previous window's passes are done with scratch)
            if (lane == 31) scratch[warp] = inc;
            __syncthreads();
            if (warp == 0) {
                const uint32_t ws = (lane < kEmitWarps) ? scratch[lane] : 0u;
                const uint32_t wi = warp_incl_scan(ws, lane);
                if (lane < kEmitWarps) scratch[32 + lane] = wi - ws;
                if (lane == kEmitWarps - 1) scratch[63] = wi;
            }
            __syncthreads();
            const uint32_t my_off = scratch[32 + warp] + inc - n;
            const uint32_t n_window = scratch[63];
            for (uint32_t s_lo = 0; s_lo < n_window; s_lo += kEmitSlots) {
                const uint32_t cnt = (n_window - s_lo < static_cast<uint32_t>(kEmitSlots)) ? n_window - s_lo : static_cast<uint32_t>(kEmitSlots);
                // ---- the primitives expand their slots of this pass
                {
                    const uint32_t b = my_off > s_lo ? my_off : s_lo;
                    const uint32_t e = (my_off + n < s_lo + cnt) ? my_off + n : s_lo + cnt;
                    for (uint32_t s = b; s < e; ++s) owner[s - s_lo] = (static_cast<uint32_t>(tid) << 20) | (s - my_off);
                }
                __syncthreads();
                // ---- fragments, warp-striped, and their rank within the warp
                const uint32_t per = ((cnt + kEmitWarps - 1) / kEmitWarps + 31u) & ~31u;     // slots per warp, <= kEmitRounds * 32
                float fcx[kEmitRounds], fcy[kEmitRounds], fa[kEmitRounds];
                uint32_t fkey[kEmitRounds], ftile[kEmitRounds], frank[kEmitRounds];
#pragma unroll
                for (int r = 0; r < kEmitRounds; ++r) {
                    const uint32_t s = warp * per + r * 32 + lane;
                    const bool valid = static_cast<uint32_t>(r * 32) < per && s < cnt;
                    ftile[r] = 0xffffffffu;
                    if (static_cast<uint32_t>(r * 32) >= per || warp * per + r * 32 >= cnt) continue;   // warp-uniform
                    if (valid) {
                        const uint32_t o = owner[s];
                        const int q = static_cast<int>(o >> 20);
                        PrimGeom P;
                        P.ma = rec[0 * kEmitThreads + q]; P.mb = rec[1 * kEmitThreads + q];
                        P.na = rec[2 * kEmitThreads + q]; P.nb = rec[3 * kEmitThreads + q];
                        P.c0 = __float_as_int(rec[4 * kEmitThreads + q]);
                        P.flags = __float_as_uint(rec[5 * kEmitThreads + q]);
                        int gx, gy; float t;
                        prim_fragment(P, o & 0xfffffu, A.g.W, A.g.H, gx, gy, t);
                        const float za = rec[6 * kEmitThreads + q], zb = rec[7 * kEmitThreads + q];
                        const float wa = rec[8 * kEmitThreads + q], wb = rec[9 * kEmitThreads + q];
                        const float aa = rec[10 * kEmitThreads + q], ab = rec[11 * kEmitThreads + q];
                        const float cx = __fadd_rn(za, __fmul_rn(t, __fsub_rn(zb, za)));
                        const float cy = __fadd_rn(wa, __fmul_rn(t, __fsub_rn(wb, wa)));
                        const float a = __fadd_rn(aa, __fmul_rn(t, __fsub_rn(ab, aa)));
                        uint32_t key = local_of(A.g, gx, gy);
                        const bool tame = is_finite(cx) && is_finite(cy) && a >= 0.0f && a <= 1.0f;
                        if (tame) key |= kKeyTame;
                        if (opaque_ok && a == 1.0f && tame && __float_as_uint(cx) != 0x80000000u && __float_as_uint(cy) != 0x80000000u)
                            key |= kKeyOpaque;
                        fcx[r] = cx; fcy[r] = cy; fa[r] = a; fkey[r] = key;
                        ftile[r] = static_cast<uint32_t>(tile_of(A.g, gx, gy));
                    }
                    const uint32_t peers = __match_any_sync(0xffffffffu, ftile[r]);
                    if (valid) {
                        const int leader = __ffs(peers) - 1;
                        uint32_t base = 0;
                        const uint32_t sh = (ftile[r] & 1u) * 16u;
                        if (lane == leader) base = (atomicAdd(&wh[warp * Tp + (ftile[r] >> 1)], static_cast<uint32_t>(__popc(peers)) << sh) >> sh) & 0xffffu;
                        base = __shfl_sync(peers, base, leader);
                        frank[r] = base + static_cast<uint32_t>(__popc(peers & lt_mask));
                    }
                }
                __syncthreads();
                // ---- across the warps: exclusive scan of the per-warp counters, two tiles per word
                for (int tp = tid; tp < Tp; tp += kEmitThreads) {
                    uint32_t run = 0;
#pragma unroll
                    for (int w = 0; w < kEmitWarps; ++w) {
                        const uint32_t c = wh[w * Tp + tp];
                        wh[w * Tp + tp] = run;
                        run += c;
                    }
                    tot[tp] = run;
                }
                __syncthreads();
                // ---- store
#pragma unroll
                for (int r = 0; r < kEmitRounds; ++r) {
                    if (ftile[r] == 0xffffffffu) continue;
                    const uint32_t tl = ftile[r], sh = (tl & 1u) * 16u;
                    const uint32_t dst = cur[tl] + ((wh[warp * Tp + (tl >> 1)] >> sh) & 0xffffu) + frank[r];
                    Frag *bin = A.bins[A.tile_owner ? A.tile_owner[tl] : 0];
                    *reinterpret_cast<float4 *>(bin + dst) = make_float4(fcx[r], fcy[r], fa[r], __uint_as_float(fkey[r]));
                }
                __syncthreads();
                for (int t = tid; t < T; t += kEmitThreads) cur[t] += (tot[t >> 1] >> ((t & 1) * 16)) & 0xffffu;
                for (int i = tid; i < kEmitWarps * Tp; i += kEmitThreads) wh[i] = 0u;
                // (the next pass / window starts with a barrier before it touches owner, wh or cur)
                __syncthreads();
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
// Pass 4: fold.  blendFunc(SRC_ALPHA, ONE_MINUS_SRC_ALPHA) on all four channels in primitive order
// (src/index.js:267-268): dst = src*a + dst*(1-a), two roundings per channel.
//
// A CTA takes work items (ticket): texels [lo, hi) of one tile live in shared memory while the tile's bin streams
// through a double-buffered shared-memory window (cp.async.bulk + mbarrier).  Warp w owns an equal share of the
// texels.  Every warp scans every chunk (one 16-byte shared load per lane and 32 fragments) and appends the
// fragments of its texels, in order, to a small ring; whenever 32 are queued it blends them: lanes whose texels
// differ go in parallel, lanes that share a texel (__match_any_sync) are chained in lane order = draw order, every
// lane of the group running the same chain.  A fragment with alpha == 1 overwrites the texel: the chain starts at
// the group's last such fragment (exact: see kKeyOpaque / kKeyTame).
// ------------------------------------------------------------------------------------------
constexpr int kFoldThreads = 256;
constexpr int kFoldNWarps = kFoldThreads / 32;
constexpr int kFoldChunkFrags = 2048;                    // fragments per shared-memory window (32 KiB)
constexpr int kFoldStages = 2;
constexpr int kFoldRing = 64;                            // queued fragments per warp

struct FoldArgs {
    TileGeom g;
    float time;
    const Frag *__restrict__ bins;
    const FoldItem *__restrict__ items;
    const uint32_t *n_items;                 // device: number of work items
    uint32_t *ticket;
    float4 *flow[kMaxBandRanks];             // [0] the grid that is read; every non-null entry is written
    int n_flow;
};

constexpr size_t kFoldSmemBytes = static_cast<size_t>(kFoldStages) * kFoldChunkFrags * sizeof(Frag)   // windows
                                  + static_cast<size_t>(kFoldTexels) * sizeof(float4)                 // the tile
                                  + static_cast<size_t>(kFoldNWarps) * kFoldRing * sizeof(Frag)       // rings
                                  + 64;                                                               // mbarriers

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

__device__ __forceinline__ void fold_batch(const Frag *ring, uint32_t head, uint32_t nb, float4 *tile, uint32_t lo, float time, int lane) {
    const bool active = static_cast<uint32_t>(lane) < nb;
    Frag f{};
    if (active) f = ring[(head + lane) & (kFoldRing - 1)];
    const uint32_t loc = active ? (f.key & kKeyLocalMask) : (0xffffff00u | static_cast<uint32_t>(lane));
    const float a = f.a;
    const float tx = __fmul_rn(f.cx, a), ty = __fmul_rn(f.cy, a), tz = __fmul_rn(time, a), tw = __fmul_rn(a, a);
    const float om = __fsub_rn(1.0f, a);
    const uint32_t peers = __match_any_sync(0xffffffffu, loc);
    const uint32_t opq = __ballot_sync(0xffffffffu, active && (f.key & kKeyOpaque)) & peers;
    const uint32_t wild = __ballot_sync(0xffffffffu, active && !(f.key & kKeyTame)) & peers;
    uint32_t rem = peers;
    if (opq) {
        const uint32_t below = (1u << (31 - __clz(opq))) - 1u;      // lanes before the group's last opaque fragment
        if ((wild & below) == 0u) rem = peers & ~below;
    }
    const uint32_t cnt = static_cast<uint32_t>(__popc(rem));
    const bool shared = __any_sync(0xffffffffu, (peers & (peers - 1u)) != 0u);   // some texel has more than one fragment here
    float4 d = make_float4(0.f, 0.f, 0.f, 0.f);
    if (active) d = tile[loc - lo];
    if (!shared) {
        d.x = __fadd_rn(tx, __fmul_rn(d.x, om)); d.y = __fadd_rn(ty, __fmul_rn(d.y, om));
        d.z = __fadd_rn(tz, __fmul_rn(d.z, om)); d.w = __fadd_rn(tw, __fmul_rn(d.w, om));
    } else {
        const uint32_t longest = __reduce_max_sync(0xffffffffu, cnt);
        for (uint32_t s = 0; s < longest; ++s) {
            const int j = rem ? __ffs(rem) - 1 : lane;
            const bool step = rem != 0u;
            rem &= rem - 1u;
            const float sx = __shfl_sync(0xffffffffu, tx, j), sy = __shfl_sync(0xffffffffu, ty, j);
            const float sz = __shfl_sync(0xffffffffu, tz, j), sw = __shfl_sync(0xffffffffu, tw, j);
            const float sm = __shfl_sync(0xffffffffu, om, j);
            if (step) {
                d.x = __fadd_rn(sx, __fmul_rn(d.x, sm)); d.y = __fadd_rn(sy, __fmul_rn(d.y, sm));
                d.z = __fadd_rn(sz, __fmul_rn(d.z, sm)); d.w = __fadd_rn(sw, __fmul_rn(d.w, sm));
            }
        }
    }
    if (active && lane == __ffs(peers) - 1) tile[loc - lo] = d;
    __syncwarp();
}

__global__ void __launch_bounds__(kFoldThreads) k_splat_fold(const FoldArgs A) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Frag *win = reinterpret_cast<Frag *>(smem_raw);                                             // [kFoldStages][kFoldChunkFrags]
    float4 *tile = reinterpret_cast<float4 *>(win + kFoldStages * kFoldChunkFrags);            // [kFoldTexels]
    Frag *rings = reinterpret_cast<Frag *>(tile + kFoldTexels);                                 // [kFoldNWarps][kFoldRing]
    unsigned long long *bars = reinterpret_cast<unsigned long long *>(rings + kFoldNWarps * kFoldRing);   // [kFoldStages]
    __shared__ uint32_t s_item;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t lt_mask = (1u << lane) - 1u;
    const uint32_t n_items = *A.n_items;
    if (tid == 0) {
        for (int s = 0; s < kFoldStages; ++s) mbar_init(&bars[s], 1u);
        mbar_fence_init();
    }
    uint32_t phase[kFoldStages];
#pragma unroll
    for (int s = 0; s < kFoldStages; ++s) phase[s] = 0u;
    Frag *ring = rings + warp * kFoldRing;
    for (;;) {
        __syncthreads();
        if (tid == 0) s_item = atomicAdd(A.ticket, 1u);
        __syncthreads();
        const uint32_t item = s_item;
        if (item >= n_items) break;
        const FoldItem it = A.items[item];
        const int tile_x = static_cast<int>(it.tile) % A.g.tiles_x, tile_y = static_cast<int>(it.tile) / A.g.tiles_x;
        const int gx0 = tile_x << A.g.txl, gy0 = tile_y << A.g.tyl;
        const uint32_t span = it.hi - it.lo;
        const uint32_t n = it.end - it.begin, n_chunks = (n + kFoldChunkFrags - 1) / kFoldChunkFrags;
        const Frag *bin = A.bins + it.begin;
        auto issue = [&](uint32_t k) {
            const uint32_t c = (n - k * kFoldChunkFrags < static_cast<uint32_t>(kFoldChunkFrags)) ? n - k * kFoldChunkFrags : static_cast<uint32_t>(kFoldChunkFrags);
            bulk_load(win + (k % kFoldStages) * kFoldChunkFrags, bin + static_cast<size_t>(k) * kFoldChunkFrags, c * static_cast<uint32_t>(sizeof(Frag)),
                      &bars[k % kFoldStages]);
        };
        if (tid == 0)
            for (uint32_t k = 0; k < kFoldStages && k < n_chunks; ++k) issue(k);
        // the tile's texels [lo, hi)
        for (uint32_t l = it.lo + tid; l < it.hi; l += kFoldThreads) {
            const int gx = gx0 + static_cast<int>(l & ((1u << A.g.txl) - 1u)), gy = gy0 + static_cast<int>(l >> A.g.txl);
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (gx < A.g.W && gy < A.g.H) v = A.flow[0][static_cast<size_t>(gy) * A.g.W + gx];
            tile[l - it.lo] = v;
        }
        __syncthreads();
        // this warp's texels: [wlo, wlo + wspan)
        const uint32_t wspan = (span + kFoldNWarps - 1) / kFoldNWarps;
        const uint32_t wlo = it.lo + warp * wspan;
        uint32_t head = 0, queued = 0;
        for (uint32_t k = 0; k < n_chunks; ++k) {
            const int st = static_cast<int>(k % kFoldStages);
            const uint32_t c = (n - k * kFoldChunkFrags < static_cast<uint32_t>(kFoldChunkFrags)) ? n - k * kFoldChunkFrags : static_cast<uint32_t>(kFoldChunkFrags);
            mbar_wait(&bars[st], phase[st]);
            phase[st] ^= 1u;
            const Frag *buf = win + st * kFoldChunkFrags;
            for (uint32_t i0 = 0; i0 < c; i0 += 32) {
                const uint32_t i = i0 + lane;
                Frag f{};
                bool mine = false;
                if (i < c) {
                    f = buf[i];
                    const uint32_t loc = f.key & kKeyLocalMask;
                    mine = (loc - wlo) < wspan && loc < it.hi;
                }
                const uint32_t m = __ballot_sync(0xffffffffu, mine);
                if (m == 0u) continue;
                if (mine) ring[(head + queued + static_cast<uint32_t>(__popc(m & lt_mask))) & (kFoldRing - 1)] = f;
                queued += static_cast<uint32_t>(__popc(m));
                if (queued >= 32u) {
                    __syncwarp();
                    fold_batch(ring, head, 32u, tile, it.lo, A.time, lane);
                    head = (head + 32u) & (kFoldRing - 1);
                    queued -= 32u;
                }
            }
            __syncthreads();                                         // every warp is done with this window
            if (tid == 0 && k + kFoldStages < n_chunks) issue(k + kFoldStages);
        }
        if (queued) {
            __syncwarp();
            fold_batch(ring, head, queued, tile, it.lo, A.time, lane);
        }
        __syncthreads();
        for (uint32_t l = it.lo + tid; l < it.hi; l += kFoldThreads) {
            const int gx = gx0 + static_cast<int>(l & ((1u << A.g.txl) - 1u)), gy = gy0 + static_cast<int>(l >> A.g.txl);
            if (gx < A.g.W && gy < A.g.H) {
                const float4 v = tile[l - it.lo];
                const size_t at = static_cast<size_t>(gy) * A.g.W + gx;
                for (int r = 0; r < A.n_flow; ++r) A.flow[r][at] = v;
            }
        }
    }
}

}  // namespace tb
