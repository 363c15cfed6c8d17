// tb_api.cu -- the C ABI of include/tendrils_b200.h over the sm_100a kernels.
//
// Mirrors, for the one hot path, the GPU-facing behaviour of the reference's
// Tendrils (src/index.js) and Particles (src/particles.js) classes: ping-pong state
// buffers, the logic pass, the flow splat, the spawn passes.  No CPU fallback.
#include "tb_kernels.cuh"
#include "tb_splat.cuh"
#include "tb_owners.cuh"
#include "tb_flowline.cuh"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <climits>
#include <cstring>
#include <string>
#include <vector>

using namespace tb;

namespace {
thread_local std::string g_create_error;
}

struct tb_ctx {
    tb_config cfg{};
    int PW = 0, PH = 0, col0 = 0, col1 = 0;
    long long n_local = 0;
    int W = 0, H = 0;
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t side = nullptr;            // noise of the next step under the previous splat (TB_OVERLAP)
    cudaEvent_t ev_state = nullptr;         // last write to the state buffers (main stream)
    cudaEvent_t ev_noise = nullptr;         // noise kernel done (side stream)
    cudaEvent_t ev_producer = nullptr;      // tb_wait_stream
    float2 *wander = nullptr;
    bool splat_since_step = false;          // something HBM-bound is queued that the noise can hide under
    bool overlap = false;

    float4 *buf[2] = {nullptr, nullptr};   // [0] current, [1] previous (src/particles.js:128)
    float4 *targets = nullptr;
    float4 *flow = nullptr;
    float4 *image = nullptr;
    size_t image_cap = 0;
    int IW = 0, IH = 0;
    float4 *layer = nullptr;
    size_t layer_cap = 0;
    uchar4 *frames = nullptr;               // optical flow: view + last, RGBA8
    size_t frames_cap = 0;
    float *line_attr = nullptr;             // flow lines: 9 attribute floats per vertex, then the shaded vertices, then the bbox
    fl::Vertex *line_verts = nullptr;
    int *line_bbox = nullptr;
    int line_cap = 0;

    // flow splat (tb_splat.cuh)
    PairEntry *pairs = nullptr;
    int n_pairs = 0;
    long long n_prims = 0;                 // local primitives = columns * n_pairs
    StripGeom geom{};
    int slab_prims = 0, n_slabs = 0;       // slabs of consecutive primitives (fixed per context)
    int slabs_per_seg = 1;
    uint32_t *slab_hist = nullptr;         // [n_slabs][kMaxBins] fragments per (slab, bin) -> per-bin exclusive scan over the slabs
    uint32_t *seg_total = nullptr;         // [2][kHistSegs][kMaxBins] the same summed over the slabs of a segment; the draws alternate
    int seg_parity = 0;
    uint32_t *bin_total = nullptr;         // [kMaxBins]
    uint32_t *bin_off = nullptr;           // [kMaxBins + 1]
    uint32_t *tickets = nullptr;           // [0] hist, [1] scatter, [2] fold work counters, [3] fold items, [4] a bin beyond 2^32
    uint32_t *items = nullptr;             // [kMaxBins] fold work items, longest bins first
    uint32_t *split_map = nullptr;         // [2][T] strip -> first bin | log2(bins) << 24; written by one draw's plan for the next
    uint32_t *bin_info = nullptr;          // [2][kMaxBins] bin -> strip | sub << 16 | log2(bins) << 24
    uint32_t *n_bins = nullptr;            // [2] device scalars
    int map_parity = 0;
    int fold_parity = 0;                   // the map the last collect used
    uint32_t split_at = 8192;              // the most fragments a bin should hold: the next draw splits a strip 2, 4, ... 256 ways to get there
    uint32_t share_at = 12288;             // fragments per bin above which 2 (4, 8) warps share the bin's fold
    PlanOut *d_plan = nullptr;
    PlanOut *h_plan = nullptr;             // pinned; valid once ev_plan has completed
    cudaEvent_t ev_plan = nullptr;
    Frag *bins = nullptr;                  // every fragment of a draw, binned by tile, draw order inside a bin
    uint32_t bin_cap = 0;
    // long bins folded in segments (tb_splat.cuh, PARITY B4)
    uint32_t seg_at = 0, seg_len = 8192;   // TB_SEG_AT (0: off, the default -- see DESIGN.md 4), TB_SEG_LEN
    bool seg_scaled = false;               // TB_SEG_SCALED=1: ... and not below 1/1024 of the draw
    uint32_t seg_out_cap = 4u << 20;       // float4 entries of segment results (64 MB)
    uint4 *seg_desc = nullptr;
    uint32_t *seg_of_bin = nullptr, *seg_cnt = nullptr;
    float4 *seg_out = nullptr;
    Frag *replay = nullptr;                // as large as `bins`: the records of the segments
    bool bin_fixed = false;                // the bin array is mapped by other ranks: it cannot grow without a reconnect
    // column-sharded run over peer memory (tb_owners.cuh): every rank maps every rank's bins, grid, totals table, flags
    bool owners_connected = false;
    int ow_rank = 0, ow_world = 1;
    uint32_t ow_epoch = 0;
    uint32_t *ow_flags = nullptr;          // [kOwnerPhases][kMaxBandRanks] the epoch each rank has reached
    uint32_t *ow_totals = nullptr;         // [kMaxBandRanks][kMaxBins] every rank's fragments per bin, written by the ranks themselves
    OwnerPeers ow_peers{};
    Frag *ow_bins[kMaxBandRanks] = {};
    float4 *ow_flow[kMaxBandRanks] = {};
    uint32_t ow_caps[kMaxBandRanks] = {};
    uint32_t *ow_scratch = nullptr;        // [4][kMaxBins]: bin_sum, scat_off, own_begin, own_count
    uint32_t *ow_last_all = nullptr;       // [kMaxBandRanks][W*H] every rank's last opaque primitive per texel, written by the ranks
    // opaque pruning (k_splat_opaque): off unless TB_PRUNE=1 -- exact, but on the bench workloads it removes only ~8 % of the
    // fragments (3 % of the lines are opaque, and not where the crowds are) and costs a pass
    uint32_t *last_local = nullptr;        // [W*H] this context's table
    uint32_t *last_global = nullptr;       // [W*H] sharded: maximum over the ranks
    uint32_t *prune_flags = nullptr;       // [2] bit 0: do not prune ([0] local, [1] over the ranks)
    int prune_mode = 0;                    // 0: never, 1: always
    bool pending = false;                  // a draw is queued whose capacity check has not been read yet
    int pending_stage = 0;                 // 1: collect only, 2: collect + fold, 3: the sharded draw
    float pending_time = 0.f;
    bool collected = false;
    float collect_time = 0.f;
    int n_sms = 148;
    int scatter_ctas = 0, fold_ctas = 0, hist_ctas = 0;

    // tb_step_streamed: column chunks over two copy streams (PCIe both ways at once) around the chunked logic pass
    static constexpr int kMaxChunks = 64;
    cudaStream_t h2d = nullptr, d2h = nullptr;
    cudaEvent_t ev_h2d[kMaxChunks] = {}, ev_int[kMaxChunks] = {}, ev_d2h[kMaxChunks] = {};
    cudaEvent_t ev_main = nullptr;         // the main stream at the start of a streamed step
    int chunks_inflight = 0;               // chunks of the last streamed step whose copies may still run
    float4 *spare = nullptr;               // third state buffer: the upload of a streamed step lands here while the previous
                                           // step's flow splat still reads the other two

    int *d_flag = nullptr;                 // device scratch flag
    int *h_flag = nullptr;                 // pinned
    bool targets_finite = true;

    // CUDA-event timing rings: [class][slot][begin/end]; class 0 = integrate, 1 = flow splat, 2 = noise (side stream)
    static constexpr int kTimingSlots = 512;
    cudaEvent_t ev_ring[3][kTimingSlots][2] = {};
    int64_t ev_count[3] = {0, 0, 0};
    // finer: the five kernels of the flow splat (hist, rows + plan, scatter, fold), same slots as class 1
    cudaEvent_t ev_stage[kTimingSlots][5] = {};
    bool stage_timing = false;

    tb_state state{};
    bool have_state = false;
    std::string err;
    int64_t launches = 0;
    int64_t last_frags = 0;
};

namespace {

int fail(tb_ctx *ctx, int code, const std::string &msg) {
    if (ctx) ctx->err = msg; else g_create_error = msg;
    return code;
}

#define TB_CUDA(ctx, expr)                                                                         \
    do {                                                                                           \
        cudaError_t e_ = (expr);                                                                   \
        if (e_ != cudaSuccess)                                                                     \
            return fail(ctx, TB_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e_));     \
    } while (0)

#define TB_REQUIRE(ctx, cond, msg)                                                                 \
    do {                                                                                           \
        if (!(cond)) return fail(ctx, TB_ERR_INVALID, std::string("tendrils-b200: ") + (msg));     \
    } while (0)

inline unsigned blocks_for(long long n, int threads) { return static_cast<unsigned>((n + threads - 1) / threads); }

// NEAREST/CLAMP_TO_EDGE texel of a normalised coordinate, float32 as the sampler computes it.
int host_texel(float u, int size) {
    volatile float prod = u * static_cast<float>(size);
    float f = std::floor(prod);
    if (!(f > 0.0f)) return 0;
    if (f > static_cast<float>(size - 1)) return size - 1;
    return static_cast<int>(f);
}

// The D6 table: Particles.generateLUT (src/particles.js:171-190) gives vertex j of a column
// uv.y = f32(j/(2PH-1)); stateAtFrame (src/state/state-at-frame.glsl:12-22) turns that into
// a texel row and a previous/current choice.  Pairs whose two vertices sample the same texel
// of the same buffer are zero-length lines and produce no fragments: dropped here.
std::vector<PairEntry> build_pairs(int PH) {
    std::vector<PairEntry> out;
    const int h = std::max(2 * PH, 2);
    const double inv = 1.0 / static_cast<double>(h - 1);
    auto vertex = [&](int j, int &row, bool &cur) {
        const float uvy = static_cast<float>(static_cast<double>(j) * inv);
        volatile float near_index = uvy * static_cast<float>(PH);
        const float fl = std::floor(near_index);
        volatile float off = near_index - fl;
        volatile float lookup = fl / static_cast<float>(PH);
        row = host_texel(lookup, PH);
        cur = off > 0.25f;
    };
    for (int k = 0; k < PH; ++k) {
        int ra, rb;
        bool ca, cb;
        vertex(2 * k, ra, ca);
        vertex(2 * k + 1, rb, cb);
        if (ra == rb && ca == cb) continue;
        PairEntry e;
        e.k = k;
        e.row_a = ra | (ca ? static_cast<int32_t>(0x80000000u) : 0);
        e.row_b = rb | (cb ? static_cast<int32_t>(0x80000000u) : 0);
        e.pad = 0;
        out.push_back(e);
    }
    return out;
}

// column sampled by vertex column i: uv.x = f32(i/(PW-1)) -> floor(uv.x*PW), must be i.
bool columns_are_identity(int PW) {
    const int w = std::max(PW, 2);
    const double inv = 1.0 / static_cast<double>(w - 1);
    for (int i = 0; i < PW; ++i)
        if (host_texel(static_cast<float>(static_cast<double>(i) * inv), PW) != i) return false;
    return true;
}

// Strips of the binning: 16 x 8 texels, grown (x first) until there are at most kMaxStrips of them.
StripGeom choose_geom(int W, int H) {
    StripGeom g{};
    g.W = W; g.H = H;
    g.sxl = 4; g.syl = 3;
    auto strips = [&]() {
        g.strips_x = (W + (1 << g.sxl) - 1) >> g.sxl;
        g.strips_y = (H + (1 << g.syl) - 1) >> g.syl;
        g.T = g.strips_x * g.strips_y;
        return g.T;
    };
    while (strips() > kMaxStrips) { if (g.sxl <= g.syl + 1) ++g.sxl; else ++g.syl; }
    return g;
}

int tiles_release(tb_ctx *c);

int alloc_flow(tb_ctx *c, int w, int h) {
    tiles_release(c);
    TB_REQUIRE(c, w >= 1 && h >= 1 && w <= 32768 && h <= 32768, "flow grid dimensions out of bounds");
    const StripGeom g = choose_geom(w, h);
    TB_REQUIRE(c, (1 << (g.sxl + g.syl)) <= kMaxStripTexels, "flow grid too large for the strip binning (at most 8192 strips of 512 texels)");
    cudaFree(c->flow); cudaFree(c->slab_hist); cudaFree(c->seg_total); cudaFree(c->bin_total); cudaFree(c->bin_off); cudaFree(c->items);
    cudaFree(c->split_map); cudaFree(c->bin_info); cudaFree(c->n_bins); cudaFree(c->last_local); cudaFree(c->last_global); cudaFree(c->prune_flags);
    cudaFree(c->seg_desc); cudaFree(c->seg_of_bin); cudaFree(c->seg_cnt); cudaFree(c->seg_out);
    c->seg_desc = nullptr; c->seg_of_bin = nullptr; c->seg_cnt = nullptr; c->seg_out = nullptr;
    c->last_local = nullptr; c->last_global = nullptr; c->prune_flags = nullptr;
    c->flow = nullptr; c->slab_hist = nullptr; c->seg_total = nullptr; c->bin_total = nullptr; c->bin_off = nullptr; c->items = nullptr;
    c->split_map = nullptr; c->bin_info = nullptr; c->n_bins = nullptr;
    c->W = w; c->H = h;
    c->geom = g;
    const size_t G = static_cast<size_t>(w) * h;
    const int T = g.T;
    TB_CUDA(c, cudaMalloc(&c->flow, G * sizeof(float4)));
    TB_CUDA(c, cudaMalloc(&c->slab_hist, static_cast<size_t>(kMaxBins) * std::max(c->n_slabs, 1) * sizeof(uint32_t)));
    TB_CUDA(c, cudaMalloc(&c->seg_total, 2 * static_cast<size_t>(kMaxBins) * kHistSegs * sizeof(uint32_t)));
    TB_CUDA(c, cudaMalloc(&c->bin_total, static_cast<size_t>(kMaxBins) * sizeof(uint32_t)));
    TB_CUDA(c, cudaMalloc(&c->bin_off, static_cast<size_t>(kMaxBins + 1) * sizeof(uint32_t)));
    TB_CUDA(c, cudaMalloc(&c->items, 16 * static_cast<size_t>(kMaxBins) * sizeof(uint32_t)));
    if (c->seg_at) {
        TB_CUDA(c, cudaMalloc(&c->seg_desc, static_cast<size_t>(kMaxBins) * sizeof(uint4)));
        TB_CUDA(c, cudaMalloc(&c->seg_of_bin, static_cast<size_t>(kMaxBins) * sizeof(uint32_t)));
        TB_CUDA(c, cudaMalloc(&c->seg_cnt, 16 * static_cast<size_t>(kMaxBins) * sizeof(uint32_t)));
        TB_CUDA(c, cudaMalloc(&c->seg_out, static_cast<size_t>(c->seg_out_cap) * sizeof(float4)));
    }
    TB_CUDA(c, cudaMalloc(&c->split_map, 2 * static_cast<size_t>(T) * sizeof(uint32_t)));
    TB_CUDA(c, cudaMalloc(&c->bin_info, 2 * static_cast<size_t>(kMaxBins) * sizeof(uint32_t)));
    TB_CUDA(c, cudaMalloc(&c->n_bins, 2 * sizeof(uint32_t)));
    TB_CUDA(c, cudaMalloc(&c->last_local, G * sizeof(uint32_t)));
    TB_CUDA(c, cudaMalloc(&c->last_global, G * sizeof(uint32_t)));
    TB_CUDA(c, cudaMalloc(&c->prune_flags, 2 * sizeof(uint32_t)));
    TB_CUDA(c, cudaMemsetAsync(c->flow, 0, G * sizeof(float4), c->stream));
    TB_CUDA(c, cudaMemsetAsync(c->seg_total, 0, 2 * static_cast<size_t>(kMaxBins) * kHistSegs * sizeof(uint32_t), c->stream));
    TB_CUDA(c, cudaMemsetAsync(c->bin_total, 0, static_cast<size_t>(kMaxBins) * sizeof(uint32_t), c->stream));
    c->seg_parity = 0;
    c->map_parity = 0;
    k_splat_map_identity<<<blocks_for(T, 256), 256, 0, c->stream>>>(T, c->split_map, c->bin_info, c->n_bins);     // one bin per strip to begin with
    if (cudaGetLastError() != cudaSuccess) return fail(c, TB_ERR_CUDA, "k_splat_map_identity failed to launch");
    // persistent grids of the splat kernels
    TB_CUDA(c, cudaFuncSetAttribute(k_splat_hist, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kHistSmemBytes)));
    TB_CUDA(c, cudaFuncSetAttribute(k_splat_scatter, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kScatterSmemBytes)));
    TB_CUDA(c, cudaFuncSetAttribute(k_splat_fold, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(fold_smem_bytes(1 << (c->geom.sxl + c->geom.syl)))));
    TB_CUDA(c, cudaFuncSetAttribute(k_splat_mend, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(fold_smem_bytes(1 << (c->geom.sxl + c->geom.syl)))));
    int per_sm = 0;
    TB_CUDA(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_splat_scatter, kEmitThreads, kScatterSmemBytes));
    c->scatter_ctas = std::max(1, per_sm) * c->n_sms;
    TB_CUDA(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_splat_fold, kFoldThreads, fold_smem_bytes(1 << (c->geom.sxl + c->geom.syl))));
    c->fold_ctas = std::max(1, per_sm) * c->n_sms;
    TB_CUDA(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_splat_hist, kHistThreads, kHistSmemBytes));
    c->hist_ctas = std::max(1, per_sm) * c->n_sms;
    // experiment knobs: resident CTAs per SM of the splat kernels (fewer leave room for the noise launch of TB_OVERLAP)
    if (const char *e = std::getenv("TB_SCATTER_CTAS")) c->scatter_ctas = std::min(c->scatter_ctas, std::max(1, std::atoi(e)) * c->n_sms);
    if (const char *e = std::getenv("TB_FOLD_CTAS")) c->fold_ctas = std::min(c->fold_ctas, std::max(1, std::atoi(e)) * c->n_sms);
    if (const char *e = std::getenv("TB_HIST_CTAS")) c->hist_ctas = std::min(c->hist_ctas, std::max(1, std::atoi(e)) * c->n_sms);
    c->collected = false;
    c->pending = false;
    return TB_OK;
}

int ensure_bin_cap(tb_ctx *c, uint64_t need) {
    if (need <= c->bin_cap) return TB_OK;
    if (need >= (1ull << 31))
        return fail(c, TB_ERR_OVERFLOW, "tendrils-b200: flow splat: " + std::to_string(need) + " fragments in one draw (limit 2^31)");
    if (c->bin_fixed)
        return fail(c, TB_ERR_OVERFLOW, "tendrils-b200: flow splat: " + std::to_string(need) + " fragments exceed the " +
                    std::to_string(c->bin_cap) + " reserved by tb_owners_export (the bins are mapped by the other ranks); "
                    "export with a larger reserve and reconnect");
    uint64_t cap = std::max<uint64_t>(need + need / 4, 1u << 16);
    if (cap >= (1ull << 31)) cap = (1ull << 31) - 1;
    TB_CUDA(c, cudaStreamSynchronize(c->stream));
    cudaFree(c->bins); cudaFree(c->replay);
    c->bins = nullptr; c->replay = nullptr; c->bin_cap = 0;
    TB_CUDA(c, cudaMalloc(&c->bins, cap * sizeof(Frag)));
    if (c->seg_at) TB_CUDA(c, cudaMalloc(&c->replay, cap * sizeof(Frag)));
    c->bin_cap = static_cast<uint32_t>(cap);
    return TB_OK;
}

// Image / frame arguments may live on the host or (unified addressing) on this device, e.g. a decoded video frame
// or a torch tensor: device-resident inputs are used in stream order, with no copy-back synchronisation.
bool is_device_pointer(const void *p) {
    cudaPointerAttributes a{};
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

int check_launch(tb_ctx *c, const char *what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(c, TB_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
    c->launches += 1;
    return TB_OK;
}

PrimSource prim_source(tb_ctx *c) {
    PrimSource S{};
    S.cur = c->buf[0];
    S.prev = c->buf[1];
    S.pairs = c->pairs;
    S.n_pairs = c->n_pairs;
    S.PH = c->PH;
    S.n_prims = c->n_prims;
    return S;
}

// Collect: rasterise this context's primitives into per-tile bins, draw order inside every bin.  Four launches, nothing
// comes back to the host on the way: the plan (fragment total, overflow flag) is copied out behind the kernels and read
// by resolve_pending() at the next API call.
BinMap bin_map(tb_ctx *c, int parity) {
    BinMap M{};
    M.map = c->split_map + static_cast<size_t>(parity) * c->geom.T;
    M.n_bins = c->n_bins + parity;
    M.lS = c->geom.sxl + c->geom.syl;
    return M;
}

// the table of last opaque primitives of this context's own primitives (k_splat_opaque)
int launch_opaque(tb_ctx *c, float time) {
    const size_t G = static_cast<size_t>(c->W) * c->H;
    TB_CUDA(c, cudaMemsetAsync(c->last_local, 0, G * sizeof(uint32_t), c->stream));
    TB_CUDA(c, cudaMemsetAsync(c->prune_flags, 0, 2 * sizeof(uint32_t), c->stream));
    if (c->n_prims <= 0) return TB_OK;
    OpaqueArgs OA{};
    OA.src = prim_source(c);
    OA.g = c->geom;
    OA.vsx = c->state.viewSize[0]; OA.vsy = c->state.viewSize[1];
    OA.speedLimit = c->state.speedLimit;
    OA.time = time;
    OA.prim_base = static_cast<long long>(c->col0) * c->n_pairs;
    OA.last = c->last_local;
    OA.flags = c->prune_flags;
    k_splat_opaque<<<c->n_sms * 8, 256, 0, c->stream>>>(OA);
    return check_launch(c, "k_splat_opaque");
}

Prune no_prune() { return Prune{nullptr, nullptr, 0}; }

// fragments per bin of this context's primitives (k_splat_hist + k_splat_rows) under split map `mp`
int launch_count(tb_ctx *c, int mp, const Prune &prune) {
    cudaEvent_t *stage = c->ev_stage[c->ev_count[1] % tb_ctx::kTimingSlots];
    if (c->n_prims > 0) {
        uint32_t *seg_now = c->seg_total + static_cast<size_t>(c->seg_parity) * kHistSegs * kMaxBins;
        uint32_t *seg_next = c->seg_total + static_cast<size_t>(c->seg_parity ^ 1) * kHistSegs * kMaxBins;
        c->seg_parity ^= 1;
        HistArgs HA{};
        HA.src = prim_source(c);
        HA.g = c->geom;
        HA.bm = bin_map(c, mp);
        HA.prune = prune;
        HA.vsx = c->state.viewSize[0]; HA.vsy = c->state.viewSize[1];
        HA.slab_prims = c->slab_prims; HA.n_slabs = c->n_slabs; HA.slabs_per_seg = c->slabs_per_seg;
        HA.slab_hist = c->slab_hist;
        HA.seg_total = seg_now;
        HA.ticket = c->tickets + 0;
        k_splat_hist<<<std::min(c->n_slabs, c->hist_ctas), kHistThreads, kHistSmemBytes, c->stream>>>(HA);
        if (int r = check_launch(c, "k_splat_hist")) return r;
        if (c->stage_timing) cudaEventRecord(stage[1], c->stream);
        k_splat_rows<<<dim3(kMaxBins / 256, kHistSegs), 256, 0, c->stream>>>(c->slab_hist, seg_now, seg_next, c->n_bins + mp, c->n_slabs,
                                                                             c->slabs_per_seg, c->bin_total, c->tickets + 4);
        if (int r = check_launch(c, "k_splat_rows")) return r;
    } else {
        TB_CUDA(c, cudaMemsetAsync(c->bin_total, 0, static_cast<size_t>(kMaxBins) * sizeof(uint32_t), c->stream));
    }
    return TB_OK;
}

int launch_scatter(tb_ctx *c, float time, int mp, const uint32_t *bin_off, Frag *const *bins, int n_ranks, const Prune &prune) {
    if (c->n_prims <= 0) return TB_OK;
    ScatterArgs SA{};
    SA.src = prim_source(c);
    SA.g = c->geom;
    SA.bm = bin_map(c, mp);
    SA.prune = prune;
    SA.vsx = c->state.viewSize[0]; SA.vsy = c->state.viewSize[1];
    SA.speedLimit = c->state.speedLimit;
    SA.time = time;
    SA.slab_prims = c->slab_prims; SA.n_slabs = c->n_slabs;
    SA.slab_hist = c->slab_hist;
    SA.bin_off = bin_off;
    SA.plan = c->d_plan;
    SA.ticket = c->tickets + 1;
    for (int r = 0; r < n_ranks; ++r) SA.bins[r] = bins[r];
    SA.n_ranks = n_ranks;
    k_splat_scatter<<<std::min(c->n_slabs, c->scatter_ctas), kEmitThreads, kScatterSmemBytes, c->stream>>>(SA);
    return check_launch(c, "k_splat_scatter");
}

SegPlan seg_plan(tb_ctx *c) {
    SegPlan G{};
    G.seg_at = c->replay ? c->seg_at : 0u;
    G.scaled = c->seg_scaled ? 1u : 0u;
    G.seg_len = std::max<uint32_t>(c->seg_len, 64u);
    G.out_cap = c->seg_out_cap;
    G.desc = c->seg_desc;
    G.of_bin = c->seg_of_bin;
    return G;
}

// the fold proper, then the join of the bins that were folded in segments
int launch_fold_kernels(tb_ctx *c, FoldArgs &FA) {
    FA.seg_desc = c->seg_desc;
    FA.seg_of_bin = c->seg_of_bin;
    FA.seg_out = c->seg_out;
    FA.seg_cnt = c->seg_cnt;
    FA.replay = c->replay;
    FA.n_seg = c->tickets + 5;
    FA.seg_ticket = c->tickets + 6;
    const size_t smem = fold_smem_bytes(1 << (c->geom.sxl + c->geom.syl));
    k_splat_fold<<<c->fold_ctas, kFoldThreads, smem, c->stream>>>(FA);
    if (int r = check_launch(c, "k_splat_fold")) return r;
    if (c->seg_at && c->replay) {
        k_splat_mend<<<std::min(c->fold_ctas, 2 * c->n_sms), kFoldThreads, smem, c->stream>>>(FA);
        if (int r = check_launch(c, "k_splat_mend")) return r;
    }
    return TB_OK;
}

int launch_collect(tb_ctx *c, float time) {
    const int T = c->geom.T;
    cudaEvent_t *stage = c->ev_stage[c->ev_count[1] % tb_ctx::kTimingSlots];
    const int mp = c->map_parity;                  // this draw's split map; its plan writes the other one for the next draw
    c->map_parity ^= 1;
    Prune prune = no_prune();
    if (c->prune_mode == 1) {
        if (int r = launch_opaque(c, time)) return r;
        prune = Prune{c->last_local, c->prune_flags, static_cast<long long>(c->col0) * c->n_pairs};
    }
    if (int r = launch_count(c, mp, prune)) return r;
    PlanArgs PA{};
    PA.T = T;
    PA.lS = c->geom.sxl + c->geom.syl;
    PA.bm = bin_map(c, mp);
    PA.bin_info = c->bin_info + static_cast<size_t>(mp) * kMaxBins;
    PA.bin_total = c->bin_total;
    PA.bin_off = c->bin_off;
    PA.items = c->items;
    PA.cap = c->bin_cap;
    PA.split_at = c->split_at;
    PA.share_at = c->share_at;
    PA.seg = seg_plan(c);
    PA.too_many = c->tickets + 4;
    PA.tickets = c->tickets;
    PA.map_next = c->split_map + static_cast<size_t>(mp ^ 1) * T;
    PA.bin_info_next = c->bin_info + static_cast<size_t>(mp ^ 1) * kMaxBins;
    PA.n_bins_next = c->n_bins + (mp ^ 1);
    PA.out = c->d_plan;
    k_splat_plan<<<1, kPlanThreads, 0, c->stream>>>(PA);
    if (int r = check_launch(c, "k_splat_plan")) return r;
    TB_CUDA(c, cudaMemcpyAsync(c->h_plan, c->d_plan, sizeof(PlanOut), cudaMemcpyDeviceToHost, c->stream));
    TB_CUDA(c, cudaEventRecord(c->ev_plan, c->stream));
    if (c->stage_timing) cudaEventRecord(stage[2], c->stream);
    Frag *one[1] = {c->bins};
    if (int r = launch_scatter(c, time, mp, c->bin_off, one, 1, prune)) return r;
    if (c->stage_timing) cudaEventRecord(stage[3], c->stream);
    c->fold_parity = mp;
    return TB_OK;
}

int launch_fold(tb_ctx *c, float time) {
    FoldArgs FA{};
    FA.g = c->geom;
    FA.time = time;
    FA.bins = c->bins;
    FA.bin_off = c->bin_off;
    FA.bin_count = nullptr;
    FA.bin_info = c->bin_info + static_cast<size_t>(c->fold_parity) * kMaxBins;
    FA.items = c->items;
    FA.n_items = c->tickets + 3;
    FA.ticket = c->tickets + 2;
    FA.flow[0] = c->flow;
    FA.n_flow = 1;
    return launch_fold_kernels(c, FA);
}

int queue_owners(tb_ctx *c, float time);

int queue_splat(tb_ctx *c, float time, int stage) {
    if (stage == 3) return queue_owners(c, time);
    cudaEvent_t *ev = c->ev_ring[1][c->ev_count[1] % tb_ctx::kTimingSlots];
    if (stage >= 1) {
        TB_CUDA(c, cudaEventRecord(ev[0], c->stream));
        if (c->stage_timing) cudaEventRecord(c->ev_stage[c->ev_count[1] % tb_ctx::kTimingSlots][0], c->stream);
        if (int r = launch_collect(c, time)) return r;
    }
    if (stage == 2) {
        if (int r = launch_fold(c, time)) return r;
        if (c->stage_timing) cudaEventRecord(c->ev_stage[c->ev_count[1] % tb_ctx::kTimingSlots][4], c->stream);
        TB_CUDA(c, cudaEventRecord(ev[1], c->stream));
    }
    return TB_OK;
}

// The capacity check of the last queued draw.  Its kernels did nothing if the fragments did not fit (the plan kernel
// saw that on the device): grow the bin array and queue the draw again -- the state buffers and the grid are as they
// were, because every entry point comes through here before it touches either.
int resolve_pending(tb_ctx *c) {
    if (!c->pending) return TB_OK;
    for (int attempt = 0; attempt < 4; ++attempt) {
        TB_CUDA(c, cudaEventSynchronize(c->ev_plan));
        c->last_frags = static_cast<int64_t>(c->h_plan->total);
        if (!c->h_plan->overflow) {
            c->pending = false;
            return TB_OK;
        }
        if (int r = ensure_bin_cap(c, c->h_plan->needed)) { c->pending = false; c->collected = false; return r; }
        if (c->pending_stage >= 2) c->ev_count[1] -= 1;              // the retry re-records the same timing slot
        const int r = queue_splat(c, c->pending_time, c->pending_stage);
        if (c->pending_stage >= 2) c->ev_count[1] += 1;
        if (r) { c->pending = false; return r; }
    }
    c->pending = false;
    return fail(c, TB_ERR_OVERFLOW, "tendrils-b200: flow splat: the fragment bins could not be sized");
}

// Make the main stream wait for the copies of the last streamed step (it owns the state buffers again afterwards).
int join_copies(tb_ctx *c) {
    for (int i = 0; i < c->chunks_inflight; ++i) {
        TB_CUDA(c, cudaStreamWaitEvent(c->stream, c->ev_h2d[i], 0));
        TB_CUDA(c, cudaStreamWaitEvent(c->stream, c->ev_d2h[i], 0));
    }
    c->chunks_inflight = 0;
    return TB_OK;
}

int collect(tb_ctx *c, float time) {
    TB_REQUIRE(c, c->have_state, "tb_set_state must be called before the flow splat");
    if (int r = resolve_pending(c)) return r;
    c->collect_time = time;
    c->collected = false;
    c->splat_since_step = true;
    c->pending_time = time;
    c->pending_stage = 1;
    if (int r = queue_splat(c, time, 1)) return r;
    c->pending = true;
    c->collected = true;
    return TB_OK;
}

int fold(tb_ctx *c) {
    TB_REQUIRE(c, c->collected, "tb_splat_fold without a preceding tb_splat_collect");
    if (int r = resolve_pending(c)) return r;          // the split form is not on the single-GPU hot path: check the collect now
    if (int r = launch_fold(c, c->collect_time)) return r;
    TB_CUDA(c, cudaEventRecord(c->ev_ring[1][c->ev_count[1] % tb_ctx::kTimingSlots][1], c->stream));
    if (c->stage_timing) cudaEventRecord(c->ev_stage[c->ev_count[1] % tb_ctx::kTimingSlots][4], c->stream);
    c->ev_count[1] += 1;
    c->collected = false;
    return TB_OK;
}

// tb_splat_flow: collect + fold queued back to back, checked lazily.
int splat(tb_ctx *c, float time) {
    TB_REQUIRE(c, c->have_state, "tb_set_state must be called before the flow splat");
    if (int r = resolve_pending(c)) return r;
    c->collect_time = time;
    c->collected = false;
    c->splat_since_step = true;
    c->pending_time = time;
    c->pending_stage = 2;
    if (int r = queue_splat(c, time, 2)) return r;
    c->pending = true;
    c->ev_count[1] += 1;
    return TB_OK;
}

// spawnShader target handling (src/particles.js:123-130): no buffer -> rotate and write
// buffers[0]; explicit targets FBO -> no rotation.  `particles` is buffers[1] either way.
float4 *spawn_out(tb_ctx *c, tb_target target) {
    if (target == TB_TARGET_TARGETS) return c->targets;
    std::swap(c->buf[0], c->buf[1]);
    return c->buf[0];
}

int after_targets_write(tb_ctx *c, tb_target target) {
    if (target != TB_TARGET_TARGETS) {
        TB_CUDA(c, cudaEventRecord(c->ev_state, c->stream));     // the state buffers changed
        return TB_OK;
    }
    TB_CUDA(c, cudaMemsetAsync(c->d_flag, 0, sizeof(int), c->stream));
    k_check_finite<<<blocks_for(c->n_local, 256), 256, 0, c->stream>>>(c->targets, c->n_local, c->d_flag);
    if (int r = check_launch(c, "k_check_finite")) return r;
    TB_CUDA(c, cudaMemcpyAsync(c->h_flag, c->d_flag, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    TB_CUDA(c, cudaStreamSynchronize(c->stream));
    c->targets_finite = (*c->h_flag == 0);
    return TB_OK;
}

bool tame(float v, float lim) { return std::isfinite(v) && std::fabs(v) < lim; }

int tiles_release(tb_ctx *c) {
    for (int j = 0; j < kMaxBandRanks; ++j) {
        if (c->owners_connected && j != c->ow_rank) {
            if (c->ow_bins[j]) cudaIpcCloseMemHandle(c->ow_bins[j]);
            if (c->ow_flow[j]) cudaIpcCloseMemHandle(c->ow_flow[j]);
            if (c->ow_peers.totals[j]) cudaIpcCloseMemHandle(c->ow_peers.totals[j]);
            if (c->ow_peers.flags[j]) cudaIpcCloseMemHandle(c->ow_peers.flags[j]);
            if (c->ow_peers.last[j]) cudaIpcCloseMemHandle(c->ow_peers.last[j]);
        }
        c->ow_bins[j] = nullptr; c->ow_flow[j] = nullptr;
        c->ow_peers.totals[j] = nullptr; c->ow_peers.flags[j] = nullptr; c->ow_peers.last[j] = nullptr;
    }
    c->owners_connected = false;
    c->bin_fixed = false;
    return TB_OK;
}

// The sharded draw (tb_owners.cuh).  Every rank queues the same sequence; the three barriers are kernels on the stream.
int queue_owners(tb_ctx *c, float time) {
    const int T = c->geom.T, P = c->ow_world;
    cudaEvent_t *ev = c->ev_ring[1][c->ev_count[1] % tb_ctx::kTimingSlots];
    TB_CUDA(c, cudaEventRecord(ev[0], c->stream));
    const uint32_t epoch = ++c->ow_epoch;
    auto barrier = [&](int phase) -> int {
        k_owners_barrier<<<1, 32, 0, c->stream>>>(c->ow_peers, c->ow_flags, phase, epoch);
        return check_launch(c, "k_owners_barrier");
    };
    // TB_OWNERS_DEBUG=<epoch>: time the phases of that draw on every rank (diagnostics, stderr)
    static const int dbg_epoch = std::getenv("TB_OWNERS_DEBUG") ? std::atoi(std::getenv("TB_OWNERS_DEBUG")) : -1;
    const bool dbg = static_cast<int>(epoch) == dbg_epoch;
    static cudaEvent_t dbg_ev[12];
    int dbg_n = 0;
    auto mark = [&]() { if (dbg) { cudaEventCreate(&dbg_ev[dbg_n]); cudaEventRecord(dbg_ev[dbg_n++], c->stream); } };
    const int mp = c->map_parity;
    c->map_parity ^= 1;
    mark();
    Prune prune = no_prune();
    if (c->prune_mode != 0) {
        // every rank's table of last opaque primitives into every rank, then the maximum: what a LATER rank overwrites is
        // not even sent
        const int G = c->W * c->H;
        if (int r = launch_opaque(c, time)) return r;
        k_owners_push_last<<<blocks_for(G, 256), 256, 0, c->stream>>>(c->last_local, c->prune_flags, G, c->ow_peers);
        if (int r = check_launch(c, "k_owners_push_last")) return r;
        if (int r = barrier(3)) return r;
        k_owners_last_max<<<blocks_for(G, 256), 256, 0, c->stream>>>(c->ow_last_all, c->ow_flags, P, G, c->last_global, c->prune_flags + 1);
        if (int r = check_launch(c, "k_owners_last_max")) return r;
        prune = Prune{c->last_global, c->prune_flags + 1, static_cast<long long>(c->col0) * c->n_pairs};
    }
    if (int r = launch_count(c, mp, prune)) return r;
    mark();
    k_owners_share<<<kMaxBins / 256, 256, 0, c->stream>>>(c->bin_total, c->n_bins + mp, c->ow_peers);
    if (int r = check_launch(c, "k_owners_share")) return r;
    if (int r = barrier(0)) return r;                      // every rank's totals are in every table
    mark();
    uint32_t *bin_sum = c->ow_scratch, *scat_off = bin_sum + kMaxBins, *own_begin = scat_off + kMaxBins, *own_count = own_begin + kMaxBins;
    OwnerPlanArgs PA{};
    PA.T = T;
    PA.lS = c->geom.sxl + c->geom.syl;
    PA.bm = bin_map(c, mp);
    PA.bin_info = c->bin_info + static_cast<size_t>(mp) * kMaxBins;
    PA.totals = c->ow_totals;
    PA.n = P; PA.me = c->ow_rank;
    for (int r = 0; r < P; ++r) PA.caps[r] = c->ow_caps[r];
    PA.bin_sum = bin_sum; PA.scat_off = scat_off; PA.own_begin = own_begin; PA.own_count = own_count;
    PA.items = c->items;
    PA.split_at = c->split_at;
    PA.share_at = c->share_at;
    PA.seg = seg_plan(c);
    PA.tickets = c->tickets;
    PA.map_next = c->split_map + static_cast<size_t>(mp ^ 1) * T;
    PA.bin_info_next = c->bin_info + static_cast<size_t>(mp ^ 1) * kMaxBins;
    PA.n_bins_next = c->n_bins + (mp ^ 1);
    PA.out = c->d_plan;
    k_owners_plan<<<1, kPlanThreads, 0, c->stream>>>(PA);
    if (int r = check_launch(c, "k_owners_plan")) return r;
    TB_CUDA(c, cudaMemcpyAsync(c->h_plan, c->d_plan, sizeof(PlanOut), cudaMemcpyDeviceToHost, c->stream));
    TB_CUDA(c, cudaEventRecord(c->ev_plan, c->stream));
    mark();
    if (int r = launch_scatter(c, time, mp, scat_off, c->ow_bins, P, prune)) return r;
    mark();
    if (int r = barrier(1)) return r;                      // every rank's fragments are in their owners' bins
    mark();
    c->fold_parity = mp;
    FoldArgs FA{};
    FA.g = c->geom;
    FA.time = time;
    FA.bins = c->bins;
    FA.bin_off = own_begin;
    FA.bin_count = own_count;
    FA.bin_info = c->bin_info + static_cast<size_t>(mp) * kMaxBins;
    FA.items = c->items;
    FA.n_items = c->tickets + 3;
    FA.ticket = c->tickets + 2;
    FA.flow[0] = c->flow;                                  // read here; the finished texels go to every rank's grid
    int nf = 1;
    for (int r = 0; r < P; ++r)
        if (r != c->ow_rank) FA.flow[nf++] = c->ow_flow[r];
    FA.n_flow = nf;
    if (int r = launch_fold_kernels(c, FA)) return r;
    mark();
    if (int r = barrier(2)) return r;                      // every grid is complete
    mark();
    if (dbg) {
        cudaStreamSynchronize(c->stream);
        float ms[8] = {};
        for (int k = 0; k + 1 < dbg_n; ++k) cudaEventElapsedTime(&ms[k], dbg_ev[k], dbg_ev[k + 1]);
        std::fprintf(stderr, "[owners dbg] rank %d: hist+rows %.0f us, share+barrier0 %.0f us, plan %.0f us, scatter %.0f us, barrier1 %.0f us, "
                             "fold %.0f us, barrier2 %.0f us (own frags %llu)\n",
                     c->ow_rank, 1e3 * ms[0], 1e3 * ms[1], 1e3 * ms[2], 1e3 * ms[3], 1e3 * ms[4], 1e3 * ms[5], 1e3 * ms[6],
                     static_cast<unsigned long long>(c->h_plan->total));
        for (int k = 0; k < dbg_n; ++k) cudaEventDestroy(dbg_ev[k]);
    }
    TB_CUDA(c, cudaEventRecord(ev[1], c->stream));
    return TB_OK;
}

}  // namespace

extern "C" {

int tb_abi_version(void) { return TB_ABI_VERSION; }

const char *tb_last_error(const tb_ctx *ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

int tb_create(const tb_config *cfg, tb_ctx **out) {
    if (!cfg || !out) return fail(nullptr, TB_ERR_INVALID, "tendrils-b200: null argument");
    *out = nullptr;
    const int PW = cfg->particles_w, PH = cfg->particles_h;
    if (PW < 2 || PH < 2 || static_cast<long long>(PW) * PH >= (1LL << 32))
        return fail(nullptr, TB_ERR_INVALID, "tendrils-b200: Texture dimensions are out of bounds");
    const int col0 = cfg->col0, col1 = cfg->col1 == 0 ? PW : cfg->col1;
    if (col0 < 0 || col1 > PW || col0 >= col1)
        return fail(nullptr, TB_ERR_INVALID, "tendrils-b200: invalid column range");
    if (col1 - col0 > 65535 || PH > (1 << 24))
        return fail(nullptr, TB_ERR_INVALID, "tendrils-b200: a context holds at most 65535 columns of at most 2^24 rows");
    if (!columns_are_identity(PW))
        return fail(nullptr, TB_ERR_UNSUPPORTED, "tendrils-b200: vertex LUT columns do not map 1:1 for this width");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(nullptr, TB_ERR_CUDA, std::string("tendrils-b200: no CUDA device (there is no CPU fallback): ") +
                                              cudaGetErrorString(e));
    if (cfg->device < 0 || cfg->device >= ndev) return fail(nullptr, TB_ERR_INVALID, "tendrils-b200: bad device ordinal");
    e = cudaSetDevice(cfg->device);
    if (e != cudaSuccess) return fail(nullptr, TB_ERR_CUDA, cudaGetErrorString(e));

    tb_ctx *c = new tb_ctx();
    c->cfg = *cfg;
    c->PW = PW; c->PH = PH; c->col0 = col0; c->col1 = col1;
    c->n_local = static_cast<long long>(col1 - col0) * PH;
    c->device = cfg->device;
    auto bail = [&](int code) {
        g_create_error = c->err;
        tb_destroy(c);
        return code;
    };
#define TB_TRY(expr) do { cudaError_t e2_ = (expr); if (e2_ != cudaSuccess) { c->err = std::string(#expr) + ": " + cudaGetErrorString(e2_); return bail(TB_ERR_CUDA); } } while (0)
    int prio_lo = 0, prio_hi = 0;
    TB_TRY(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
    TB_TRY(cudaStreamCreateWithPriority(&c->stream, cudaStreamNonBlocking, prio_hi));
    {
        const char *sp = std::getenv("TB_SIDE_PRIO");     // experiment knob: "equal" | "high" (default low)
        int prio = prio_lo;
        if (sp && std::string(sp) == "equal") prio = prio_hi;
        if (sp && std::string(sp) == "high") { prio = prio_hi; prio_hi = prio_lo; }
        if (sp && std::string(sp) == "high") {
            TB_TRY(cudaStreamDestroy(c->stream));
            TB_TRY(cudaStreamCreateWithPriority(&c->stream, cudaStreamNonBlocking, prio_lo));
        }
        TB_TRY(cudaStreamCreateWithPriority(&c->side, cudaStreamNonBlocking, prio));
    }
    TB_TRY(cudaEventCreateWithFlags(&c->ev_state, cudaEventDisableTiming));
    TB_TRY(cudaEventCreateWithFlags(&c->ev_noise, cudaEventDisableTiming));
    // On by default (TB_OVERLAP=0 or tb_set_overlap turn it off): the two simplex noises of the NEXT step do not read the flow
    // grid, so they run on the low-priority side stream under the flow splat, whose kernels leave issue slots free
    // (measured at cfg3: 1.50 -> 1.31 ms per step).  The noise launch is time-sliced then: its own duration says nothing
    // about the kernel -- for roofline numbers measure with the overlap off.
    c->overlap = !(std::getenv("TB_OVERLAP") && std::atoi(std::getenv("TB_OVERLAP")) == 0);
    c->stage_timing = std::getenv("TB_STAGE_TIMING") != nullptr;
    if (const char *e = std::getenv("TB_SPLIT_AT")) c->split_at = static_cast<uint32_t>(std::max(64, std::atoi(e)));
    if (const char *e = std::getenv("TB_PRUNE")) c->prune_mode = std::atoi(e) != 0 ? 1 : 0;
    if (const char *e = std::getenv("TB_SHARE_AT")) c->share_at = static_cast<uint32_t>(std::max(64, std::atoi(e)));
    if (const char *e = std::getenv("TB_SEG_AT")) c->seg_at = static_cast<uint32_t>(std::max(0, std::atoi(e)));
    if (const char *e = std::getenv("TB_SEG_SCALED")) c->seg_scaled = std::atoi(e) != 0;
    if (const char *e = std::getenv("TB_SEG_LEN")) c->seg_len = static_cast<uint32_t>(std::max(64, std::atoi(e)));
    TB_TRY(cudaDeviceGetAttribute(&c->n_sms, cudaDevAttrMultiProcessorCount, c->device));
    const size_t bytes = static_cast<size_t>(c->n_local) * sizeof(float4);
    TB_TRY(cudaMalloc(&c->buf[0], bytes));
    TB_TRY(cudaMalloc(&c->buf[1], bytes));
    TB_TRY(cudaMalloc(&c->targets, bytes));
    TB_TRY(cudaMalloc(&c->wander, static_cast<size_t>(c->n_local) * sizeof(float2)));
    TB_TRY(cudaMemsetAsync(c->targets, 0, bytes, c->stream));       // FBO textures start zeroed
    TB_TRY(cudaMemsetAsync(c->buf[0], 0, bytes, c->stream));
    TB_TRY(cudaMemsetAsync(c->buf[1], 0, bytes, c->stream));
    TB_TRY(cudaMalloc(&c->d_flag, sizeof(int)));
    TB_TRY(cudaMallocHost(&c->h_flag, sizeof(int)));
    TB_TRY(cudaMallocHost(&c->h_plan, sizeof(PlanOut)));
    std::memset(c->h_plan, 0, sizeof(PlanOut));
    TB_TRY(cudaMalloc(&c->d_plan, sizeof(PlanOut)));
    TB_TRY(cudaMemsetAsync(c->d_plan, 0, sizeof(PlanOut), c->stream));
    TB_TRY(cudaMalloc(&c->tickets, 8 * sizeof(uint32_t)));
    TB_TRY(cudaMemsetAsync(c->tickets, 0, 8 * sizeof(uint32_t), c->stream));
    TB_TRY(cudaEventCreateWithFlags(&c->ev_plan, cudaEventDisableTiming));
    if (c->stage_timing)
        for (int i = 0; i < tb_ctx::kTimingSlots; ++i)
            for (int j = 0; j < 5; ++j) TB_TRY(cudaEventCreate(&c->ev_stage[i][j]));
    for (int k = 0; k < 3; ++k)
        for (int i = 0; i < tb_ctx::kTimingSlots; ++i)
            for (int j = 0; j < 2; ++j) TB_TRY(cudaEventCreate(&c->ev_ring[k][i][j]));
    const std::vector<PairEntry> pairs = build_pairs(PH);
    c->n_pairs = static_cast<int>(pairs.size());
    if (c->n_pairs) {
        TB_TRY(cudaMalloc(&c->pairs, pairs.size() * sizeof(PairEntry)));
        TB_TRY(cudaMemcpyAsync(c->pairs, pairs.data(), pairs.size() * sizeof(PairEntry), cudaMemcpyHostToDevice, c->stream));
        TB_TRY(cudaStreamSynchronize(c->stream));
    }
    c->n_prims = static_cast<long long>(col1 - col0) * c->n_pairs;
    if (c->n_prims >= (1LL << 31)) { c->err = "tendrils-b200: too many primitives per context"; return bail(TB_ERR_INVALID); }
    // slabs of consecutive primitives: the unit of work of the count and emit passes (about eight per SM)
    {
        const long long want = (c->n_prims + 6LL * c->n_sms - 1) / (6LL * c->n_sms);      // two slabs per resident CTA of the count and emit kernels
        c->slab_prims = static_cast<int>(std::max<long long>(4 * kEmitThreads, (want + kEmitThreads - 1) / kEmitThreads * kEmitThreads));
        c->n_slabs = static_cast<int>((c->n_prims + c->slab_prims - 1) / c->slab_prims);
        c->slabs_per_seg = std::max(1, (c->n_slabs + kHistSegs - 1) / kHistSegs);
    }
#undef TB_TRY
    const int fw = cfg->flow_w > 0 ? cfg->flow_w : 1, fh = cfg->flow_h > 0 ? cfg->flow_h : 1;
    if (int r = alloc_flow(c, fw, fh)) return bail(r);
    if (int r = ensure_bin_cap(c, static_cast<uint64_t>(c->n_prims) * 6)) return bail(r);
    if (int r = tb_reset(c)) return bail(r);
    *out = c;
    return TB_OK;
}

int tb_destroy(tb_ctx *c) {
    if (!c) return TB_OK;
    cudaSetDevice(c->device);
    if (c->side) cudaStreamSynchronize(c->side);
    if (c->stream) cudaStreamSynchronize(c->stream);
    if (c->h2d) { cudaStreamSynchronize(c->h2d); cudaStreamDestroy(c->h2d); }
    if (c->d2h) { cudaStreamSynchronize(c->d2h); cudaStreamDestroy(c->d2h); }
    if (c->ev_main) cudaEventDestroy(c->ev_main);
    for (int i = 0; i < tb_ctx::kMaxChunks; ++i) {
        if (c->ev_h2d[i]) cudaEventDestroy(c->ev_h2d[i]);
        if (c->ev_int[i]) cudaEventDestroy(c->ev_int[i]);
        if (c->ev_d2h[i]) cudaEventDestroy(c->ev_d2h[i]);
    }
    tiles_release(c);
    cudaFree(c->ow_flags); cudaFree(c->ow_totals); cudaFree(c->ow_scratch); cudaFree(c->ow_last_all);
    cudaFree(c->buf[0]); cudaFree(c->buf[1]); cudaFree(c->spare); cudaFree(c->targets); cudaFree(c->flow);
    cudaFree(c->frames);
    cudaFree(c->line_attr); cudaFree(c->line_verts); cudaFree(c->line_bbox);
    cudaFree(c->image); cudaFree(c->layer); cudaFree(c->pairs); cudaFree(c->d_flag);
    cudaFree(c->slab_hist); cudaFree(c->seg_total); cudaFree(c->bin_total); cudaFree(c->bin_off); cudaFree(c->tickets); cudaFree(c->items);
    cudaFree(c->split_map); cudaFree(c->bin_info); cudaFree(c->n_bins);
    cudaFree(c->d_plan); cudaFree(c->bins); cudaFree(c->replay);
    cudaFree(c->seg_desc); cudaFree(c->seg_of_bin); cudaFree(c->seg_cnt); cudaFree(c->seg_out);
    if (c->h_flag) cudaFreeHost(c->h_flag);
    if (c->h_plan) cudaFreeHost(c->h_plan);
    if (c->ev_plan) cudaEventDestroy(c->ev_plan);
    for (int k = 0; k < 3; ++k)
        for (int i = 0; i < tb_ctx::kTimingSlots; ++i)
            for (int j = 0; j < 2; ++j)
                if (c->ev_ring[k][i][j]) cudaEventDestroy(c->ev_ring[k][i][j]);
    for (int i = 0; i < tb_ctx::kTimingSlots; ++i)
        for (int j = 0; j < 5; ++j)
            if (c->ev_stage[i][j]) cudaEventDestroy(c->ev_stage[i][j]);
    if (c->ev_state) cudaEventDestroy(c->ev_state);
    if (c->ev_producer) cudaEventDestroy(c->ev_producer);
    if (c->ev_noise) cudaEventDestroy(c->ev_noise);
    if (c->side) cudaStreamDestroy(c->side);
    cudaFree(c->wander);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
    return TB_OK;
}

int tb_set_state(tb_ctx *c, const tb_state *s) {
    TB_REQUIRE(c, c && s, "null argument");
    c->state = *s;
    c->have_state = true;
    return TB_OK;
}

int tb_set_overlap(tb_ctx *c, int32_t on) {
    TB_REQUIRE(c, c, "null context");
    c->overlap = on != 0;
    return TB_OK;
}

int tb_resize_flow(tb_ctx *c, int32_t w, int32_t h) {
    TB_REQUIRE(c, c, "null context");
    TB_CUDA(c, cudaSetDevice(c->device));
    if (int r = resolve_pending(c)) return r;
    TB_CUDA(c, cudaStreamSynchronize(c->stream));
    return alloc_flow(c, w, h);
}

int tb_clear_flow(tb_ctx *c) {
    TB_REQUIRE(c, c, "null context");
    TB_CUDA(c, cudaSetDevice(c->device));
    if (int r = resolve_pending(c)) return r;
    TB_CUDA(c, cudaMemsetAsync(c->flow, 0, static_cast<size_t>(c->W) * c->H * sizeof(float4), c->stream));
    return TB_OK;
}

namespace {
IntegrateArgs integrate_args(tb_ctx *c, float time, float dt) {
    const tb_state &S = c->state;
    IntegrateArgs A{};
    A.S = S;
    A.in = c->buf[1];
    A.out = c->buf[0];
    A.targets = c->targets;
    A.flow = c->flow;
    A.PW = c->PW; A.PH = c->PH; A.W = c->W; A.H = c->H;
    A.col0 = c->col0;
    A.cols = c->col1 - c->col0;
    A.time = time; A.dt = dt;
    // The target / noise terms are exactly +-0 for every finite particle when their weight is 0
    // and the variances are tame; the kernel still takes the full path for wild positions.
    A.use_targets = !(S.target == 0.0f && tame(S.varyTarget, 1e6f) && c->targets_finite);
    A.use_noise = !(S.noiseWeight == 0.0f && tame(S.varyNoise, 1e6f) && tame(S.noiseScale, 1e6f) &&
                    tame(S.varyNoiseScale, 1e6f) && tame(S.noiseSpeed, 1e6f) && tame(S.varyNoiseSpeed, 1e6f) &&
                    tame(time, 1e9f) && tame(dt, 1e6f));
    auto is_pow2 = [](int v) { return v > 0 && (v & (v - 1)) == 0; };
    A.pow2_res = is_pow2(c->PW) && is_pow2(c->PH) && static_cast<long long>(c->PW) * c->PH <= (1LL << 40);
    A.inv_resx = 1.0f / static_cast<float>(c->PW);
    A.inv_resy = 1.0f / static_cast<float>(c->PH);
    A.inv_n = 1.0f / (static_cast<float>(c->PW) * static_cast<float>(c->PH));
    static const bool scalar_noise = std::getenv("TB_SCALAR_NOISE") != nullptr;    // A/B switch for profiling
    A.packed_noise = scalar_noise ? 0 : 1;
    A.pk.one = 1.0f; A.pk.neg_one = -1.0f; A.pk.neg_zero = -0.0f;
    A.wander = c->wander;
    return A;
}
}  // namespace

int tb_step(tb_ctx *c, float time, float dt) {
    TB_REQUIRE(c, c, "null context");
    TB_REQUIRE(c, c->have_state, "tb_set_state must be called before tb_step");
    TB_CUDA(c, cudaSetDevice(c->device));
    if (int r = resolve_pending(c)) return r;
    if (int r = join_copies(c)) return r;
    std::swap(c->buf[0], c->buf[1]);                       // utils.step(buffers), src/particles.js:128
    const IntegrateArgs A = integrate_args(c, time, dt);
    cudaEvent_t *ev = c->ev_ring[0][c->ev_count[0] % tb_ctx::kTimingSlots];
    const dim3 grid(blocks_for(c->PH, 256), static_cast<unsigned>(A.cols));
    if (A.use_noise && c->overlap && c->splat_since_step) {
        // The noise does not read the flow grid: evaluate it on the low-priority side stream, where it
        // runs under the previous step's flow splat still queued on the main stream; the rest of the
        // shader follows on the main stream.  The side stream only waits for the state to be final.
        cudaEvent_t *evn = c->ev_ring[2][c->ev_count[2] % tb_ctx::kTimingSlots];
        TB_CUDA(c, cudaStreamWaitEvent(c->side, c->ev_state, 0));
        TB_CUDA(c, cudaEventRecord(evn[0], c->side));
        k_integrate<kNoise><<<grid, 256, 0, c->side>>>(A);
        if (int r = check_launch(c, "k_integrate<noise>")) return r;
        TB_CUDA(c, cudaEventRecord(evn[1], c->side));
        TB_CUDA(c, cudaEventRecord(c->ev_noise, c->side));
        c->ev_count[2] += 1;
        TB_CUDA(c, cudaStreamWaitEvent(c->stream, c->ev_noise, 0));
        TB_CUDA(c, cudaEventRecord(ev[0], c->stream));
        k_integrate<kFinish><<<grid, 256, 0, c->stream>>>(A);
        if (int r = check_launch(c, "k_integrate<finish>")) return r;
    } else {
        TB_CUDA(c, cudaEventRecord(ev[0], c->stream));
        k_integrate<kFused><<<grid, 256, 0, c->stream>>>(A);
        if (int r = check_launch(c, "k_integrate")) return r;
    }
    TB_CUDA(c, cudaEventRecord(ev[1], c->stream));
    TB_CUDA(c, cudaEventRecord(c->ev_state, c->stream));
    c->ev_count[0] += 1;
    c->splat_since_step = false;
    return TB_OK;
}

// tb_upload(CURRENT, host_in) + tb_step + tb_download(CURRENT, host_out) as one pipelined pass over column chunks: chunk i
// goes up on the h2d stream, through the logic pass on the main stream and down on the d2h stream while chunk i + 1 goes up.
// Returns as soon as everything is queued; host_out is complete after tb_sync.
int tb_step_streamed(tb_ctx *c, float time, float dt, const float *host_in, float *host_out, int32_t n_chunks) {
    TB_REQUIRE(c, c && host_in && host_out, "null argument");
    TB_REQUIRE(c, c->have_state, "tb_set_state must be called before tb_step_streamed");
    TB_REQUIRE(c, n_chunks >= 1 && n_chunks <= tb_ctx::kMaxChunks, "tb_step_streamed: 1..64 chunks");
    TB_CUDA(c, cudaSetDevice(c->device));
    if (int r = resolve_pending(c)) return r;
    const int cols = c->col1 - c->col0;
    const int C = std::min<int>(n_chunks, cols);
    if (!c->h2d) {
        TB_CUDA(c, cudaStreamCreateWithFlags(&c->h2d, cudaStreamNonBlocking));
        TB_CUDA(c, cudaStreamCreateWithFlags(&c->d2h, cudaStreamNonBlocking));
        TB_CUDA(c, cudaEventCreateWithFlags(&c->ev_main, cudaEventDisableTiming));
        for (int i = 0; i < tb_ctx::kMaxChunks; ++i) {
            TB_CUDA(c, cudaEventCreateWithFlags(&c->ev_h2d[i], cudaEventDisableTiming));
            TB_CUDA(c, cudaEventCreateWithFlags(&c->ev_int[i], cudaEventDisableTiming));
            TB_CUDA(c, cudaEventCreateWithFlags(&c->ev_d2h[i], cudaEventDisableTiming));
        }
    }
    if (!c->spare) TB_CUDA(c, cudaMalloc(&c->spare, static_cast<size_t>(c->n_local) * sizeof(float4)));
    // The upload goes into the spare buffer, which nothing queued reads (the previous flow splat reads the other two), so it
    // only has to follow the previous streamed step's download chunk by chunk: on the device that download is long done
    // with the spare buffer's former life, and when the caller round-trips the state through one host buffer it is the
    // host memory that must be complete.  The logic pass then writes over the PREVIOUS state, in stream order after the splat.
    const int prev_chunks = c->chunks_inflight;
    if (prev_chunks != C) {                                 // another chunking than last time: no chunk-wise overlap across the steps
        for (int i = 0; i < prev_chunks; ++i) TB_CUDA(c, cudaStreamWaitEvent(c->h2d, c->ev_d2h[i], 0));
    }
    float4 *in = c->spare, *out = c->buf[1];
    c->spare = c->buf[0];                                   // utils.step(buffers) over three buffers
    c->buf[0] = out;
    c->buf[1] = in;
    IntegrateArgs A = integrate_args(c, time, dt);
    cudaEvent_t *ev = c->ev_ring[0][c->ev_count[0] % tb_ctx::kTimingSlots];
    TB_CUDA(c, cudaEventRecord(ev[0], c->stream));
    for (int i = 0; i < C; ++i) {
        const int x0 = static_cast<int>(static_cast<long long>(cols) * i / C), x1 = static_cast<int>(static_cast<long long>(cols) * (i + 1) / C);
        const size_t off = static_cast<size_t>(x0) * c->PH, n = static_cast<size_t>(x1 - x0) * c->PH;
        if (prev_chunks == C) TB_CUDA(c, cudaStreamWaitEvent(c->h2d, c->ev_d2h[i], 0));
        TB_CUDA(c, cudaMemcpyAsync(c->buf[1] + off, host_in + 4 * off, n * sizeof(float4), cudaMemcpyHostToDevice, c->h2d));
        TB_CUDA(c, cudaEventRecord(c->ev_h2d[i], c->h2d));
        TB_CUDA(c, cudaStreamWaitEvent(c->stream, c->ev_h2d[i], 0));
        IntegrateArgs Ai = A;
        Ai.in = A.in + off; Ai.out = A.out + off; Ai.targets = A.targets + off;
        Ai.col0 = c->col0 + x0; Ai.cols = x1 - x0;
        k_integrate<kFused><<<dim3(blocks_for(c->PH, 256), static_cast<unsigned>(x1 - x0)), 256, 0, c->stream>>>(Ai);
        if (int r = check_launch(c, "k_integrate")) return r;
        TB_CUDA(c, cudaEventRecord(c->ev_int[i], c->stream));
        TB_CUDA(c, cudaStreamWaitEvent(c->d2h, c->ev_int[i], 0));
        TB_CUDA(c, cudaMemcpyAsync(host_out + 4 * off, c->buf[0] + off, n * sizeof(float4), cudaMemcpyDeviceToHost, c->d2h));
        TB_CUDA(c, cudaEventRecord(c->ev_d2h[i], c->d2h));
    }
    TB_CUDA(c, cudaEventRecord(ev[1], c->stream));
    TB_CUDA(c, cudaEventRecord(c->ev_state, c->stream));
    c->ev_count[0] += 1;
    c->splat_since_step = false;
    c->chunks_inflight = C;
    return TB_OK;
}

namespace {
struct OwnerHandles {                 // what tb_owners_export hands to every other rank
    cudaIpcMemHandle_t bins, flow, totals, flags, last;
    int32_t w, h;
    uint32_t bin_cap, pad;
};
}  // namespace

int64_t tb_owners_handle_bytes(void) { return static_cast<int64_t>(sizeof(OwnerHandles)); }

int tb_owners_export(tb_ctx *c, int64_t reserve_fragments, void *out, int64_t n_bytes) {
    TB_REQUIRE(c, c && out, "null argument");
    TB_REQUIRE(c, n_bytes == static_cast<int64_t>(sizeof(OwnerHandles)), "tb_owners_export: buffer must be tb_owners_handle_bytes() long");
    TB_REQUIRE(c, reserve_fragments >= 0 && reserve_fragments < (1ll << 31), "tb_owners_export: reserve out of range");
    TB_CUDA(c, cudaSetDevice(c->device));
    if (int r = resolve_pending(c)) return r;
    TB_CUDA(c, cudaStreamSynchronize(c->stream));
    tiles_release(c);
    if (int e = ensure_bin_cap(c, std::max<uint64_t>(static_cast<uint64_t>(reserve_fragments), 1u << 16))) return e;
    if (!c->ow_flags) {
        TB_CUDA(c, cudaMalloc(&c->ow_flags, (kOwnerPhases + 1) * kMaxBandRanks * sizeof(uint32_t)));
        TB_CUDA(c, cudaMalloc(&c->ow_totals, static_cast<size_t>(kMaxBandRanks) * kMaxBins * sizeof(uint32_t)));
        TB_CUDA(c, cudaMalloc(&c->ow_scratch, 4 * static_cast<size_t>(kMaxBins) * sizeof(uint32_t)));
    }
    TB_CUDA(c, cudaMemset(c->ow_flags, 0, (kOwnerPhases + 1) * kMaxBandRanks * sizeof(uint32_t)));
    cudaFree(c->ow_last_all);
    c->ow_last_all = nullptr;
    TB_CUDA(c, cudaMalloc(&c->ow_last_all, static_cast<size_t>(kMaxBandRanks) * c->W * c->H * sizeof(uint32_t)));
    TB_CUDA(c, cudaMemset(c->ow_last_all, 0, static_cast<size_t>(kMaxBandRanks) * c->W * c->H * sizeof(uint32_t)));
    TB_CUDA(c, cudaMemset(c->ow_totals, 0, static_cast<size_t>(kMaxBandRanks) * kMaxBins * sizeof(uint32_t)));
    c->ow_epoch = 0;
    OwnerHandles hnd{};
    TB_CUDA(c, cudaIpcGetMemHandle(&hnd.bins, c->bins));
    TB_CUDA(c, cudaIpcGetMemHandle(&hnd.flow, c->flow));
    TB_CUDA(c, cudaIpcGetMemHandle(&hnd.totals, c->ow_totals));
    TB_CUDA(c, cudaIpcGetMemHandle(&hnd.flags, c->ow_flags));
    TB_CUDA(c, cudaIpcGetMemHandle(&hnd.last, c->ow_last_all));
    hnd.w = c->W; hnd.h = c->H; hnd.bin_cap = c->bin_cap;
    std::memcpy(out, &hnd, sizeof(hnd));
    c->bin_fixed = true;
    return TB_OK;
}

int tb_owners_connect(tb_ctx *c, int32_t rank, int32_t world, const void *all_handles, int64_t n_bytes) {
    TB_REQUIRE(c, c && all_handles, "null argument");
    TB_REQUIRE(c, world >= 2 && world <= kMaxBandRanks && rank >= 0 && rank < world, "tb_owners_connect: bad rank/world");
    TB_REQUIRE(c, n_bytes == static_cast<int64_t>(sizeof(OwnerHandles)) * world, "tb_owners_connect: expected world x tb_owners_handle_bytes()");
    TB_REQUIRE(c, c->bin_fixed && c->ow_flags, "tb_owners_export must be called before tb_owners_connect");
    TB_CUDA(c, cudaSetDevice(c->device));
    const auto *h = static_cast<const unsigned char *>(all_handles);
    c->ow_rank = rank; c->ow_world = world;
    c->owners_connected = true;                // from here on tiles_release skips this rank's own slots
    for (int j = 0; j < world; ++j) {
        OwnerHandles hnd;
        std::memcpy(&hnd, h + sizeof(OwnerHandles) * j, sizeof(hnd));
        c->ow_caps[j] = hnd.bin_cap;
        if (j == rank) {
            c->ow_bins[j] = c->bins; c->ow_flow[j] = c->flow;
            c->ow_peers.totals[j] = c->ow_totals; c->ow_peers.flags[j] = c->ow_flags; c->ow_peers.last[j] = c->ow_last_all;
            continue;
        }
        if (hnd.w != c->W || hnd.h != c->H) {
            tiles_release(c);
            c->bin_fixed = true;
            return fail(c, TB_ERR_INVALID, "tendrils-b200: tb_owners_connect: rank " + std::to_string(j) + " has another flow grid shape");
        }
        void *p[5] = {};
        const cudaIpcMemHandle_t *hs[5] = {&hnd.bins, &hnd.flow, &hnd.totals, &hnd.flags, &hnd.last};
        for (int k = 0; k < 5; ++k) {
            cudaError_t e = cudaIpcOpenMemHandle(&p[k], *hs[k], cudaIpcMemLazyEnablePeerAccess);
            // keep what was opened so far where tiles_release will find it
            if (k == 0) c->ow_bins[j] = static_cast<Frag *>(p[0]);
            if (k == 1) c->ow_flow[j] = static_cast<float4 *>(p[1]);
            if (k == 2) c->ow_peers.totals[j] = static_cast<uint32_t *>(p[2]);
            if (k == 3) c->ow_peers.flags[j] = static_cast<uint32_t *>(p[3]);
            if (k == 4) c->ow_peers.last[j] = static_cast<uint32_t *>(p[4]);
            if (e != cudaSuccess) {
                tiles_release(c);
                c->bin_fixed = true;
                return fail(c, TB_ERR_CUDA, std::string("cudaIpcOpenMemHandle (rank ") + std::to_string(j) + "): " + cudaGetErrorString(e));
            }
        }
    }
    c->ow_peers.n = world; c->ow_peers.me = rank;
    return TB_OK;
}

int tb_splat_flow_owners(tb_ctx *c, float time) {
    TB_REQUIRE(c, c, "null context");
    TB_REQUIRE(c, c->have_state, "tb_set_state must be called before the flow splat");
    TB_REQUIRE(c, c->owners_connected, "tb_owners_connect must be called first");
    TB_CUDA(c, cudaSetDevice(c->device));
    if (int r = resolve_pending(c)) return r;
    c->collect_time = time;
    c->collected = false;
    c->splat_since_step = true;
    c->pending_time = time;
    c->pending_stage = 3;
    if (int r = queue_owners(c, time)) return r;
    c->pending = true;
    c->ev_count[1] += 1;
    return TB_OK;
}

int tb_splat_collect(tb_ctx *c, float time) {
    TB_REQUIRE(c, c, "null context");
    TB_CUDA(c, cudaSetDevice(c->device));
    return collect(c, time);
}

int tb_splat_fold(tb_ctx *c) {
    TB_REQUIRE(c, c, "null context");
    TB_CUDA(c, cudaSetDevice(c->device));
    return fold(c);
}

int tb_splat_flow(tb_ctx *c, float time) {
    TB_REQUIRE(c, c, "null context");
    TB_CUDA(c, cudaSetDevice(c->device));
    return splat(c, time);
}

int tb_reset(tb_ctx *c) {
    TB_REQUIRE(c, c, "null context");
    TB_CUDA(c, cudaSetDevice(c->device));
    if (int r = resolve_pending(c)) return r;
    if (int r = join_copies(c)) return r;
    for (int b = 0; b < 2; ++b) {
        k_spawn_init<<<blocks_for(c->n_local, 256), 256, 0, c->stream>>>(c->buf[b], c->n_local);
        if (int r = check_launch(c, "k_spawn_init")) return r;
    }
    TB_CUDA(c, cudaEventRecord(c->ev_state, c->stream));
    return TB_OK;
}

int tb_spawn_init(tb_ctx *c, tb_target target) {
    TB_REQUIRE(c, c, "null context");
    TB_CUDA(c, cudaSetDevice(c->device));
    if (int r = resolve_pending(c)) return r;
    if (int r = join_copies(c)) return r;
    float4 *out = spawn_out(c, target);
    k_spawn_init<<<blocks_for(c->n_local, 256), 256, 0, c->stream>>>(out, c->n_local);
    if (int r = check_launch(c, "k_spawn_init")) return r;
    return after_targets_write(c, target);
}

int tb_spawn_ball(tb_ctx *c, float radius, float speed, tb_target target) {
    TB_REQUIRE(c, c, "null context");
    TB_CUDA(c, cudaSetDevice(c->device));
    if (int r = resolve_pending(c)) return r;
    if (int r = join_copies(c)) return r;
    SpawnArgs A{};
    A.out = spawn_out(c, target);
    A.PW = c->PW; A.PH = c->PH;
    A.p0 = static_cast<long long>(c->col0) * c->PH;
    A.n = c->n_local;
    A.radius = radius; A.speed = speed;
    k_spawn_ball<<<blocks_for(A.n, 256), 256, 0, c->stream>>>(A);
    if (int r = check_launch(c, "k_spawn_ball")) return r;
    return after_targets_write(c, target);
}

int tb_set_spawn_image(tb_ctx *c, const float *rgba, int32_t w, int32_t h) {
    TB_REQUIRE(c, c && rgba, "null argument");
    TB_REQUIRE(c, w >= 1 && h >= 1, "gl-texture2d: Texture dimensions are out of bounds");
    TB_CUDA(c, cudaSetDevice(c->device));
    if (int r = resolve_pending(c)) return r;
    const size_t need = static_cast<size_t>(w) * h;
    if (need > c->image_cap) {
        TB_CUDA(c, cudaStreamSynchronize(c->stream));
        if (c->image) cudaFree(c->image);
        c->image = nullptr; c->image_cap = 0;
        TB_CUDA(c, cudaMalloc(&c->image, need * sizeof(float4)));
        c->image_cap = need;
    }
    TB_CUDA(c, cudaMemcpyAsync(c->image, rgba, need * sizeof(float4), cudaMemcpyDefault, c->stream));
    if (!is_device_pointer(rgba)) TB_CUDA(c, cudaStreamSynchronize(c->stream));       // a host pointer is only borrowed
    c->IW = w; c->IH = h;
    return TB_OK;
}

int tb_spawn_pixels(tb_ctx *c, const tb_pixel_spawner *params, tb_spawn_variant variant, tb_spawn_source source,
                    float time, tb_target target) {
    TB_REQUIRE(c, c && params, "null argument");
    TB_REQUIRE(c, c->have_state, "tb_set_state must be called before tb_spawn_pixels");
    TB_CUDA(c, cudaSetDevice(c->device));
    if (int r = resolve_pending(c)) return r;
    if (int r = join_copies(c)) return r;
    SpawnArgs A{};
    A.U = *params;
    A.PW = c->PW; A.PH = c->PH;
    A.p0 = static_cast<long long>(c->col0) * c->PH;
    A.n = c->n_local;
    A.time = time;
    A.flowDecay = c->state.flowDecay;
    switch (variant) {
        case TB_SPAWN_DIRECT:        A.apply = APPLY_COLOR;     A.vignette = 1; A.samples = 0; break;
        case TB_SPAWN_BEST_SAMPLE:   A.apply = APPLY_COLOR;     A.vignette = 1; A.samples = 6; break;
        case TB_SPAWN_BRIGHT_SAMPLE: A.apply = APPLY_BRIGHTEST; A.vignette = 0; A.samples = 6; break;
        case TB_SPAWN_COLOR_SAMPLE:  A.apply = APPLY_COLOR;     A.vignette = 0; A.samples = 3; break;
        case TB_SPAWN_DATA_SAMPLE:   A.apply = APPLY_IDENTITY;  A.vignette = 1; A.samples = 2; break;
        case TB_SPAWN_FLOW_SAMPLE:   A.apply = APPLY_FLOW;      A.vignette = 0; A.samples = 5; break;
        default: return fail(c, TB_ERR_UNSUPPORTED, "tendrils-b200: custom spawn shaders are not supported");
    }
    // `particles` and spawnData are bound BEFORE the ping-pong rotates (src/particles.js:124-141):
    // particles = buffers[1] after rotation = the current state before it.
    const float4 *current_before = c->buf[0];
    switch (source) {
        case TB_SOURCE_IMAGE:
            TB_REQUIRE(c, c->image && c->IW > 0, "tb_set_spawn_image must be called first");
            A.image = c->image; A.IW = c->IW; A.IH = c->IH; A.image_xmajor = 0;
            break;
        case TB_SOURCE_FLOW:
            A.image = c->flow; A.IW = c->W; A.IH = c->H; A.image_xmajor = 0;
            break;
        case TB_SOURCE_PARTICLES:
            TB_REQUIRE(c, c->col0 == 0 && c->col1 == c->PW, "spawning from the particle texture needs an unsharded context");
            A.image = current_before; A.IW = c->PW; A.IH = c->PH; A.image_xmajor = 1;
            break;
        default: return fail(c, TB_ERR_INVALID, "tendrils-b200: bad spawn source");
    }
    A.out = spawn_out(c, target);
    A.state = c->buf[1];
    if (variant == TB_SPAWN_DIRECT) {
        k_spawn_direct<<<blocks_for(A.n, 256), 256, 0, c->stream>>>(A);
        if (int r = check_launch(c, "k_spawn_direct")) return r;
    } else {
        k_spawn_sample<<<blocks_for(A.n, 256), 256, 0, c->stream>>>(A);
        if (int r = check_launch(c, "k_spawn_sample")) return r;
    }
    return after_targets_write(c, target);
}

static int buffer_of(tb_ctx *c, tb_buffer which, float4 **ptr, int64_t *n_floats) {
    switch (which) {
        case TB_BUF_CURRENT:  *ptr = c->buf[0];  *n_floats = 4 * c->n_local; return TB_OK;
        case TB_BUF_PREVIOUS: *ptr = c->buf[1];  *n_floats = 4 * c->n_local; return TB_OK;
        case TB_BUF_TARGETS:  *ptr = c->targets; *n_floats = 4 * c->n_local; return TB_OK;
        case TB_BUF_FLOW:     *ptr = c->flow;    *n_floats = 4LL * c->W * c->H; return TB_OK;
    }
    return fail(c, TB_ERR_INVALID, "tendrils-b200: bad buffer id");
}

int tb_upload(tb_ctx *c, tb_buffer which, const float *host, int64_t n_floats) {
    TB_REQUIRE(c, c && host, "null argument");
    TB_CUDA(c, cudaSetDevice(c->device));
    if (int r = resolve_pending(c)) return r;
    if (int r = join_copies(c)) return r;
    float4 *dst; int64_t n;
    if (int r = buffer_of(c, which, &dst, &n)) return r;
    TB_REQUIRE(c, n == n_floats, "tb_upload: size mismatch");
    TB_CUDA(c, cudaStreamSynchronize(c->side));
    TB_CUDA(c, cudaMemcpyAsync(dst, host, static_cast<size_t>(n) * sizeof(float), cudaMemcpyHostToDevice, c->stream));
    TB_CUDA(c, cudaStreamSynchronize(c->stream));
    if (which == TB_BUF_TARGETS) return after_targets_write(c, TB_TARGET_TARGETS);
    TB_CUDA(c, cudaEventRecord(c->ev_state, c->stream));
    return TB_OK;
}

int tb_download(tb_ctx *c, tb_buffer which, float *host, int64_t n_floats) {
    TB_REQUIRE(c, c && host, "null argument");
    TB_CUDA(c, cudaSetDevice(c->device));
    if (int r = resolve_pending(c)) return r;
    if (int r = join_copies(c)) return r;
    float4 *src; int64_t n;
    if (int r = buffer_of(c, which, &src, &n)) return r;
    TB_REQUIRE(c, n == n_floats, "tb_download: size mismatch");
    TB_CUDA(c, cudaMemcpyAsync(host, src, static_cast<size_t>(n) * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    TB_CUDA(c, cudaStreamSynchronize(c->stream));
    return TB_OK;
}

int tb_blend_into_flow(tb_ctx *c, const float *rgba, int32_t w, int32_t h) {
    TB_REQUIRE(c, c && rgba, "null argument");
    TB_REQUIRE(c, w == c->W && h == c->H, "tb_blend_into_flow: layer must have the flow grid's shape");
    TB_CUDA(c, cudaSetDevice(c->device));
    if (int r = resolve_pending(c)) return r;
    const size_t G = static_cast<size_t>(w) * h;
    if (G > c->layer_cap) {
        TB_CUDA(c, cudaStreamSynchronize(c->stream));
        if (c->layer) cudaFree(c->layer);
        c->layer = nullptr; c->layer_cap = 0;
        TB_CUDA(c, cudaMalloc(&c->layer, G * sizeof(float4)));
        c->layer_cap = G;
    }
    TB_CUDA(c, cudaMemcpyAsync(c->layer, rgba, G * sizeof(float4), cudaMemcpyHostToDevice, c->stream));
    k_blend_layer<<<blocks_for(static_cast<long long>(G), 256), 256, 0, c->stream>>>(c->flow, c->layer, static_cast<int>(G));
    if (int r = check_launch(c, "k_blend_layer")) return r;
    TB_CUDA(c, cudaStreamSynchronize(c->stream));
    return TB_OK;
}

int tb_flow_line(tb_ctx *c, const tb_flow_line_params *params, int32_t n_vertices, const float *position, const float *normal,
                 const float *miter, const float *previous, const float *time, const float *dt) {
    TB_REQUIRE(c, c && params, "null argument");
    TB_REQUIRE(c, n_vertices >= 0 && n_vertices <= (1 << 24), "tb_flow_line: vertex count out of range");
    if (n_vertices < 3) return TB_OK;                       // a strip needs three vertices to make a triangle
    TB_REQUIRE(c, position && normal && miter && previous && time && dt, "null attribute array");
    TB_CUDA(c, cudaSetDevice(c->device));
    if (int r = resolve_pending(c)) return r;
    const int n = n_vertices;
    if (n > c->line_cap) {
        TB_CUDA(c, cudaStreamSynchronize(c->stream));
        cudaFree(c->line_attr); cudaFree(c->line_verts);
        c->line_attr = nullptr; c->line_verts = nullptr; c->line_cap = 0;
        const int cap = std::max(n, 256);
        TB_CUDA(c, cudaMalloc(&c->line_attr, static_cast<size_t>(cap) * 9 * sizeof(float)));
        TB_CUDA(c, cudaMalloc(&c->line_verts, static_cast<size_t>(cap) * sizeof(fl::Vertex)));
        c->line_cap = cap;
    }
    if (!c->line_bbox) TB_CUDA(c, cudaMalloc(&c->line_bbox, 4 * sizeof(int)));
    // attribute arrays as gl-geometry holds them (src/geom/line/index.js:119-123): one buffer per attribute
    float *d_pos = c->line_attr, *d_nor = d_pos + 2 * n, *d_mit = d_nor + 2 * n, *d_prv = d_mit + n, *d_tim = d_prv + 2 * n,
          *d_dt = d_tim + n;
    TB_CUDA(c, cudaMemcpyAsync(d_pos, position, 2 * n * sizeof(float), cudaMemcpyDefault, c->stream));
    TB_CUDA(c, cudaMemcpyAsync(d_nor, normal, 2 * n * sizeof(float), cudaMemcpyDefault, c->stream));
    TB_CUDA(c, cudaMemcpyAsync(d_mit, miter, n * sizeof(float), cudaMemcpyDefault, c->stream));
    TB_CUDA(c, cudaMemcpyAsync(d_prv, previous, 2 * n * sizeof(float), cudaMemcpyDefault, c->stream));
    TB_CUDA(c, cudaMemcpyAsync(d_tim, time, n * sizeof(float), cudaMemcpyDefault, c->stream));
    TB_CUDA(c, cudaMemcpyAsync(d_dt, dt, n * sizeof(float), cudaMemcpyDefault, c->stream));
    const int empty_box[4] = {INT_MAX, INT_MAX, INT_MIN, INT_MIN};
    TB_CUDA(c, cudaMemcpyAsync(c->line_bbox, empty_box, sizeof(empty_box), cudaMemcpyHostToDevice, c->stream));
    fl::Uniforms U{params->viewSize[0], params->viewSize[1], params->rad, params->speed, params->speedLimit, params->crestShape};
    k_flow_line_vertices<<<blocks_for(n, 128), 128, 0, c->stream>>>(U, n, d_pos, d_nor, d_mit, d_prv, d_tim, d_dt, c->W, c->H,
                                                                     c->line_verts, c->line_bbox);
    if (int r = check_launch(c, "k_flow_line_vertices")) return r;
    const dim3 grid(blocks_for(c->W, 256), static_cast<unsigned>(c->H));
    k_flow_line_raster<<<grid, 256, 0, c->stream>>>(c->line_verts, n, U.crestShape, c->line_bbox, c->flow, c->W, c->H);
    if (int r = check_launch(c, "k_flow_line_raster")) return r;
    TB_CUDA(c, cudaStreamSynchronize(c->stream));           // the host arrays are only borrowed
    return TB_OK;
}

int tb_optical_flow(tb_ctx *c, const tb_optical_flow_params *params, const uint8_t *view_rgba8, const uint8_t *last_rgba8,
                    int32_t w, int32_t h) {
    TB_REQUIRE(c, c && params && view_rgba8 && last_rgba8, "null argument");
    TB_REQUIRE(c, w >= 1 && h >= 1, "gl-texture2d: Texture dimensions are out of bounds");
    TB_CUDA(c, cudaSetDevice(c->device));
    if (int r = resolve_pending(c)) return r;
    const size_t n = static_cast<size_t>(w) * h;
    if (2 * n > c->frames_cap) {
        TB_CUDA(c, cudaStreamSynchronize(c->stream));
        if (c->frames) cudaFree(c->frames);
        c->frames = nullptr; c->frames_cap = 0;
        TB_CUDA(c, cudaMalloc(&c->frames, 2 * n * sizeof(uchar4)));
        c->frames_cap = 2 * n;
    }
    const bool resident = is_device_pointer(view_rgba8) && is_device_pointer(last_rgba8);
    TB_CUDA(c, cudaMemcpyAsync(c->frames, view_rgba8, n * sizeof(uchar4), cudaMemcpyDefault, c->stream));
    TB_CUDA(c, cudaMemcpyAsync(c->frames + n, last_rgba8, n * sizeof(uchar4), cudaMemcpyDefault, c->stream));
    OpticalArgs A{};
    A.flow = c->flow; A.view = c->frames; A.last = c->frames + n;
    A.W = c->W; A.H = c->H; A.IW = w; A.IH = h;
    A.U = *params;
    k_optical_flow<<<blocks_for(static_cast<long long>(c->W) * c->H, 256), 256, 0, c->stream>>>(A);
    if (int r = check_launch(c, "k_optical_flow")) return r;
    if (!resident) TB_CUDA(c, cudaStreamSynchronize(c->stream));       // host frames are only borrowed
    return TB_OK;
}

int tb_debug_max_bins(void) { return kMaxBins; }

int tb_debug_bins(tb_ctx *c, uint32_t *offsets, uint32_t *info, int32_t *n_bins, int32_t *strip_w, int32_t *strip_h) {
    TB_REQUIRE(c, c && offsets && info && n_bins, "null argument");
    TB_CUDA(c, cudaSetDevice(c->device));
    if (int r = resolve_pending(c)) return r;
    uint32_t nb = 0;
    TB_CUDA(c, cudaMemcpyAsync(&nb, c->n_bins + c->fold_parity, sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
    TB_CUDA(c, cudaMemcpyAsync(offsets, c->bin_off, static_cast<size_t>(kMaxBins + 1) * sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
    TB_CUDA(c, cudaMemcpyAsync(info, c->bin_info + static_cast<size_t>(c->fold_parity) * kMaxBins, static_cast<size_t>(kMaxBins) * sizeof(uint32_t),
                               cudaMemcpyDeviceToHost, c->stream));
    TB_CUDA(c, cudaStreamSynchronize(c->stream));
    *n_bins = static_cast<int32_t>(nb);
    if (strip_w) *strip_w = 1 << c->geom.sxl;
    if (strip_h) *strip_h = 1 << c->geom.syl;
    return TB_OK;
}

int tb_debug_segments(tb_ctx *c, int32_t *bins, int32_t *segments, int64_t *records) {
    TB_REQUIRE(c, c && bins && segments && records, "null argument");
    TB_CUDA(c, cudaSetDevice(c->device));
    if (int r = resolve_pending(c)) return r;
    *bins = 0; *segments = 0; *records = 0;
    if (!c->seg_at || !c->replay) return TB_OK;
    uint32_t nb = 0;
    TB_CUDA(c, cudaMemcpyAsync(&nb, c->tickets + 5, sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
    TB_CUDA(c, cudaStreamSynchronize(c->stream));
    if (nb == 0 || nb > static_cast<uint32_t>(kMaxBins)) return TB_OK;
    std::vector<uint4> desc(nb);
    std::vector<uint32_t> cnt(16 * static_cast<size_t>(kMaxBins));
    TB_CUDA(c, cudaMemcpyAsync(desc.data(), c->seg_desc, nb * sizeof(uint4), cudaMemcpyDeviceToHost, c->stream));
    TB_CUDA(c, cudaMemcpyAsync(cnt.data(), c->seg_cnt, cnt.size() * sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
    TB_CUDA(c, cudaStreamSynchronize(c->stream));
    *bins = static_cast<int32_t>(nb);
    for (const uint4 &d : desc) {
        *segments += 1 << d.y;
        for (uint32_t part = 1; part < (1u << d.y); ++part) *records += cnt[d.w + part];
    }
    return TB_OK;
}

int tb_device_ptr(tb_ctx *c, tb_buffer which, void **ptr, int64_t *n_floats) {
    TB_REQUIRE(c, c && ptr && n_floats, "null argument");
    float4 *p; int64_t n;
    if (int r = buffer_of(c, which, &p, &n)) return r;
    *ptr = p; *n_floats = n;
    return TB_OK;
}

// Order this context's stream after everything queued so far on `producer_stream` (a cudaStream_t of this device): for inputs
// that live in device memory and were written by someone else's stream (a decoder, torch's current stream).
int tb_wait_stream(tb_ctx *c, void *producer_stream) {
    TB_REQUIRE(c, c, "null context");
    TB_CUDA(c, cudaSetDevice(c->device));
    if (!c->ev_producer) TB_CUDA(c, cudaEventCreateWithFlags(&c->ev_producer, cudaEventDisableTiming));
    TB_CUDA(c, cudaEventRecord(c->ev_producer, static_cast<cudaStream_t>(producer_stream)));
    TB_CUDA(c, cudaStreamWaitEvent(c->stream, c->ev_producer, 0));
    return TB_OK;
}

int tb_stream(tb_ctx *c, void **cuda_stream) {
    TB_REQUIRE(c, c && cuda_stream, "null argument");
    *cuda_stream = c->stream;
    return TB_OK;
}

int tb_sync(tb_ctx *c) {
    TB_REQUIRE(c, c, "null context");
    TB_CUDA(c, cudaSetDevice(c->device));
    if (int r = resolve_pending(c)) return r;
    TB_CUDA(c, cudaStreamSynchronize(c->side));
    TB_CUDA(c, cudaStreamSynchronize(c->stream));
    if (c->h2d) TB_CUDA(c, cudaStreamSynchronize(c->h2d));
    if (c->d2h) TB_CUDA(c, cudaStreamSynchronize(c->d2h));
    return TB_OK;
}

int tb_stats(tb_ctx *c, int64_t *kernel_launches, int64_t *last_fragments) {
    TB_REQUIRE(c, c, "null context");
    TB_CUDA(c, cudaSetDevice(c->device));
    if (int r = resolve_pending(c)) return r;
    if (kernel_launches) *kernel_launches = c->launches;
    if (last_fragments) *last_fragments = c->last_frags;
    return TB_OK;
}

int tb_timing(tb_ctx *c, int reset, int64_t *n_integrate, float *integrate_ms, int64_t *n_splat, float *splat_ms,
              int64_t *n_noise, float *noise_ms) {
    TB_REQUIRE(c, c, "null context");
    TB_CUDA(c, cudaSetDevice(c->device));
    if (int r = resolve_pending(c)) return r;
    TB_CUDA(c, cudaStreamSynchronize(c->side));
    TB_CUDA(c, cudaStreamSynchronize(c->stream));
    int64_t *n_out[3] = {n_integrate, n_splat, n_noise};
    float *ms_out[3] = {integrate_ms, splat_ms, noise_ms};
    for (int k = 0; k < 3; ++k) {
        const int64_t n = std::min<int64_t>(c->ev_count[k], tb_ctx::kTimingSlots);
        double total = 0.0;
        for (int64_t i = 0; i < n; ++i) {
            float ms = 0.f;
            TB_CUDA(c, cudaEventElapsedTime(&ms, c->ev_ring[k][i][0], c->ev_ring[k][i][1]));
            total += ms;
        }
        if (n_out[k]) *n_out[k] = n;
        if (ms_out[k]) *ms_out[k] = static_cast<float>(total);
        if (reset) c->ev_count[k] = 0;
    }
    return TB_OK;
}

}  // extern "C"
