// tb_api.cu -- the C ABI of include/tendrils_b200.h over the sm_100a kernels.
//
// Mirrors, for the one hot path, the GPU-facing behaviour of the reference's
// Tendrils (src/index.js) and Particles (src/particles.js) classes: ping-pong state
// buffers, the logic pass, the flow splat, the spawn passes.  No CPU fallback.
#include "tb_kernels.cuh"
#include "tb_flowline.cuh"

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

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
    cudaStream_t side = nullptr;            // low priority: noise of the next step under the previous splat
    cudaEvent_t ev_state = nullptr;         // last write to the state buffers (main stream)
    cudaEvent_t ev_noise = nullptr;         // noise kernel done (side stream)
    float2 *wander = nullptr;
    bool splat_since_step = false;          // something HBM-bound is queued that the noise can hide under
    bool overlap = true;

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

    // flow splat scratch
    PairEntry *pairs = nullptr;
    int n_pairs = 0;
    long long n_prims = 0;                 // local primitives = columns * n_pairs
    uint32_t *prim_off = nullptr;          // n_prims+1: fragments per primitive -> exclusive offsets, [n_prims] = total
    int32_t *row_pair = nullptr;           // per texture row: pair index | kind << 30, or -1 (fused count in k_integrate)
    bool fuse_count = false;               // every pair reads one particle's prev/cur: the count can ride in k_integrate
    bool fuse_partial = false;             // TB_FUSE_PARTIAL: all but a few pairs ride there, k_splat_count_odd counts the rest
    int n_odd = 0;
    int32_t *odd_pairs = nullptr;          // indices of the pairs that do not ride
    bool count_valid = false;              // prim_off/total on the host belong to the current (state, flow shape, viewSize)
    float count_vs[2] = {0.f, 0.f};
    int count_wh[2] = {0, 0};
    void *scan_tmp = nullptr;
    size_t scan_tmp_bytes = 0;
    uint32_t *keys[2] = {nullptr, nullptr};   // texel of each fragment (draw order / sorted)
    FragVal *vals[2] = {nullptr, nullptr};
    void *sort_tmp = nullptr;
    size_t sort_tmp_bytes = 0;
    uint32_t frag_cap = 0;
    uint32_t *seg = nullptr;               // 2*G: [begin,end) of every texel's sorted segment
    uint32_t *hot = nullptr;               // [0] = count, [1] = cursor, [2..G+1] = worklist of hot texels
    int n_sms = 148;
    uint32_t hot_threshold = kFoldHot;

    // sharded ring fold over peer memory (tb_ring_*): my inbox + flags, the next rank's mapped over NVLink
    static constexpr int kRingMaxChunks = 64;
    int ring_chunks = 16;
    int ring_rank = 0, ring_world = 1;
    float4 *inbox = nullptr;               // the previous rank writes its folded chunks here
    uint32_t *ring_flags = nullptr;        // [0..C): inbox chunk ready (epoch), [C..2C): final chunk in my grid (epoch)
    uint32_t *ring_hot_counts = nullptr;   // per chunk hot-texel counters
    float4 *next_inbox = nullptr, *next_flow = nullptr;
    uint32_t *next_flags = nullptr;
    bool ring_connected = false;
    uint32_t ring_epoch = 0;
    cudaStream_t ring_stream = nullptr;    // forwards final chunks around the ring
    cudaEvent_t ev_ring_fwd = nullptr, ev_ring_begin = nullptr;
    // band fold over peer memory (tb_bands_*): every rank maps every rank's sorted fragments, segments, grid, flags
    bool bands_connected = false;
    bool frag_fixed = false;               // keys/vals are exported through IPC handles: they must not be reallocated
    int bands_rank = 0, bands_world = 1;
    uint32_t bands_epoch = 0;
    uint32_t *bands_flags = nullptr;       // [kBandPhases][kMaxBandRanks]: the epoch each rank has reached
    BandSources band_src{};                // every rank's segment table (own slot: local pointer)
    BandSinks band_sinks{};                // every rank's merged fragment array (its vals[0]) and offset table
    BandPeers band_peers{};                // every rank's flow grid and flags
    int bands_mine = 0;                    // 32-texel tiles this rank folds
    uint32_t *bands_len = nullptr;         // [mine*32*world + 1] fragments per (local texel, source), texel-major; last = 0
    uint32_t *bands_off = nullptr;         // its exclusive scan: where each segment goes in the merged array
    uint32_t *bands_dst = nullptr;         // [G] written by the owners: where MY segment of each texel goes in its owner's array
    uint32_t *bands_seg = nullptr;         // 2*G: merged segment table (only this rank's texels are meaningful)
    void *bands_scan_tmp = nullptr;
    size_t bands_scan_bytes = 0;
    int *bands_overflow = nullptr;         // device flag: the merged fragments did not fit
    int *h_bands_overflow = nullptr;       // pinned copy, read at the start of the next fold
    cudaEvent_t ev_bands = nullptr;
    int key_bits = 1;
    uint32_t *h_total = nullptr;           // pinned
    cudaEvent_t ev_total = nullptr;
    bool collected = false;
    float collect_time = 0.f;

    int *d_flag = nullptr;                 // device scratch flag
    int *h_flag = nullptr;                 // pinned
    bool targets_finite = true;

    // CUDA-event timing rings: [class][slot][begin/end]; class 0 = integrate, 1 = flow splat
    static constexpr int kTimingSlots = 512;
    cudaEvent_t ev_ring[3][kTimingSlots][2] = {};   // 0 integrate (main stream), 1 flow splat, 2 noise (side stream)
    int64_t ev_count[3] = {0, 0, 0};

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

int ring_release(tb_ctx *c);
int bands_release(tb_ctx *c);

int alloc_flow(tb_ctx *c, int w, int h) {
    ring_release(c);
    bands_release(c);
    TB_REQUIRE(c, w >= 1 && h >= 1 && static_cast<long long>(w) * h < (1LL << 31), "flow grid dimensions out of bounds");
    if (c->flow) cudaFree(c->flow);
    if (c->seg) cudaFree(c->seg);
    if (c->hot) cudaFree(c->hot);
    c->flow = nullptr; c->seg = nullptr; c->hot = nullptr;
    c->W = w; c->H = h;
    const size_t G = static_cast<size_t>(w) * h;
    TB_CUDA(c, cudaMalloc(&c->flow, G * sizeof(float4)));
    TB_CUDA(c, cudaMalloc(&c->seg, 2 * G * sizeof(uint32_t)));
    TB_CUDA(c, cudaMalloc(&c->hot, (G + 2) * sizeof(uint32_t)));
    TB_CUDA(c, cudaMemsetAsync(c->flow, 0, G * sizeof(float4), c->stream));
    c->key_bits = 1;
    while ((1ull << c->key_bits) < G) c->key_bits += 1;
    c->collected = false;
    return TB_OK;
}

int ensure_frag_cap(tb_ctx *c, uint64_t need) {
    if (need <= c->frag_cap) return TB_OK;
    TB_REQUIRE(c, need < (1ull << 31), "flow splat: more than 2^31 fragments in one draw");
    if (c->frag_fixed)
        return fail(c, TB_ERR_UNSUPPORTED, "flow splat: " + std::to_string(need) + " fragments exceed the " +
                    std::to_string(c->frag_cap) + " reserved by tb_bands_export (the buffers are mapped by the other ranks); "
                    "export with a larger reserve and reconnect");
    uint64_t cap = std::max<uint64_t>(need + need / 4, 1u << 16);
    if (cap >= (1ull << 31)) cap = (1ull << 31) - 1;
    TB_CUDA(c, cudaStreamSynchronize(c->stream));
    for (int i = 0; i < 2; ++i) {
        if (c->keys[i]) cudaFree(c->keys[i]);
        if (c->vals[i]) cudaFree(c->vals[i]);
        c->keys[i] = nullptr; c->vals[i] = nullptr;
    }
    if (c->sort_tmp) cudaFree(c->sort_tmp);
    c->sort_tmp = nullptr;
    c->frag_cap = 0;
    for (int i = 0; i < 2; ++i) {
        TB_CUDA(c, cudaMalloc(&c->keys[i], cap * sizeof(uint32_t)));
        TB_CUDA(c, cudaMalloc(&c->vals[i], cap * sizeof(FragVal)));
    }
    c->sort_tmp_bytes = 0;
    TB_CUDA(c, cub::DeviceRadixSort::SortPairs(nullptr, c->sort_tmp_bytes, c->keys[0], c->keys[1], c->vals[0], c->vals[1],
                                               static_cast<int64_t>(cap), 0, 32, c->stream));
    TB_CUDA(c, cudaMalloc(&c->sort_tmp, c->sort_tmp_bytes));
    c->frag_cap = static_cast<uint32_t>(cap);
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

SplatArgs splat_args(tb_ctx *c) {
    SplatArgs A{};
    A.cur = c->buf[0];
    A.prev = c->buf[1];
    A.pairs = c->pairs;
    A.n_pairs = c->n_pairs;
    A.PH = c->PH;
    A.cols = c->col1 - c->col0;
    A.W = c->W;
    A.H = c->H;
    A.vsx = c->state.viewSize[0];
    A.vsy = c->state.viewSize[1];
    A.speedLimit = c->state.speedLimit;
    A.prim_off = c->prim_off;
    A.keys = c->keys[0];
    A.vals = c->vals[0];
    A.cap = c->frag_cap;
    A.total = c->prim_off + c->n_prims;      // slot n_prims holds the total after the scan
    return A;
}

// exclusive scan of the per-primitive counts in place and the total to the host (async; ev_total)
int scan_counts(tb_ctx *c) {
    const long long threads = c->n_prims;
    TB_CUDA(c, cub::DeviceScan::ExclusiveSum(c->scan_tmp, c->scan_tmp_bytes, c->prim_off, c->prim_off,
                                             static_cast<int>(threads + 1), c->stream));
    c->launches += 2;    // DeviceScanInitKernel + DeviceScanKernel
    TB_CUDA(c, cudaMemcpyAsync(c->h_total, c->prim_off + threads, sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
    TB_CUDA(c, cudaEventRecord(c->ev_total, c->stream));
    return TB_OK;
}

// Rasterise this context's primitives into per-texel fragment segments in draw order:
//   count per primitive -> exclusive scan -> emit in draw order (no atomics) -> stable radix
//   sort by texel -> segment bounds.
int collect(tb_ctx *c, float time) {
    TB_REQUIRE(c, c->have_state, "tb_set_state must be called before the flow splat");
    const size_t G = static_cast<size_t>(c->W) * c->H;
    const long long threads = c->n_prims;
    c->collect_time = time;
    c->collected = false;
    c->last_frags = 0;
    c->splat_since_step = true;
    // splat timing: from the start of the collect to the end of the fold (in a sharded run this
    // includes waiting for the grid from the previous rank)
    TB_CUDA(c, cudaEventRecord(c->ev_ring[1][c->ev_count[1] % tb_ctx::kTimingSlots][0], c->stream));
    if (threads <= 0) *c->h_total = 0;
    if (threads > 0) {
        const bool have_count = c->count_valid && c->count_wh[0] == c->W && c->count_wh[1] == c->H &&
                                c->count_vs[0] == c->state.viewSize[0] && c->count_vs[1] == c->state.viewSize[1];
        if (!have_count) {      // the count did not ride in k_integrate (or the draw parameters changed since)
            TB_CUDA(c, cudaMemsetAsync(c->prim_off, 0, (threads + 1) * sizeof(uint32_t), c->stream));
            SplatArgs A = splat_args(c);
            k_splat_count<<<blocks_for(threads, 256), 256, 0, c->stream>>>(A);
            if (int r = check_launch(c, "k_splat_count")) return r;
            if (int r = scan_counts(c)) return r;
        }
        c->count_valid = false;                                // consumed: the offsets are about to be used
        TB_CUDA(c, cudaEventSynchronize(c->ev_total));        // the sort needs the count on the host
        if (int r = ensure_frag_cap(c, *c->h_total)) return r;
        c->last_frags = *c->h_total;
    }
    const uint32_t F = *c->h_total;
    if (F > 0) {
        SplatArgs A = splat_args(c);
        k_splat_emit<<<blocks_for(threads, 256), 256, 0, c->stream>>>(A);
        if (int r = check_launch(c, "k_splat_emit")) return r;
        TB_CUDA(c, cub::DeviceRadixSort::SortPairs(c->sort_tmp, c->sort_tmp_bytes, c->keys[0], c->keys[1], c->vals[0],
                                                   c->vals[1], static_cast<int64_t>(F), 0, c->key_bits, c->stream));
        c->launches += 1 + (c->key_bits + 7) / 8;   // histogram + one onesweep pass per 8 key bits (+ scan, not counted)
        TB_CUDA(c, cudaMemsetAsync(c->seg, 0, 2 * G * sizeof(uint32_t), c->stream));
        k_splat_bounds<<<blocks_for((static_cast<long long>(F) + 3) / 4, 256), 256, 0, c->stream>>>(c->keys[1], F, c->seg);
        if (int r = check_launch(c, "k_splat_bounds")) return r;
    } else if (c->bands_connected) {
        TB_CUDA(c, cudaMemsetAsync(c->seg, 0, 2 * G * sizeof(uint32_t), c->stream));   // the other ranks read this table
    }
    c->collected = true;
    return TB_OK;
}

int fold(tb_ctx *c) {
    TB_REQUIRE(c, c->collected, "tb_splat_fold without a preceding tb_splat_collect");
    const int G = c->W * c->H;
    if (c->last_frags > 0) {
        FoldIO io{};
        io.src = c->flow; io.dst = c->flow; io.dst2 = nullptr;
        io.t_begin = 0; io.t_end = G; io.copy_all = 0;
        TB_CUDA(c, cudaMemsetAsync(c->hot, 0, 2 * sizeof(uint32_t), c->stream));       // [0] count, [1] cursor
        k_splat_fold<<<blocks_for(G, kFoldWarps * 32), kFoldWarps * 32, 0, c->stream>>>(
            io, reinterpret_cast<const uint2 *>(c->seg), c->vals[1], c->collect_time, c->hot, c->hot + 2, c->hot_threshold);
        if (int r = check_launch(c, "k_splat_fold")) return r;
        k_splat_fold_hot<<<c->n_sms * 4, kHotWarps * 32, 0, c->stream>>>(
            io, reinterpret_cast<const uint2 *>(c->seg), c->vals[1], c->collect_time, c->hot, c->hot + 2, c->hot + 1);
        if (int r = check_launch(c, "k_splat_fold_hot")) return r;
    }
    TB_CUDA(c, cudaEventRecord(c->ev_ring[1][c->ev_count[1] % tb_ctx::kTimingSlots][1], c->stream));
    c->ev_count[1] += 1;
    c->collected = false;
    return TB_OK;
}

// spawnShader target handling (src/particles.js:123-130): no buffer -> rotate and write
// buffers[0]; explicit targets FBO -> no rotation.  `particles` is buffers[1] either way.
float4 *spawn_out(tb_ctx *c, tb_target target) {
    if (target == TB_TARGET_TARGETS) return c->targets;
    c->count_valid = false;
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

}  // namespace

// ------------------------------------------------------------------------------------------------
// Sharded ring fold over peer memory
// ------------------------------------------------------------------------------------------------
namespace {

struct RingHandles {                 // what tb_ring_export hands to the neighbour (3 x 64 bytes + sizes)
    cudaIpcMemHandle_t flow, inbox, flags;
    int32_t w, h, chunks, pad;
};

int ring_release(tb_ctx *c) {
    if (c->next_inbox) cudaIpcCloseMemHandle(c->next_inbox);
    if (c->next_flow) cudaIpcCloseMemHandle(c->next_flow);
    if (c->next_flags) cudaIpcCloseMemHandle(c->next_flags);
    c->next_inbox = c->next_flow = nullptr;
    c->next_flags = nullptr;
    c->ring_connected = false;
    return TB_OK;
}

// The ordered fold of a column-sharded run, chunk by chunk:
//   rank r waits for chunk k of its inbox (rank 0: reads its own grid), folds its fragments onto it and
//   writes the result straight into rank r+1's inbox over NVLink, then raises that rank's flag;
//   the last rank's result is final: it goes to its own grid and to rank 0's, and travels on around the
//   ring (ring_stream) so that every rank ends the step with the same grid.
// Fold order = rank order = primitive order, so the result equals the single-GPU fold bit for bit.
int ring_fold(tb_ctx *c) {
    TB_REQUIRE(c, c->collected, "tb_splat_fold_ring without a preceding tb_splat_collect");
    TB_REQUIRE(c, c->ring_connected, "tb_ring_connect must be called first");
    const int G = c->W * c->H, C = c->ring_chunks;
    const int r = c->ring_rank, P = c->ring_world;
    const bool first = r == 0, last = r == P - 1;
    const uint32_t epoch = ++c->ring_epoch;
    const int per = ((G + C - 1) / C + 127) / 128 * 128;                 // texels per chunk, CTA aligned
    constexpr int S = tb_ctx::kRingMaxChunks;                              // flag stride: [0,S) inbox ready, [S,2S) final ready
    uint32_t *in_flag = c->ring_flags, *fin_flag = c->ring_flags + S;
    TB_CUDA(c, cudaMemsetAsync(c->ring_hot_counts, 0, 2 * tb_ctx::kRingMaxChunks * sizeof(uint32_t), c->stream));   // counts, cursors
    if (c->last_frags == 0)   // no fragments on this rank: collect() left the segments of an earlier draw behind
        TB_CUDA(c, cudaMemsetAsync(c->seg, 0, 2 * static_cast<size_t>(G) * sizeof(uint32_t), c->stream));
    TB_CUDA(c, cudaEventRecord(c->ev_ring_begin, c->stream));
    TB_CUDA(c, cudaStreamWaitEvent(c->ring_stream, c->ev_ring_begin, 0));
    // TB_RING_DEBUG=<epoch>: time the phases of that ring fold on every rank (diagnostics, stderr)
    static const int dbg_epoch = std::getenv("TB_RING_DEBUG") ? std::atoi(std::getenv("TB_RING_DEBUG")) : -1;
    const bool dbg = static_cast<int>(epoch) == dbg_epoch;
    static cudaEvent_t dbg_ev[tb_ctx::kRingMaxChunks][5];
    if (dbg)
        for (int k = 0; k < C; ++k)
            for (int j = 0; j < 5; ++j) cudaEventCreate(&dbg_ev[k][j]);
    for (int k = 0; k < C; ++k) {
        const int t0 = std::min(G, k * per), t1 = std::min(G, (k + 1) * per);
        if (t0 >= t1) continue;
        FoldIO io{};
        if (dbg) cudaEventRecord(dbg_ev[k][0], c->stream);
        io.src = first ? c->flow : c->inbox;
        io.dst = last ? c->flow : c->next_inbox;
        io.dst2 = (last && P > 1) ? c->next_flow : nullptr;             // rank 0's grid
        io.t_begin = t0; io.t_end = t1; io.copy_all = 1;
        if (!first) {
            k_ring_wait<<<1, 1, 0, c->stream>>>(in_flag + k, epoch);
            if (int e = check_launch(c, "k_ring_wait")) return e;
        }
        if (dbg) cudaEventRecord(dbg_ev[k][1], c->stream);
        k_splat_fold<<<blocks_for(t1 - t0, kFoldWarps * 32), kFoldWarps * 32, 0, c->stream>>>(
            io, reinterpret_cast<const uint2 *>(c->seg), c->vals[1], c->collect_time, c->ring_hot_counts + k, c->hot + 2,
            c->last_frags > 0 ? c->hot_threshold : 0xffffffffu);
        if (int e = check_launch(c, "k_splat_fold")) return e;
        if (dbg) cudaEventRecord(dbg_ev[k][2], c->stream);
        k_splat_fold_hot<<<c->n_sms * 4, kHotWarps * 32, 0, c->stream>>>(
            io, reinterpret_cast<const uint2 *>(c->seg), c->vals[1], c->collect_time, c->ring_hot_counts + k, c->hot + 2,
            c->ring_hot_counts + tb_ctx::kRingMaxChunks + k);
        if (int e = check_launch(c, "k_splat_fold_hot")) return e;
        // chunk k of the next rank's inbox (or, from the last rank, of rank 0's grid) is complete
        if (dbg) cudaEventRecord(dbg_ev[k][3], c->stream);
        k_ring_signal<<<1, 1, 0, c->stream>>>(last ? c->next_flags + S + k : c->next_flags + k, epoch);
        if (int e = check_launch(c, "k_ring_signal")) return e;
        if (dbg) cudaEventRecord(dbg_ev[k][4], c->stream);
    }
    if (dbg) {
        cudaStreamSynchronize(c->stream);
        for (int k = 0; k < C; ++k) {
            float w = 0, m = 0, h = 0, sgl = 0;
            cudaEventElapsedTime(&w, dbg_ev[k][0], dbg_ev[k][1]);
            cudaEventElapsedTime(&m, dbg_ev[k][1], dbg_ev[k][2]);
            cudaEventElapsedTime(&h, dbg_ev[k][2], dbg_ev[k][3]);
            cudaEventElapsedTime(&sgl, dbg_ev[k][3], dbg_ev[k][4]);
            std::fprintf(stderr, "[ring dbg] rank %d chunk %d: wait %.0f us, fold %.0f us, hot %.0f us, signal %.0f us (frags %lld)\n",
                         r, k, w * 1e3f, m * 1e3f, h * 1e3f, sgl * 1e3f, static_cast<long long>(c->last_frags));
        }
    }
    // final chunks travel 0 -> 1 -> ... -> P-2 (the last rank already has them)
    if (!last) {
        for (int k = 0; k < C; ++k) {
            const int t0 = std::min(G, k * per), t1 = std::min(G, (k + 1) * per);
            if (t0 >= t1) continue;
            k_ring_wait<<<1, 1, 0, c->ring_stream>>>(fin_flag + k, epoch);
            if (int e = check_launch(c, "k_ring_wait")) return e;
            if (r < P - 2) {
                TB_CUDA(c, cudaMemcpyAsync(c->next_flow + t0, c->flow + t0, static_cast<size_t>(t1 - t0) * sizeof(float4),
                                           cudaMemcpyDeviceToDevice, c->ring_stream));
                k_ring_signal<<<1, 1, 0, c->ring_stream>>>(c->next_flags + S + k, epoch);
                if (int e = check_launch(c, "k_ring_signal")) return e;
            }
        }
        TB_CUDA(c, cudaEventRecord(c->ev_ring_fwd, c->ring_stream));
        TB_CUDA(c, cudaStreamWaitEvent(c->stream, c->ev_ring_fwd, 0));    // the next integrate reads the final grid
    }
    TB_CUDA(c, cudaEventRecord(c->ev_ring[1][c->ev_count[1] % tb_ctx::kTimingSlots][1], c->stream));
    c->ev_count[1] += 1;
    c->collected = false;
    return TB_OK;
}

}  // namespace

extern "C" {

int tb_ring_export(tb_ctx *c, void *out, int64_t n_bytes) {
    TB_REQUIRE(c, c && out, "null argument");
    TB_REQUIRE(c, n_bytes == static_cast<int64_t>(sizeof(RingHandles)), "tb_ring_export: buffer must be tb_ring_handle_bytes() long");
    TB_CUDA(c, cudaSetDevice(c->device));
    const size_t G = static_cast<size_t>(c->W) * c->H;
    if (const char *e = std::getenv("TB_RING_CHUNKS")) c->ring_chunks = std::max(1, std::min<int>(tb_ctx::kRingMaxChunks, std::atoi(e)));
    else c->ring_chunks = 0;      // chosen from the world size in tb_ring_connect
    const int C = tb_ctx::kRingMaxChunks;
    TB_CUDA(c, cudaStreamSynchronize(c->stream));
    ring_release(c);
    if (c->inbox) cudaFree(c->inbox);
    if (c->ring_flags) cudaFree(c->ring_flags);
    if (c->ring_hot_counts) cudaFree(c->ring_hot_counts);
    c->inbox = nullptr; c->ring_flags = nullptr; c->ring_hot_counts = nullptr;
    TB_CUDA(c, cudaMalloc(&c->inbox, G * sizeof(float4)));
    TB_CUDA(c, cudaMalloc(&c->ring_flags, 2 * C * sizeof(uint32_t)));
    TB_CUDA(c, cudaMalloc(&c->ring_hot_counts, 2 * C * sizeof(uint32_t)));
    TB_CUDA(c, cudaMemset(c->ring_flags, 0, 2 * C * sizeof(uint32_t)));
    c->ring_epoch = 0;
    if (!c->ring_stream) {
        TB_CUDA(c, cudaStreamCreateWithFlags(&c->ring_stream, cudaStreamNonBlocking));
        TB_CUDA(c, cudaEventCreateWithFlags(&c->ev_ring_fwd, cudaEventDisableTiming));
        TB_CUDA(c, cudaEventCreateWithFlags(&c->ev_ring_begin, cudaEventDisableTiming));
    }
    RingHandles hnd{};
    TB_CUDA(c, cudaIpcGetMemHandle(&hnd.flow, c->flow));
    TB_CUDA(c, cudaIpcGetMemHandle(&hnd.inbox, c->inbox));
    TB_CUDA(c, cudaIpcGetMemHandle(&hnd.flags, c->ring_flags));
    hnd.w = c->W; hnd.h = c->H; hnd.chunks = c->ring_chunks;
    std::memcpy(out, &hnd, sizeof(hnd));
    return TB_OK;
}

int64_t tb_ring_handle_bytes(void) { return static_cast<int64_t>(sizeof(RingHandles)); }

int tb_ring_connect(tb_ctx *c, int32_t rank, int32_t world, const void *next_rank_handles, int64_t n_bytes) {
    TB_REQUIRE(c, c && next_rank_handles, "null argument");
    TB_REQUIRE(c, n_bytes == static_cast<int64_t>(sizeof(RingHandles)), "tb_ring_connect: bad handle size");
    TB_REQUIRE(c, world >= 2 && rank >= 0 && rank < world, "tb_ring_connect: bad rank/world");
    TB_REQUIRE(c, c->inbox != nullptr, "tb_ring_export must be called before tb_ring_connect");
    TB_CUDA(c, cudaSetDevice(c->device));
    RingHandles hnd;
    std::memcpy(&hnd, next_rank_handles, sizeof(hnd));
    TB_REQUIRE(c, hnd.w == c->W && hnd.h == c->H && hnd.chunks == c->ring_chunks, "tb_ring_connect: the next rank's flow grid has another shape");
    // Measured on 8xB200 (profiles/r01_multi_gpu.txt): every chunk adds the tail of its own hot texels, so few
    // chunks win on short rings and about world/2 on long ones.
    if (c->ring_chunks == 0) c->ring_chunks = std::max(1, world / 2);
    ring_release(c);
    TB_CUDA(c, cudaIpcOpenMemHandle(reinterpret_cast<void **>(&c->next_flow), hnd.flow, cudaIpcMemLazyEnablePeerAccess));
    TB_CUDA(c, cudaIpcOpenMemHandle(reinterpret_cast<void **>(&c->next_inbox), hnd.inbox, cudaIpcMemLazyEnablePeerAccess));
    TB_CUDA(c, cudaIpcOpenMemHandle(reinterpret_cast<void **>(&c->next_flags), hnd.flags, cudaIpcMemLazyEnablePeerAccess));
    c->ring_rank = rank; c->ring_world = world;
    c->ring_connected = true;
    // a rank waiting for its predecessor's grid has idle SMs: let the next step's noise run there
    c->overlap = true;
    return TB_OK;
}

}  // extern "C"

// ------------------------------------------------------------------------------------------------
// Band fold over peer memory: all ranks fold in parallel, each its own tiles of the grid
// ------------------------------------------------------------------------------------------------
namespace {

struct BandHandles {                 // what tb_bands_export hands to every other rank
    cudaIpcMemHandle_t seg, merged, dst, flow, flags;
    int32_t w, h;
    uint32_t frag_cap, pad;
};

int bands_release(tb_ctx *c) {
    for (int j = 0; j < kMaxBandRanks; ++j) {
        if (j != c->bands_rank || !c->bands_connected) {
            if (c->band_src.seg[j]) cudaIpcCloseMemHandle(const_cast<uint2 *>(c->band_src.seg[j]));
            if (c->band_sinks.merged[j]) cudaIpcCloseMemHandle(c->band_sinks.merged[j]);
            if (c->band_sinks.dst[j]) cudaIpcCloseMemHandle(c->band_sinks.dst[j]);
            if (c->band_peers.flow[j]) cudaIpcCloseMemHandle(c->band_peers.flow[j]);
            if (c->band_peers.flags[j]) cudaIpcCloseMemHandle(c->band_peers.flags[j]);
        }
        c->band_src.seg[j] = nullptr;
        c->band_sinks.merged[j] = nullptr; c->band_sinks.dst[j] = nullptr;
        c->band_peers.flow[j] = nullptr; c->band_peers.flags[j] = nullptr;
    }
    cudaFree(c->bands_len); cudaFree(c->bands_off); cudaFree(c->bands_seg); cudaFree(c->bands_scan_tmp);
    c->bands_len = c->bands_off = c->bands_seg = nullptr;
    c->bands_scan_tmp = nullptr;
    c->bands_connected = false;
    c->frag_fixed = false;
    return TB_OK;
}

// Every rank folds the tiles it owns (tile % world == rank).  The fragments of those tiles are first brought
// together in one local array -- per texel the sources side by side in rank order = column order = primitive
// order, and inside a source the stable sort kept the draw order -- then the local fold runs on it.
// The result equals the single-GPU fold bit for bit, and no rank waits for another rank's fold.
int bands_fold(tb_ctx *c) {
    TB_REQUIRE(c, c->collected, "tb_splat_fold_bands without a preceding tb_splat_collect");
    TB_REQUIRE(c, c->bands_connected, "tb_bands_connect must be called first");
    const int G = c->W * c->H, P = c->bands_world, r = c->bands_rank, mine = c->bands_mine;
    if (c->bands_epoch > 0) {                        // the previous fold's overflow flag has long arrived
        TB_CUDA(c, cudaEventSynchronize(c->ev_bands));
        if (*c->h_bands_overflow)
            return fail(c, TB_ERR_UNSUPPORTED, "flow splat (bands): the fragments of this rank's tiles exceeded the reserve of " +
                        std::to_string(c->frag_cap) + "; export with a larger reserve (TB_BANDS_RESERVE) and reconnect");
    }
    const uint32_t epoch = ++c->bands_epoch;
    // TB_RING_DEBUG=<epoch>: time the phases of that fold on every rank (diagnostics, stderr)
    static const int dbg_epoch = std::getenv("TB_RING_DEBUG") ? std::atoi(std::getenv("TB_RING_DEBUG")) : -1;
    const bool dbg = static_cast<int>(epoch) == dbg_epoch;
    static cudaEvent_t dbg_ev[12];
    int dbg_n = 0;
    auto mark = [&]() { if (dbg) { cudaEventCreate(&dbg_ev[dbg_n]); cudaEventRecord(dbg_ev[dbg_n++], c->stream); } };
    auto barrier = [&](int phase) -> int {
        k_bands_barrier<<<1, 32, 0, c->stream>>>(c->band_peers, c->bands_flags, phase, epoch);
        return check_launch(c, "k_bands_barrier");
    };
    mark();
    // barrier 0: every rank's sorted fragments and segment table of this step are in place (and every rank is done
    // with the merged array and the grid of the previous step)
    if (int e = barrier(0)) return e;
    mark();
    const long long warps = static_cast<long long>(mine) * P;
    const int n_seg = mine * 32 * P + 1;
    k_bands_lengths<<<blocks_for(warps * 32, 256), 256, 0, c->stream>>>(c->band_src, P, r, mine, G, c->bands_len);
    if (int e = check_launch(c, "k_bands_lengths")) return e;
    TB_CUDA(c, cub::DeviceScan::ExclusiveSum(c->bands_scan_tmp, c->bands_scan_bytes, c->bands_len, c->bands_off, n_seg, c->stream));
    c->launches += 2;
    k_bands_offsets<<<blocks_for(warps * 32, 256), 256, 0, c->stream>>>(c->band_sinks, P, r, mine, G, c->bands_off, c->frag_cap,
                                                                           c->bands_seg, c->bands_overflow);
    if (int e = check_launch(c, "k_bands_offsets")) return e;
    TB_CUDA(c, cudaMemcpyAsync(c->h_bands_overflow, c->bands_overflow, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    TB_CUDA(c, cudaEventRecord(c->ev_bands, c->stream));
    mark();
    // barrier 1: every owner has told every source where its segments go
    if (int e = barrier(1)) return e;
    mark();
    if (c->last_frags > 0) {
        const uint32_t F = static_cast<uint32_t>(c->last_frags);
        k_bands_push<<<blocks_for(F, 256), 256, 0, c->stream>>>(c->keys[1], c->vals[1], F, reinterpret_cast<const uint2 *>(c->seg),
                                                                  c->bands_dst, c->band_sinks, P, c->frag_cap);
        if (int e = check_launch(c, "k_bands_push")) return e;
    }
    mark();
    // barrier 2: every source's fragments have landed in the owners' merged arrays
    if (int e = barrier(2)) return e;
    mark();
    FoldIO io{};
    io.src = c->flow; io.dst = c->flow; io.dst2 = nullptr;
    io.t_begin = 0; io.t_end = G; io.copy_all = 0;
    io.tile_first = r; io.tile_stride = P;
    TB_CUDA(c, cudaMemsetAsync(c->hot, 0, 2 * sizeof(uint32_t), c->stream));       // [0] count, [1] cursor
    k_splat_fold<<<blocks_for(mine, kFoldWarps), kFoldWarps * 32, 0, c->stream>>>(
        io, reinterpret_cast<const uint2 *>(c->bands_seg), c->vals[0], c->collect_time, c->hot, c->hot + 2, c->hot_threshold);
    if (int e = check_launch(c, "k_splat_fold")) return e;
    k_splat_fold_hot<<<c->n_sms * 4, kHotWarps * 32, 0, c->stream>>>(
        io, reinterpret_cast<const uint2 *>(c->bands_seg), c->vals[0], c->collect_time, c->hot, c->hot + 2, c->hot + 1);
    if (int e = check_launch(c, "k_splat_fold_hot")) return e;
    mark();
    k_bands_publish<<<blocks_for(static_cast<long long>(mine) * 32, 256), 256, 0, c->stream>>>(c->flow, c->band_peers, G);
    if (int e = check_launch(c, "k_bands_publish")) return e;
    mark();
    // barrier 3: every tile has landed in every grid
    if (int e = barrier(3)) return e;
    mark();
    if (dbg) {
        cudaStreamSynchronize(c->stream);
        float ms[8] = {};
        for (int k = 0; k < 8; ++k) cudaEventElapsedTime(&ms[k], dbg_ev[k], dbg_ev[k + 1]);
        std::fprintf(stderr, "[bands dbg] rank %d: barrier0 %.0f us, lengths+scan+offsets %.0f us, barrier1 %.0f us, push %.0f us, "
                             "barrier2 %.0f us, fold+hot %.0f us, publish %.0f us, barrier3 %.0f us (own frags %lld)\n",
                     r, 1e3 * ms[0], 1e3 * ms[1], 1e3 * ms[2], 1e3 * ms[3], 1e3 * ms[4], 1e3 * ms[5], 1e3 * ms[6], 1e3 * ms[7],
                     static_cast<long long>(c->last_frags));
        for (int k = 0; k < 9; ++k) cudaEventDestroy(dbg_ev[k]);
    }
    TB_CUDA(c, cudaEventRecord(c->ev_ring[1][c->ev_count[1] % tb_ctx::kTimingSlots][1], c->stream));
    c->ev_count[1] += 1;
    c->collected = false;
    return TB_OK;
}

}  // namespace

extern "C" {

int64_t tb_bands_handle_bytes(void) { return static_cast<int64_t>(sizeof(BandHandles)); }

int tb_bands_export(tb_ctx *c, int64_t reserve_fragments, void *out, int64_t n_bytes) {
    TB_REQUIRE(c, c && out, "null argument");
    TB_REQUIRE(c, n_bytes == static_cast<int64_t>(sizeof(BandHandles)), "tb_bands_export: buffer must be tb_bands_handle_bytes() long");
    TB_REQUIRE(c, reserve_fragments >= 0 && reserve_fragments < (1ll << 31), "tb_bands_export: reserve out of range");
    TB_CUDA(c, cudaSetDevice(c->device));
    TB_CUDA(c, cudaStreamSynchronize(c->stream));
    bands_release(c);
    if (int e = ensure_frag_cap(c, std::max<uint64_t>(static_cast<uint64_t>(reserve_fragments), 1u << 16))) return e;
    if (!c->bands_flags) {
        TB_CUDA(c, cudaMalloc(&c->bands_flags, kBandPhases * kMaxBandRanks * sizeof(uint32_t)));
        TB_CUDA(c, cudaMalloc(&c->bands_overflow, sizeof(int)));
        TB_CUDA(c, cudaMallocHost(&c->h_bands_overflow, sizeof(int)));
        TB_CUDA(c, cudaEventCreateWithFlags(&c->ev_bands, cudaEventDisableTiming));
    }
    TB_CUDA(c, cudaMemset(c->bands_overflow, 0, sizeof(int)));
    *c->h_bands_overflow = 0;
    TB_CUDA(c, cudaMemset(c->bands_flags, 0, kBandPhases * kMaxBandRanks * sizeof(uint32_t)));
    c->bands_epoch = 0;
    if (c->bands_dst) cudaFree(c->bands_dst);
    c->bands_dst = nullptr;
    TB_CUDA(c, cudaMalloc(&c->bands_dst, static_cast<size_t>(c->W) * c->H * sizeof(uint32_t)));
    TB_CUDA(c, cudaMemset(c->bands_dst, 0, static_cast<size_t>(c->W) * c->H * sizeof(uint32_t)));
    BandHandles hnd{};
    TB_CUDA(c, cudaIpcGetMemHandle(&hnd.seg, c->seg));
    TB_CUDA(c, cudaIpcGetMemHandle(&hnd.merged, c->vals[0]));
    TB_CUDA(c, cudaIpcGetMemHandle(&hnd.dst, c->bands_dst));
    TB_CUDA(c, cudaIpcGetMemHandle(&hnd.flow, c->flow));
    TB_CUDA(c, cudaIpcGetMemHandle(&hnd.flags, c->bands_flags));
    hnd.w = c->W; hnd.h = c->H; hnd.frag_cap = c->frag_cap;
    std::memcpy(out, &hnd, sizeof(hnd));
    c->frag_fixed = true;
    return TB_OK;
}

int tb_bands_connect(tb_ctx *c, int32_t rank, int32_t world, const void *all_handles, int64_t n_bytes) {
    TB_REQUIRE(c, c && all_handles, "null argument");
    TB_REQUIRE(c, world >= 2 && world <= kMaxBandRanks && rank >= 0 && rank < world, "tb_bands_connect: bad rank/world");
    TB_REQUIRE(c, n_bytes == static_cast<int64_t>(sizeof(BandHandles)) * world, "tb_bands_connect: expected world x tb_bands_handle_bytes()");
    TB_REQUIRE(c, c->frag_fixed && c->bands_flags, "tb_bands_export must be called before tb_bands_connect");
    TB_CUDA(c, cudaSetDevice(c->device));
    const bool fixed = c->frag_fixed;
    bands_release(c);
    c->frag_fixed = fixed;
    c->bands_rank = rank; c->bands_world = world;
    c->bands_connected = true;                 // from here on bands_release skips this rank's own slots
    const auto *h = static_cast<const unsigned char *>(all_handles);
    for (int j = 0; j < world; ++j) {
        if (j == rank) {
            c->band_src.seg[j] = reinterpret_cast<const uint2 *>(c->seg);
            c->band_sinks.merged[j] = c->vals[0];
            c->band_sinks.dst[j] = c->bands_dst;
            c->band_peers.flow[j] = c->flow;
            c->band_peers.flags[j] = c->bands_flags;
            continue;
        }
        BandHandles hnd;
        std::memcpy(&hnd, h + sizeof(BandHandles) * j, sizeof(hnd));
        if (hnd.w != c->W || hnd.h != c->H) {
            bands_release(c);
            return fail(c, TB_ERR_INVALID, "tendrils-b200: tb_bands_connect: rank " + std::to_string(j) + " has another flow grid shape");
        }
        void *p[5] = {};
        const cudaIpcMemHandle_t *hs[5] = {&hnd.seg, &hnd.merged, &hnd.dst, &hnd.flow, &hnd.flags};
        for (int k = 0; k < 5; ++k) {
            cudaError_t e = cudaIpcOpenMemHandle(&p[k], *hs[k], cudaIpcMemLazyEnablePeerAccess);
            // keep what was opened so far where bands_release will find it
            if (k == 0) c->band_src.seg[j] = static_cast<const uint2 *>(p[0]);
            if (k == 1) c->band_sinks.merged[j] = static_cast<FragVal *>(p[1]);
            if (k == 2) c->band_sinks.dst[j] = static_cast<uint32_t *>(p[2]);
            if (k == 3) c->band_peers.flow[j] = static_cast<float4 *>(p[3]);
            if (k == 4) c->band_peers.flags[j] = static_cast<uint32_t *>(p[4]);
            if (e != cudaSuccess) {
                bands_release(c);
                return fail(c, TB_ERR_CUDA, std::string("cudaIpcOpenMemHandle (rank ") + std::to_string(j) + "): " + cudaGetErrorString(e));
            }
        }
    }
    c->band_peers.n = world; c->band_peers.me = rank;
    // barriers, NVLink-bound pushes and latency-bound blend chains leave SMs idle: let the next step's noise run there
    c->overlap = true;
    // scratch of the owner side: lengths / offsets per (local texel, source), the merged segment table
    const size_t G = static_cast<size_t>(c->W) * c->H;
    const int tiles = static_cast<int>((G + 31) / 32);
    c->bands_mine = (tiles - rank + world - 1) / world;
    const size_t n_seg = static_cast<size_t>(c->bands_mine) * 32 * world + 1;
    TB_CUDA(c, cudaMalloc(&c->bands_len, n_seg * sizeof(uint32_t)));
    TB_CUDA(c, cudaMalloc(&c->bands_off, n_seg * sizeof(uint32_t)));
    TB_CUDA(c, cudaMalloc(&c->bands_seg, 2 * G * sizeof(uint32_t)));
    TB_CUDA(c, cudaMemset(c->bands_len, 0, n_seg * sizeof(uint32_t)));
    c->bands_scan_bytes = 0;
    TB_CUDA(c, cub::DeviceScan::ExclusiveSum(nullptr, c->bands_scan_bytes, c->bands_len, c->bands_off, static_cast<int>(n_seg), c->stream));
    TB_CUDA(c, cudaMalloc(&c->bands_scan_tmp, std::max<size_t>(c->bands_scan_bytes, 16)));
    return TB_OK;
}

int tb_splat_fold_bands(tb_ctx *c) {
    TB_REQUIRE(c, c, "null context");
    TB_CUDA(c, cudaSetDevice(c->device));
    return bands_fold(c);
}

}  // extern "C"

extern "C" {

// ---- band exchange ("a2a"): every rank folds ONE band of the grid with the fragments of ALL ranks -------------
int tb_splat_band_offsets(tb_ctx *c, int32_t n_bands, int32_t band_texels, int64_t *host_offsets) {
    TB_REQUIRE(c, c && host_offsets, "null argument");
    TB_REQUIRE(c, c->collected, "tb_splat_band_offsets without a preceding tb_splat_collect");
    TB_REQUIRE(c, n_bands >= 1 && n_bands <= 64, "bad band count");
    TB_CUDA(c, cudaSetDevice(c->device));
    const uint32_t F = static_cast<uint32_t>(c->last_frags);
    uint32_t *d_out = c->hot;                                  // scratch: free between collect and fold
    if (F > 0) {
        k_band_offsets<<<1, 128, 0, c->stream>>>(c->keys[1], F, band_texels, n_bands, d_out);
        if (int r = check_launch(c, "k_band_offsets")) return r;
        uint32_t tmp[65];
        TB_CUDA(c, cudaMemcpyAsync(tmp, d_out, (n_bands + 1) * sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
        TB_CUDA(c, cudaStreamSynchronize(c->stream));
        for (int b = 0; b <= n_bands; ++b) host_offsets[b] = tmp[b];
    } else {
        for (int b = 0; b <= n_bands; ++b) host_offsets[b] = 0;
    }
    return TB_OK;
}

// pointers for the exchange: send = the sorted fragments, recv = the (now free) pre-sort buffers, grown if needed
int tb_splat_exchange_buffers(tb_ctx *c, int64_t recv_items, void **send_keys, void **send_vals, void **recv_keys,
                              void **recv_vals, int32_t *val_bytes) {
    TB_REQUIRE(c, c && send_keys && send_vals && recv_keys && recv_vals && val_bytes, "null argument");
    TB_CUDA(c, cudaSetDevice(c->device));
    if (static_cast<uint64_t>(recv_items) > c->frag_cap) {
        // growing reallocates all four buffers: keep the sorted fragments
        const uint32_t F = static_cast<uint32_t>(c->last_frags);
        uint32_t *k_old = nullptr; FragVal *v_old = nullptr;
        TB_CUDA(c, cudaStreamSynchronize(c->stream));
        if (F > 0) {
            TB_CUDA(c, cudaMalloc(&k_old, F * sizeof(uint32_t)));
            TB_CUDA(c, cudaMalloc(&v_old, F * sizeof(FragVal)));
            TB_CUDA(c, cudaMemcpy(k_old, c->keys[1], F * sizeof(uint32_t), cudaMemcpyDeviceToDevice));
            TB_CUDA(c, cudaMemcpy(v_old, c->vals[1], F * sizeof(FragVal), cudaMemcpyDeviceToDevice));
        }
        if (int r = ensure_frag_cap(c, static_cast<uint64_t>(recv_items))) return r;
        if (F > 0) {
            TB_CUDA(c, cudaMemcpy(c->keys[1], k_old, F * sizeof(uint32_t), cudaMemcpyDeviceToDevice));
            TB_CUDA(c, cudaMemcpy(c->vals[1], v_old, F * sizeof(FragVal), cudaMemcpyDeviceToDevice));
            cudaFree(k_old); cudaFree(v_old);
        }
    }
    *send_keys = c->keys[1]; *send_vals = c->vals[1];
    *recv_keys = c->keys[0]; *recv_vals = c->vals[0];
    *val_bytes = static_cast<int32_t>(sizeof(FragVal));
    return TB_OK;
}

// blend one received piece (fragments of ONE source rank, sorted, in draw order) onto texels [t_begin, t_end)
int tb_splat_fold_piece(tb_ctx *c, int64_t piece_offset, int64_t piece_items, int32_t t_begin, int32_t t_end) {
    TB_REQUIRE(c, c, "null context");
    TB_REQUIRE(c, t_begin >= 0 && t_begin <= t_end && t_end <= c->W * c->H, "bad texel range");
    TB_REQUIRE(c, piece_offset >= 0 && piece_items >= 0 && piece_offset + piece_items <= static_cast<int64_t>(c->frag_cap), "bad piece");
    TB_CUDA(c, cudaSetDevice(c->device));
    if (piece_items == 0 || t_begin == t_end) return TB_OK;
    const uint32_t n = static_cast<uint32_t>(piece_items);
    const uint32_t *keys = c->keys[0] + piece_offset;
    const FragVal *vals = c->vals[0] + piece_offset;
    TB_CUDA(c, cudaMemsetAsync(c->seg + 2ull * t_begin, 0, 2ull * (t_end - t_begin) * sizeof(uint32_t), c->stream));
    k_splat_bounds<<<blocks_for((static_cast<long long>(n) + 3) / 4, 256), 256, 0, c->stream>>>(keys, n, c->seg);
    if (int r = check_launch(c, "k_splat_bounds")) return r;
    FoldIO io{};
    io.src = c->flow; io.dst = c->flow; io.dst2 = nullptr;
    io.t_begin = t_begin; io.t_end = t_end; io.copy_all = 0;
    TB_CUDA(c, cudaMemsetAsync(c->hot, 0, 2 * sizeof(uint32_t), c->stream));
    k_splat_fold<<<blocks_for(t_end - t_begin, kFoldWarps * 32), kFoldWarps * 32, 0, c->stream>>>(
        io, reinterpret_cast<const uint2 *>(c->seg), vals, c->collect_time, c->hot, c->hot + 2, c->hot_threshold);
    if (int r = check_launch(c, "k_splat_fold")) return r;
    k_splat_fold_hot<<<c->n_sms * 4, kHotWarps * 32, 0, c->stream>>>(
        io, reinterpret_cast<const uint2 *>(c->seg), vals, c->collect_time, c->hot, c->hot + 2, c->hot + 1);
    if (int r = check_launch(c, "k_splat_fold_hot")) return r;
    return TB_OK;
}

// closes the splat timing span of a band-exchange draw and marks the collect as consumed
int tb_splat_exchange_done(tb_ctx *c) {
    TB_REQUIRE(c, c, "null context");
    TB_CUDA(c, cudaSetDevice(c->device));
    TB_CUDA(c, cudaEventRecord(c->ev_ring[1][c->ev_count[1] % tb_ctx::kTimingSlots][1], c->stream));
    c->ev_count[1] += 1;
    c->collected = false;
    return TB_OK;
}

int tb_splat_fold_ring(tb_ctx *c) {
    TB_REQUIRE(c, c, "null context");
    TB_CUDA(c, cudaSetDevice(c->device));
    return ring_fold(c);
}

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
    // Opt-in (TB_OVERLAP=1): measured +4.7 % step throughput at cfg3, but the low-priority noise launch is
    // time-sliced under the sort, which makes its own duration meaningless as a roofline input.
    c->overlap = std::getenv("TB_OVERLAP") != nullptr;
    if (const char *e = std::getenv("TB_FOLD_HOT")) c->hot_threshold = static_cast<uint32_t>(std::max(1, std::atoi(e)));
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
    TB_TRY(cudaMallocHost(&c->h_total, sizeof(uint32_t)));
    TB_TRY(cudaEventCreateWithFlags(&c->ev_total, cudaEventDisableTiming));
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
    {
        // Row -> pair map of the count fused into k_integrate.  A pair rides there if both its vertices read ONE texel row
        // (one particle's previous and current state) and no earlier pair has claimed that row: the D6 table of many
        // non-power-of-two heights draws some rows twice (e.g. PH = 47: pairs 0 and 1 both draw row 0), and tall
        // textures have a few pairs that join two different rows (3 of 4104 at PH = 8192).
        std::vector<int32_t> rp(static_cast<size_t>(PH), -1), odd;
        for (size_t k = 0; k < pairs.size(); ++k) {
            const int ra = pairs[k].row_a & 0x7fffffff, rb = pairs[k].row_b & 0x7fffffff;
            const bool ca = pairs[k].row_a < 0, cb = pairs[k].row_b < 0;
            const bool rides = ra == rb && ca != cb && k < (1u << 30) && rp[static_cast<size_t>(ra)] == -1;
            if (rides) rp[static_cast<size_t>(ra)] = static_cast<int32_t>(static_cast<uint32_t>(k) | ((cb ? 1u : 2u) << 30));   // 1: prev->cur, 2: cur->prev
            else odd.push_back(static_cast<int32_t>(k));
        }
        c->fuse_count = !pairs.empty() && odd.empty();
        // experimental, off unless TB_FUSE_PARTIAL is set (not yet run on a GPU): ride anyway when only a few pairs cannot
        c->fuse_partial = std::getenv("TB_FUSE_PARTIAL") != nullptr && !pairs.empty() && !odd.empty() && odd.size() * 16 <= pairs.size();
        c->n_odd = static_cast<int>(odd.size());
        if (c->fuse_partial) {
            TB_TRY(cudaMalloc(&c->odd_pairs, odd.size() * sizeof(int32_t)));
            TB_TRY(cudaMemcpyAsync(c->odd_pairs, odd.data(), odd.size() * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
        }
        TB_TRY(cudaMalloc(&c->row_pair, rp.size() * sizeof(int32_t)));
        TB_TRY(cudaMemcpyAsync(c->row_pair, rp.data(), rp.size() * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
        TB_TRY(cudaStreamSynchronize(c->stream));
    }
    c->n_prims = static_cast<long long>(col1 - col0) * c->n_pairs;
    if (c->n_prims >= (1LL << 31)) { c->err = "tendrils-b200: too many primitives per context"; return bail(TB_ERR_INVALID); }
    TB_TRY(cudaMalloc(&c->prim_off, (c->n_prims + 1) * sizeof(uint32_t)));
    TB_TRY(cub::DeviceScan::ExclusiveSum(nullptr, c->scan_tmp_bytes, c->prim_off, c->prim_off,
                                         static_cast<int>(c->n_prims + 1), c->stream));
    TB_TRY(cudaMalloc(&c->scan_tmp, c->scan_tmp_bytes));
#undef TB_TRY
    const int fw = cfg->flow_w > 0 ? cfg->flow_w : 1, fh = cfg->flow_h > 0 ? cfg->flow_h : 1;
    if (int r = alloc_flow(c, fw, fh)) return bail(r);
    if (int r = ensure_frag_cap(c, static_cast<uint64_t>(c->n_prims) * 2)) return bail(r);
    if (int r = tb_reset(c)) return bail(r);
    *out = c;
    return TB_OK;
}

int tb_destroy(tb_ctx *c) {
    if (!c) return TB_OK;
    cudaSetDevice(c->device);
    if (c->side) cudaStreamSynchronize(c->side);
    if (c->stream) cudaStreamSynchronize(c->stream);
    cudaFree(c->buf[0]); cudaFree(c->buf[1]); cudaFree(c->targets); cudaFree(c->flow);
    cudaFree(c->frames);
    cudaFree(c->line_attr); cudaFree(c->line_verts); cudaFree(c->line_bbox);
    cudaFree(c->image); cudaFree(c->layer); cudaFree(c->pairs); cudaFree(c->prim_off); cudaFree(c->row_pair); cudaFree(c->odd_pairs);
    cudaFree(c->scan_tmp); cudaFree(c->sort_tmp); cudaFree(c->seg); cudaFree(c->hot); cudaFree(c->d_flag);
    for (int i = 0; i < 2; ++i) { cudaFree(c->keys[i]); cudaFree(c->vals[i]); }
    if (c->h_flag) cudaFreeHost(c->h_flag);
    if (c->h_total) cudaFreeHost(c->h_total);
    if (c->ev_total) cudaEventDestroy(c->ev_total);
    for (int k = 0; k < 3; ++k)
        for (int i = 0; i < tb_ctx::kTimingSlots; ++i)
            for (int j = 0; j < 2; ++j)
                if (c->ev_ring[k][i][j]) cudaEventDestroy(c->ev_ring[k][i][j]);
    ring_release(c);
    bands_release(c);
    cudaFree(c->bands_flags); cudaFree(c->bands_overflow); cudaFree(c->bands_dst);
    if (c->h_bands_overflow) cudaFreeHost(c->h_bands_overflow);
    if (c->ev_bands) cudaEventDestroy(c->ev_bands);
    cudaFree(c->inbox); cudaFree(c->ring_flags); cudaFree(c->ring_hot_counts);
    if (c->ring_stream) { cudaStreamSynchronize(c->ring_stream); cudaStreamDestroy(c->ring_stream); }
    if (c->ev_ring_fwd) cudaEventDestroy(c->ev_ring_fwd);
    if (c->ev_ring_begin) cudaEventDestroy(c->ev_ring_begin);
    if (c->ev_state) cudaEventDestroy(c->ev_state);
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

int tb_resize_flow(tb_ctx *c, int32_t w, int32_t h) {
    TB_REQUIRE(c, c, "null context");
    TB_CUDA(c, cudaSetDevice(c->device));
    TB_CUDA(c, cudaStreamSynchronize(c->stream));
    return alloc_flow(c, w, h);
}

int tb_clear_flow(tb_ctx *c) {
    TB_REQUIRE(c, c, "null context");
    TB_CUDA(c, cudaSetDevice(c->device));
    TB_CUDA(c, cudaMemsetAsync(c->flow, 0, static_cast<size_t>(c->W) * c->H * sizeof(float4), c->stream));
    return TB_OK;
}

int tb_step(tb_ctx *c, float time, float dt) {
    TB_REQUIRE(c, c, "null context");
    TB_REQUIRE(c, c->have_state, "tb_set_state must be called before tb_step");
    TB_CUDA(c, cudaSetDevice(c->device));
    std::swap(c->buf[0], c->buf[1]);                       // utils.step(buffers), src/particles.js:128
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
    cudaEvent_t *ev = c->ev_ring[0][c->ev_count[0] % tb_ctx::kTimingSlots];
    auto is_pow2 = [](int v) { return v > 0 && (v & (v - 1)) == 0; };
    A.pow2_res = is_pow2(c->PW) && is_pow2(c->PH) && static_cast<long long>(c->PW) * c->PH <= (1LL << 40);
    A.inv_resx = 1.0f / static_cast<float>(c->PW);
    A.inv_resy = 1.0f / static_cast<float>(c->PH);
    A.inv_n = 1.0f / (static_cast<float>(c->PW) * static_cast<float>(c->PH));
    static const bool scalar_noise = std::getenv("TB_SCALAR_NOISE") != nullptr;    // A/B switch for profiling
    A.packed_noise = scalar_noise ? 0 : 1;
    A.pk.one = 1.0f; A.pk.neg_one = -1.0f; A.pk.neg_zero = -0.0f;
    A.wander = c->wander;
    const bool fuse = (c->fuse_count || c->fuse_partial) && c->n_prims > 0;
    A.row_pair = c->row_pair;
    A.prim_off = fuse ? c->prim_off : nullptr;
    A.n_pairs = c->n_pairs;
    c->count_valid = false;
    if (fuse) TB_CUDA(c, cudaMemsetAsync(c->prim_off, 0, (c->n_prims + 1) * sizeof(uint32_t), c->stream));
    const dim3 grid(blocks_for(c->PH, 256), static_cast<unsigned>(A.cols));
    if (A.use_noise && c->overlap && c->splat_since_step) {
        // The noise does not read the flow grid: evaluate it on the low-priority side stream, where it
        // runs under the previous step's sort + fold still queued on the main stream; the rest of the
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
    if (fuse && c->fuse_partial) {          // the pairs that could not ride: the state buffers are final on this stream
        const SplatArgs SA = splat_args(c);
        k_splat_count_odd<<<blocks_for(static_cast<long long>(SA.cols) * c->n_odd, 256), 256, 0, c->stream>>>(SA, c->odd_pairs, c->n_odd);
        if (int r = check_launch(c, "k_splat_count_odd")) return r;
    }
    if (fuse) {
        // the scan and the 4-byte total travel to the host now, so that the next tb_splat_flow finds the
        // fragment count waiting instead of stalling the GPU on a round trip
        if (int r = scan_counts(c)) return r;
        c->count_valid = true;
        c->count_wh[0] = c->W; c->count_wh[1] = c->H;
        c->count_vs[0] = S.viewSize[0]; c->count_vs[1] = S.viewSize[1];
    }
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
    if (int r = collect(c, time)) return r;
    return fold(c);
}

int tb_reset(tb_ctx *c) {
    TB_REQUIRE(c, c, "null context");
    c->count_valid = false;
    TB_CUDA(c, cudaSetDevice(c->device));
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
    float4 *out = spawn_out(c, target);
    k_spawn_init<<<blocks_for(c->n_local, 256), 256, 0, c->stream>>>(out, c->n_local);
    if (int r = check_launch(c, "k_spawn_init")) return r;
    return after_targets_write(c, target);
}

int tb_spawn_ball(tb_ctx *c, float radius, float speed, tb_target target) {
    TB_REQUIRE(c, c, "null context");
    TB_CUDA(c, cudaSetDevice(c->device));
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
    float4 *dst; int64_t n;
    if (int r = buffer_of(c, which, &dst, &n)) return r;
    TB_REQUIRE(c, n == n_floats, "tb_upload: size mismatch");
    if (which != TB_BUF_FLOW && which != TB_BUF_TARGETS) c->count_valid = false;
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

int tb_debug_segments(tb_ctx *c, uint32_t *host, int64_t n_words) {
    TB_REQUIRE(c, c && host, "null argument");
    TB_REQUIRE(c, n_words == 2LL * c->W * c->H, "tb_debug_segments: size mismatch");
    TB_CUDA(c, cudaSetDevice(c->device));
    TB_CUDA(c, cudaMemcpyAsync(host, c->seg, static_cast<size_t>(n_words) * sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
    TB_CUDA(c, cudaStreamSynchronize(c->stream));
    return TB_OK;
}

int tb_optical_flow(tb_ctx *c, const tb_optical_flow_params *params, const uint8_t *view_rgba8, const uint8_t *last_rgba8,
                    int32_t w, int32_t h) {
    TB_REQUIRE(c, c && params && view_rgba8 && last_rgba8, "null argument");
    TB_REQUIRE(c, w >= 1 && h >= 1, "gl-texture2d: Texture dimensions are out of bounds");
    TB_CUDA(c, cudaSetDevice(c->device));
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

int tb_device_ptr(tb_ctx *c, tb_buffer which, void **ptr, int64_t *n_floats) {
    TB_REQUIRE(c, c && ptr && n_floats, "null argument");
    float4 *p; int64_t n;
    if (int r = buffer_of(c, which, &p, &n)) return r;
    *ptr = p; *n_floats = n;
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
    TB_CUDA(c, cudaStreamSynchronize(c->side));
    TB_CUDA(c, cudaStreamSynchronize(c->stream));
    return TB_OK;
}

int tb_stats(tb_ctx *c, int64_t *kernel_launches, int64_t *last_fragments) {
    TB_REQUIRE(c, c, "null context");
    if (kernel_launches) *kernel_launches = c->launches;
    if (last_fragments) *last_fragments = c->last_frags;
    return TB_OK;
}

int tb_timing(tb_ctx *c, int reset, int64_t *n_integrate, float *integrate_ms, int64_t *n_splat, float *splat_ms,
              int64_t *n_noise, float *noise_ms) {
    TB_REQUIRE(c, c, "null context");
    TB_CUDA(c, cudaSetDevice(c->device));
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
