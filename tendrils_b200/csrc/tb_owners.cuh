// tb_owners.cuh -- the flow splat of a column-sharded run (one process per GPU, one node), sm_100a.
//
// Particles are sharded by contiguous column blocks = contiguous ranges of the draw order
// (src/particles.js:182-186), the flow grid is replicated.  The ordered blend cannot be summed over ranks;
// instead the BINS of tb_splat.cuh are owned round-robin (bin b belongs to rank b % P), and per draw:
//
//   every rank   counts its fragments per bin (k_splat_hist, k_splat_rows) and stores the totals into every
//                rank's table over NVLink (k_owners_share)                                      | barrier
//   every rank   runs the same plan on the same table (k_owners_plan): per bin the ranks' fragments lie side by
//                side in rank order = column order = draw order in the OWNER's bin array
//   every rank   rasterises (k_splat_scatter) -- the fragments go straight to their place in the owner's array,
//                posted 16-byte stores over NVLink for (P-1)/P of them                           | barrier
//   every rank   folds the bins it owns (k_splat_fold) and stores the finished texels into every rank's grid | barrier
//
// The result equals the single-GPU draw bit for bit.  Nothing but these kernels touches the data path: the
// process group of the host layer only carries the IPC handles once.
#pragma once

#include "tb_splat.cuh"

namespace tb {

constexpr int kOwnerPhases = 4;

struct OwnerPeers {
    uint32_t *totals[kMaxBandRanks];     // every rank's table [n][kMaxBins]: row r = rank r's fragments per bin
    uint32_t *flags[kMaxBandRanks];      // every rank's barrier flags [kOwnerPhases][kMaxBandRanks], then [kMaxBandRanks] prune flags
    uint32_t *last[kMaxBandRanks];       // every rank's table [n][W*H]: row r = rank r's "last opaque primitive" per texel
    int n, me;
};

// opaque pruning across the ranks: my table of last opaque primitives (and my "do not prune" flag) into row `me` of every rank's
__global__ void __launch_bounds__(256) k_owners_push_last(const uint32_t *__restrict__ last, const uint32_t *__restrict__ flags, int G,
                                                           const OwnerPeers P) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i == 0) {
        const uint32_t f = *flags;
        for (int r = 0; r < P.n; ++r) P.flags[r][kOwnerPhases * kMaxBandRanks + P.me] = f;
    }
    if (i >= G) return;
    const uint32_t v = last[i];
    for (int r = 0; r < P.n; ++r) P.last[r][static_cast<size_t>(P.me) * G + i] = v;
}
// ... and, behind a barrier, the maximum over the ranks (primitive indices are global: rank order = draw order)
__global__ void __launch_bounds__(256) k_owners_last_max(const uint32_t *__restrict__ all, const uint32_t *__restrict__ my_flags, int n, int G,
                                                          uint32_t *__restrict__ last, uint32_t *flags) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i == 0) {
        uint32_t f = 0u;
        for (int r = 0; r < n; ++r) f |= my_flags[kOwnerPhases * kMaxBandRanks + r];
        *flags = f;
    }
    if (i >= G) return;
    uint32_t m = 0u;
    for (int r = 0; r < n; ++r) { const uint32_t v = all[static_cast<size_t>(r) * G + i]; m = v > m ? v : m; }
    last[i] = m;
}

// my fragments per bin into row `me` of every rank's table
__global__ void __launch_bounds__(256) k_owners_share(const uint32_t *__restrict__ bin_total, const uint32_t *__restrict__ n_bins, const OwnerPeers P) {
    const int b = blockIdx.x * 256 + threadIdx.x;
    if (b >= static_cast<int>(*n_bins)) return;
    const uint32_t v = bin_total[b];
    for (int r = 0; r < P.n; ++r) P.totals[r][static_cast<size_t>(P.me) * kMaxBins + b] = v;
}

// all-rank barrier over peer memory: thread j tells rank j "I reached `epoch`" and waits for rank j to say so
__global__ void k_owners_barrier(const OwnerPeers P, uint32_t *my_flags, int phase, uint32_t epoch) {
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

constexpr int kOwnerPlanPer = 12;                         // bins per thread and owner class (kMaxBins / P / (1024 / P), rounded up)

struct OwnerPlanArgs {
    int T, lS;
    BinMap bm;
    const uint32_t *__restrict__ bin_info;     // this draw's bins
    const uint32_t *__restrict__ totals;       // [n][kMaxBins] every rank's fragments per bin (identical on every rank)
    int n, me;
    uint32_t caps[kMaxBandRanks];              // capacity of every rank's bin array
    uint32_t *__restrict__ bin_sum;            // [kMaxBins] scratch: fragments per bin over all ranks
    uint32_t *__restrict__ scat_off;           // [kMaxBins] where MY fragments of bin b start in the array of rank b % n
    uint32_t *__restrict__ own_begin;          // [kMaxBins] for the bins I own: where the bin starts in my array ...
    uint32_t *__restrict__ own_count;          // [kMaxBins] ... and how many fragments it holds
    uint32_t *__restrict__ items;              // [16 kMaxBins] fold work items of the bins I own, longest first
    uint32_t split_at, share_at;
    SegPlan seg;
    uint32_t *tickets;
    uint32_t *map_next, *bin_info_next, *n_bins_next;
    PlanOut *out;
};

__global__ void __launch_bounds__(kPlanThreads) k_owners_plan(const OwnerPlanArgs A) {
    __shared__ unsigned long long s_cls[kPlanThreads];       // per (owner class, thread of the class): fragments of its bins
    __shared__ unsigned long long s_owner_total[kMaxBandRanks];
    __shared__ unsigned long long s_warp[32];
    __shared__ unsigned long long s_total;
    __shared__ uint32_t s_bucket[33];
    __shared__ uint32_t s_seg[3];
    const int B = static_cast<int>(*A.bm.n_bins);
    const int P = A.n;
    if (threadIdx.x < 3) s_seg[threadIdx.x] = 0u;
    const int C = kPlanThreads / P;                          // threads per owner class
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int o = tid % P, i = tid / P;                      // this thread: owner class o, i-th thread of the class
    const bool in_class = i < C;
    const int per_class = (kMaxBins + P - 1) / P;            // bins of a class: b = o + P*m, m < per_class
    const int K = (per_class + C - 1) / C;                   // <= kOwnerPlanPer
    uint32_t S[kOwnerPlanPer], pre[kOwnerPlanPer];
    unsigned long long mine = 0ull, emitted = 0ull;
#pragma unroll
    for (int k = 0; k < kOwnerPlanPer; ++k) {
        S[k] = 0u; pre[k] = 0u;
        const int b = o + P * (i * K + k);
        if (in_class && k < K && b < B) {
            unsigned long long all = 0ull;
            for (int r = 0; r < P; ++r) {
                const uint32_t v = A.totals[static_cast<size_t>(r) * kMaxBins + b];
                if (r < A.me) pre[k] += v;
                if (r == A.me) emitted += v;
                all += v;
            }
            S[k] = all > 0xffffffffull ? 0xffffffffu : static_cast<uint32_t>(all);
            mine += all;
            A.bin_sum[b] = S[k];
        }
    }
    if (tid < 33) s_bucket[tid] = 0u;
    if (in_class) s_cls[o * C + i] = mine;
    __syncthreads();
    // per class: exclusive scan over its threads (warp w scans class w)
    for (int cls = warp; cls < P; cls += kPlanThreads / 32) {
        const int per_lane = (C + 31) / 32;
        unsigned long long sum = 0ull;
        for (int k = 0; k < per_lane; ++k) {
            const int idx = lane * per_lane + k;
            if (idx < C) sum += s_cls[cls * C + idx];
        }
        unsigned long long inc = sum;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const unsigned long long u = __shfl_up_sync(0xffffffffu, inc, d);
            if (lane >= d) inc += u;
        }
        unsigned long long run = inc - sum;
        for (int k = 0; k < per_lane; ++k) {
            const int idx = lane * per_lane + k;
            if (idx < C) { const unsigned long long v = s_cls[cls * C + idx]; s_cls[cls * C + idx] = run; run += v; }
        }
        if (lane == 31) s_owner_total[cls] = inc;
    }
    __syncthreads();
    // the same verdict on every rank: every owner's fragments fit its array, or nobody scatters
    bool ok = true;
    for (int r = 0; r < P; ++r) ok = ok && s_owner_total[r] <= static_cast<unsigned long long>(A.caps[r]);
    block_excl_scan64(emitted, s_warp, &s_total);
    const unsigned long long total_emitted = s_total;          // this rank's fragments
    if (tid == 0) {
        A.out->total = total_emitted;
        A.out->needed = s_owner_total[A.me];
        A.out->overflow = ok ? 0u : 1u;
        A.tickets[0] = 0u; A.tickets[1] = 0u; A.tickets[2] = 0u;
    }
    unsigned long long run = in_class ? s_cls[o * C + i] : 0ull;
    uint32_t rank[kOwnerPlanPer], lparts[kOwnerPlanPer];
    int bucket[kOwnerPlanPer];
#pragma unroll
    for (int k = 0; k < kOwnerPlanPer; ++k) {
        rank[k] = 0u; bucket[k] = 0; lparts[k] = 0u;
        const int b = o + P * (i * K + k);
        if (in_class && k < K && b < B) {
            A.scat_off[b] = ok ? static_cast<uint32_t>(run) + pre[k] : 0u;
            if (o == A.me) {
                A.own_begin[b] = ok ? static_cast<uint32_t>(run) : 0u;
                A.own_count[b] = ok ? S[k] : 0u;
                bucket[k] = __clz(S[k] | 1u);
                if (ok && S[k]) {
                    const uint32_t R = (1u << A.lS) >> (A.bin_info[b] >> 24);
                    const uint32_t lseg = plan_segments(static_cast<uint32_t>(b), S[k], R, 1u << A.lS, s_owner_total[A.me], A.seg, s_seg);
                    lparts[k] = lseg ? (lseg | 0x80u) : fold_lparts(S[k], R, A.share_at);
                    rank[k] = atomicAdd(&s_bucket[bucket[k]], 1u << (lparts[k] & 0x7fu));
                }
            }
            run += S[k];
        }
    }
    __syncthreads();
    if (tid == 0) {
        uint32_t r = 0;
        for (int b = 0; b < 32; ++b) { const uint32_t c = s_bucket[b]; s_bucket[b] = r; r += c; }
        A.out->n_items = r;
        A.tickets[3] = r;
        A.tickets[5] = s_seg[0];
        A.tickets[6] = 0u;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < kOwnerPlanPer; ++k) {
        const int b = o + P * (i * K + k);
        if (in_class && k < K && b < B && o == A.me && ok && S[k])
            for (uint32_t part = 0; part < (1u << (lparts[k] & 0x7fu)); ++part)
                A.items[s_bucket[bucket[k]] + rank[k] + part] = static_cast<uint32_t>(b) | (part << 16) | (lparts[k] << 24);
    }
    // the next draw's split map, from the fragments per strip over ALL ranks: identical on every rank
    unsigned long long all_ranks = 0ull;
    for (int r = 0; r < P; ++r) all_ranks += s_owner_total[r];
    plan_next_map(A.T, A.lS, A.bm.map, A.bin_sum, all_ranks, A.split_at, A.map_next, A.bin_info_next, A.n_bins_next, s_warp, &s_total);
}

}  // namespace tb
