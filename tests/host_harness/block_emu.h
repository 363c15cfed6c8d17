/* TEST STUB -- one CUDA thread block on the CPU: one std::thread per CUDA thread, __syncthreads and the warp-synchronous
 * intrinsics as barriers, __shared__ as static storage (one block runs at a time).  Good for single-CTA kernels whose
 * control flow is uniform around the barriers (the plan kernels of the flow splat are), which tests/test_plan_host.py
 * checks against a numpy restatement of what they have to produce.  Never used by the product. */
#ifndef TB_TEST_BLOCK_EMU_H
#define TB_TEST_BLOCK_EMU_H
#include <barrier>
#include <cstdint>
#include <cstring>
#include <functional>
#include <memory>
#include <thread>
#include <vector>

struct HostIdx { unsigned x, y, z; };
static thread_local HostIdx tb_host_blockIdx, tb_host_threadIdx;
static HostIdx tb_host_blockDim, tb_host_gridDim;
#define blockIdx tb_host_blockIdx
#define threadIdx tb_host_threadIdx
#define blockDim tb_host_blockDim
#define gridDim tb_host_gridDim
#define __shared__ static
#define __global__
#define __launch_bounds__(...)
#define __restrict__

struct WarpEmu {
    std::barrier<> bar{32};
    uint64_t slot[32];
};
static thread_local WarpEmu *tb_warp = nullptr;
static thread_local int tb_lane = 0;
static std::barrier<> *tb_block_bar = nullptr;

static inline void __syncthreads() { tb_block_bar->arrive_and_wait(); }
static inline void __syncwarp() { tb_warp->bar.arrive_and_wait(); }
static inline uint64_t tb_exchange(uint64_t v, int src) {
    tb_warp->slot[tb_lane] = v;
    tb_warp->bar.arrive_and_wait();
    const uint64_t r = tb_warp->slot[src & 31];
    tb_warp->bar.arrive_and_wait();
    return r;
}
template <class T> static inline T __shfl_sync(unsigned, T v, int src) {
    static_assert(sizeof(T) == 4 || sizeof(T) == 8, "32- and 64-bit shuffles");
    uint64_t u = 0; std::memcpy(&u, &v, sizeof(T));
    u = tb_exchange(u, src);
    T r; std::memcpy(&r, &u, sizeof(T));
    return r;
}
template <class T> static inline T __shfl_up_sync(unsigned m, T v, int delta) { return __shfl_sync(m, v, tb_lane >= delta ? tb_lane - delta : tb_lane); }
static inline unsigned tb_gather(uint64_t v, uint64_t *all) {       /* every lane's value, to every lane */
    tb_warp->slot[tb_lane] = v;
    tb_warp->bar.arrive_and_wait();
    for (int i = 0; i < 32; ++i) all[i] = tb_warp->slot[i];
    tb_warp->bar.arrive_and_wait();
    return 0;
}
static inline unsigned __ballot_sync(unsigned, int pred) {
    uint64_t all[32]; tb_gather(pred ? 1u : 0u, all);
    unsigned m = 0;
    for (int i = 0; i < 32; ++i) m |= static_cast<unsigned>(all[i]) << i;
    return m;
}
static inline int __any_sync(unsigned m, int pred) { return __ballot_sync(m, pred) != 0u; }
static inline unsigned __match_any_sync(unsigned, uint32_t v) {
    uint64_t all[32]; tb_gather(v, all);
    unsigned m = 0;
    for (int i = 0; i < 32; ++i) m |= (all[i] == v ? 1u : 0u) << i;
    return m;
}
static inline unsigned __reduce_max_sync(unsigned, uint32_t v) {
    uint64_t all[32]; tb_gather(v, all);
    uint64_t m = 0;
    for (int i = 0; i < 32; ++i) m = all[i] > m ? all[i] : m;
    return static_cast<unsigned>(m);
}
template <class T> static inline T __ldg(const T *p) { return *p; }
template <class T> static inline T __ldcs(const T *p) { return *p; }
static inline void __threadfence_block() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
static inline uint32_t atomicMax(uint32_t *p, uint32_t v) {
    uint32_t old = __atomic_load_n(p, __ATOMIC_SEQ_CST);
    while (old < v && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {}
    return old;
}
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline int __ffs(unsigned v) { return __builtin_ffs((int)v); }
static inline int __clz(unsigned v) { return v ? __builtin_clz(v) : 32; }
static inline uint32_t atomicAdd(uint32_t *p, uint32_t v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
static inline uint32_t atomicOr(uint32_t *p, uint32_t v) { return __atomic_fetch_or(p, v, __ATOMIC_SEQ_CST); }

/* a kernel without barriers: its threads one after the other, no OS threads (grid gx x gy blocks of n_threads) */
static inline void tb_run_serial(unsigned gx, unsigned gy, unsigned n_threads, const std::function<void()> &body) {
    tb_host_blockDim = {n_threads, 1, 1};
    tb_host_gridDim = {gx, gy, 1};
    for (unsigned by = 0; by < gy; ++by)
        for (unsigned bx = 0; bx < gx; ++bx)
            for (unsigned t = 0; t < n_threads; ++t) {
                tb_host_blockIdx = {bx, by, 0};
                tb_host_threadIdx = {t, 0, 0};
                body();
            }
}

/* run one block of n_threads (a multiple of 32): body() is called by every thread with threadIdx.x set */
static inline void tb_run_block(unsigned n_threads, const std::function<void()> &body) {
    const unsigned n_warps = n_threads / 32;
    std::vector<std::unique_ptr<WarpEmu>> warps;
    for (unsigned w = 0; w < n_warps; ++w) warps.emplace_back(new WarpEmu);
    std::barrier<> block_bar{static_cast<std::ptrdiff_t>(n_threads)};
    tb_block_bar = &block_bar;
    tb_host_blockDim = {n_threads, 1, 1};
    tb_host_gridDim = {1, 1, 1};
    std::vector<std::thread> threads;
    for (unsigned t = 0; t < n_threads; ++t)
        threads.emplace_back([&, t] {
            tb_warp = warps[t / 32].get();
            tb_lane = static_cast<int>(t % 32);
            tb_host_blockIdx = {0, 0, 0};
            tb_host_threadIdx = {t, 0, 0};
            body();
            tb_warp->bar.arrive_and_drop();              /* a thread that returns early must not stall the others */
            block_bar.arrive_and_drop();
        });
    for (auto &t : threads) t.join();
    tb_block_bar = nullptr;
}
#endif
