/* TEST STUB -- a 32-lane warp on the CPU: one std::thread per lane, the warp-synchronous intrinsics as barriers.
 * Good for kernels whose control flow is uniform around __shfl / __any / __syncwarp (the fold kernels are), which
 * tests/test_fold_host.py runs against a plain sequential fold.  Never used by the product. */
#ifndef TB_TEST_WARP_EMU_H
#define TB_TEST_WARP_EMU_H
#include <barrier>
#include <cstdint>
#include <cstring>
#include <functional>
#include <thread>
#include <vector>

struct HostIdx { unsigned x, y, z; };
static thread_local HostIdx tb_host_blockIdx, tb_host_threadIdx;
static HostIdx tb_host_blockDim, tb_host_gridDim;
#define blockIdx tb_host_blockIdx
#define threadIdx tb_host_threadIdx
#define blockDim tb_host_blockDim
#define gridDim tb_host_gridDim

struct WarpEmu {
    std::barrier<> bar{32};
    uint32_t slot[32];
};
static WarpEmu *tb_warp = nullptr;
static thread_local int tb_lane = 0;

static inline uint32_t tb_exchange(uint32_t v, int src) {
    tb_warp->slot[tb_lane] = v;
    tb_warp->bar.arrive_and_wait();
    const uint32_t r = tb_warp->slot[src & 31];
    tb_warp->bar.arrive_and_wait();
    return r;
}
template <class T> static inline T __shfl_sync(unsigned, T v, int src) {
    static_assert(sizeof(T) == 4, "32-bit shuffles only");
    uint32_t u; std::memcpy(&u, &v, 4);
    u = tb_exchange(u, src);
    T r; std::memcpy(&r, &u, 4);
    return r;
}
template <class T> static inline T __shfl_xor_sync(unsigned m, T v, int lane_mask) { return __shfl_sync(m, v, tb_lane ^ lane_mask); }
static inline int __any_sync(unsigned, int pred) {
    tb_warp->slot[tb_lane] = pred ? 1u : 0u;
    tb_warp->bar.arrive_and_wait();
    uint32_t any = 0;
    for (int i = 0; i < 32; ++i) any |= tb_warp->slot[i];
    tb_warp->bar.arrive_and_wait();
    return any != 0;
}
static inline void __syncwarp() { tb_warp->bar.arrive_and_wait(); }
template <class T> static inline T __shfl_up_sync(unsigned m, T v, int delta) { return __shfl_sync(m, v, tb_lane >= delta ? tb_lane - delta : tb_lane); }
static inline unsigned __ballot_sync(unsigned, int pred) {
    tb_warp->slot[tb_lane] = pred ? 1u : 0u;
    tb_warp->bar.arrive_and_wait();
    uint32_t m = 0;
    for (int i = 0; i < 32; ++i) m |= tb_warp->slot[i] << i;
    tb_warp->bar.arrive_and_wait();
    return m;
}
static inline unsigned __match_any_sync(unsigned, uint32_t v) {
    tb_warp->slot[tb_lane] = v;
    tb_warp->bar.arrive_and_wait();
    uint32_t m = 0;
    for (int i = 0; i < 32; ++i) m |= (tb_warp->slot[i] == v ? 1u : 0u) << i;
    tb_warp->bar.arrive_and_wait();
    return m;
}
static inline unsigned __reduce_max_sync(unsigned, uint32_t v) {
    tb_warp->slot[tb_lane] = v;
    tb_warp->bar.arrive_and_wait();
    uint32_t m = 0;
    for (int i = 0; i < 32; ++i) m = tb_warp->slot[i] > m ? tb_warp->slot[i] : m;
    tb_warp->bar.arrive_and_wait();
    return m;
}
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline int __ffs(unsigned v) { return __builtin_ffs((int)v); }
static inline int __clz(unsigned v) { return v ? __builtin_clz(v) : 32; }
static inline uint32_t atomicAdd(uint32_t *p, uint32_t v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
static inline uint32_t atomicOr(uint32_t *p, uint32_t v) { return __atomic_fetch_or(p, v, __ATOMIC_SEQ_CST); }

/* run one warp: body() is called by 32 lane threads with threadIdx.x = first_thread + lane */
static inline void tb_run_warp(unsigned block, unsigned first_thread, const std::function<void()> &body) {
    WarpEmu warp;
    tb_warp = &warp;
    std::vector<std::thread> lanes;
    for (int l = 0; l < 32; ++l)
        lanes.emplace_back([&, l] {
            tb_lane = l;
            tb_host_blockIdx = {block, 0, 0};
            tb_host_threadIdx = {first_thread + (unsigned)l, 0, 0};
            body();
            warp.bar.arrive_and_drop();                  /* a lane that returns early must not stall the others */
        });
    for (auto &t : lanes) t.join();
    tb_warp = nullptr;
}
#endif
