"""The single-CTA plan kernels of the flow splat (k_splat_plan of tendrils_b200/csrc/tb_splat.cuh, k_owners_plan of
tb_owners.cuh) cut out of the product source unchanged and run on a CPU emulation of a thread block -- one thread per CUDA
thread, __syncthreads / shuffles as barriers, __shared__ as static storage (tests/host_harness/block_emu.h) -- against a
numpy restatement of what they have to produce: bin offsets, the capacity verdict, a fold work list that covers every
non-empty bin exactly once (longest first), the next draw's split map, and, for a sharded run, the SAME verdict and map on
every rank with the ranks' fragments side by side, in rank order, inside the owner's bin.  Nothing here is used by the product."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MAX_BINS, MAX_RANKS = 11264, 16
_up = C.POINTER(C.c_uint32)

HARNESS = r'''
#include <algorithm>
#include <cstddef>
#include <cstdint>
#include <cuda_runtime.h>
#undef __shared__
#undef __global__
#undef __launch_bounds__
#include "cuda_intrinsics_shim.h"
#include "block_emu.h"
namespace tb {
using std::min; using std::max;
constexpr int kMaxStrips = 8192;
constexpr int kMaxBins = 11264;
constexpr int kMaxBandRanks = 16;
struct BinMap { const uint32_t *map; const uint32_t *n_bins; int lS; };
struct PlanOut { unsigned long long total, needed; uint32_t overflow, n_items; };
%(plan)s
%(owners)s
}
using namespace tb;

struct Out { unsigned long long total, needed; uint32_t overflow, n_items; };

extern "C" void ph_plan(int T, int lS, const uint32_t *map, const uint32_t *n_bins, const uint32_t *bin_info, const uint32_t *bin_total,
                        uint32_t cap, uint32_t split_at, uint32_t share_at, uint32_t seg_at, uint32_t seg_len, uint32_t seg_scaled,
                        uint32_t seg_cap, uint32_t *bin_off, uint32_t *items, uint32_t *tickets, uint32_t *map_next,
                        uint32_t *bin_info_next, uint32_t *n_bins_next, uint32_t *seg_desc, uint32_t *seg_of_bin, Out *out) {
    PlanArgs A{};
    A.T = T; A.lS = lS;
    A.bm = BinMap{map, n_bins, lS};
    A.bin_info = bin_info; A.bin_total = bin_total; A.bin_off = bin_off; A.items = items;
    A.cap = cap; A.split_at = split_at; A.share_at = share_at;
    A.seg = SegPlan{seg_at, seg_scaled, seg_len, seg_cap, reinterpret_cast<uint4 *>(seg_desc), seg_of_bin};
    A.too_many = tickets + 4; A.tickets = tickets;
    A.map_next = map_next; A.bin_info_next = bin_info_next; A.n_bins_next = n_bins_next;
    PlanOut po{};
    A.out = &po;
    tb_run_block(kPlanThreads, [&] { k_splat_plan(A); });
    out->total = po.total; out->needed = po.needed; out->overflow = po.overflow; out->n_items = po.n_items;
}

extern "C" void ph_owners_plan(int T, int lS, const uint32_t *map, const uint32_t *n_bins, const uint32_t *bin_info, const uint32_t *totals,
                               int n, int me, const uint32_t *caps, uint32_t split_at, uint32_t share_at, uint32_t seg_at, uint32_t seg_len,
                               uint32_t seg_cap, uint32_t *scratch /* 4 kMaxBins */, uint32_t *items, uint32_t *tickets, uint32_t *map_next,
                               uint32_t *bin_info_next, uint32_t *n_bins_next, uint32_t *seg_desc, uint32_t *seg_of_bin, Out *out) {
    OwnerPlanArgs A{};
    A.T = T; A.lS = lS;
    A.bm = BinMap{map, n_bins, lS};
    A.bin_info = bin_info; A.totals = totals; A.n = n; A.me = me;
    for (int r = 0; r < n; ++r) A.caps[r] = caps[r];
    A.bin_sum = scratch; A.scat_off = scratch + kMaxBins; A.own_begin = scratch + 2 * kMaxBins; A.own_count = scratch + 3 * kMaxBins;
    A.items = items; A.split_at = split_at; A.share_at = share_at;
    A.seg = SegPlan{seg_at, 0u, seg_len, seg_cap, reinterpret_cast<uint4 *>(seg_desc), seg_of_bin};
    A.tickets = tickets;
    A.map_next = map_next; A.bin_info_next = bin_info_next; A.n_bins_next = n_bins_next;
    PlanOut po{};
    A.out = &po;
    tb_run_block(kPlanThreads, [&] { k_owners_plan(A); });
    out->total = po.total; out->needed = po.needed; out->overflow = po.overflow; out->n_items = po.n_items;
}
'''


class Out(C.Structure):
    _fields_ = [("total", C.c_ulonglong), ("needed", C.c_ulonglong), ("overflow", C.c_uint32), ("n_items", C.c_uint32)]


@pytest.fixture(scope="module")
def ph(tmp_path_factory):
    return build_harness(tmp_path_factory.mktemp("ph"))


def build_harness(d):
    csrc = os.path.join(ROOT, "tendrils_b200", "csrc")
    src = open(os.path.join(csrc, "tb_splat.cuh")).read()
    plan = src[src.index("constexpr int kPlanThreads = 1024;"):src.index("// the identity map (one bin per strip)")]
    osrc = open(os.path.join(csrc, "tb_owners.cuh")).read()
    owners = osrc[osrc.index("constexpr int kOwnerPlanPer"):osrc.rindex("}  // namespace tb")]
    assert "k_splat_plan" in plan and "k_owners_plan" in owners and "asm" not in plan + owners
    strip = lambda s: s.replace("__host__", "").replace("__device__", "")
    cpp = d / "plan_host.cpp"
    cpp.write_text(HARNESS % {"plan": strip(plan), "owners": strip(owners)})
    out = d / "libplan_host.so"
    subprocess.run(["g++", "-O1", "-std=c++20", "-fPIC", "-shared", "-pthread", "-Wno-unknown-pragmas", "-Wno-unused-function", "-Wno-attributes",
                    "-I/usr/local/cuda/include", "-I", os.path.join(ROOT, "tests", "host_harness"), "-o", str(out), str(cpp)], check=True)
    L = C.CDLL(str(out))
    L.ph_plan.restype = None
    L.ph_owners_plan.restype = None
    return L


def u32(a):
    return np.ascontiguousarray(a, np.uint32)


def ptr(a):
    return a.ctypes.data_as(_up)


def random_map(rng, T, lS, split_frac):
    """a split map as a previous draw could have left it: some strips split 2^ls ways"""
    ls = np.where(rng.random(T) < split_frac, rng.integers(1, min(lS, 8) + 1, T), 0)
    while (1 << ls).sum() > MAX_BINS:
        ls[np.argmax(ls)] = 0
    first = np.concatenate([[0], np.cumsum(1 << ls)[:-1]])
    n_bins = int((1 << ls).sum())
    info = np.zeros(MAX_BINS, np.uint32)
    for t in range(T):
        for sub in range(1 << ls[t]):
            info[first[t] + sub] = t | (sub << 16) | (int(ls[t]) << 24)
    return u32(first | (ls << 24)), u32([n_bins]), info, ls, first, n_bins


def want(frags, at, lS):
    ls = 0
    while ls < min(lS, 8) and (int(frags) >> ls) > at:          # at most 256 bins per strip: `sub` has 8 bits in bin_info
        ls += 1
    return ls


def expected_next_map(strip_frags, total, split_at, lS):
    """as many bins (a power of two) as it takes to bring a strip's bins under 1/4096 of the draw (between 256 and
    split_at fragments), the target raised by a quarter until the map fits kMaxBins."""
    at = min(max(total // 4096, 256), split_at)
    while sum(1 << want(f, at, lS) for f in strip_frags) > MAX_BINS:
        at += (at >> 2) + 1
    return np.array([want(f, at, lS) for f in strip_frags])


def check_next_map(strip_frags, total, split_at, lS, map_next, info_next, n_bins_next):
    ls = expected_next_map(strip_frags, total, split_at, lS)
    first = np.concatenate([[0], np.cumsum(1 << ls)[:-1]])
    assert n_bins_next == (1 << ls).sum() <= MAX_BINS
    assert np.array_equal(map_next & 0xffffff, first) and np.array_equal(map_next >> 24, ls)
    for t in np.flatnonzero(ls > 0)[:50].tolist() + [0, len(ls) - 1]:
        for sub in range(1 << ls[t]):
            assert info_next[first[t] + sub] == t | (sub << 16) | (int(ls[t]) << 24)


def lparts_shared(n, R, share_at):
    lp = 0
    while lp < 3 and (R >> (lp + 1)) >= 1 and (n >> lp) > share_at:
        lp += 1
    return lp


def check_items(items, n_items, counts, R_of, owned, expect_lp):
    """every non-empty owned bin exactly once per part, parts 0 .. 2^lp - 1, longest bins (by power of two) first"""
    it = items[:n_items]
    bins, part, lp, seg = it & 0xffff, (it >> 16) & 0xff, (it >> 24) & 0x7f, it >> 31
    seen = {}
    for b, p, l, s in zip(bins.tolist(), part.tolist(), lp.tolist(), seg.tolist()):
        seen.setdefault(b, []).append((p, l, s))
    want_bins = [b for b in owned if counts[b] > 0]
    assert sorted(seen) == sorted(want_bins)
    for b, parts in seen.items():
        l, s = parts[0][1], parts[0][2]
        assert sorted(p for p, _, _ in parts) == list(range(1 << l)) and all(q[1:] == (l, s) for q in parts)
        assert (l, s) == expect_lp(b, int(counts[b]), R_of(b)), (b, int(counts[b]), R_of(b), l, s)
    clz = np.array([32 - int(counts[b]).bit_length() for b in bins.tolist()])
    assert np.all(np.diff(clz) >= 0), "work list not longest first"


@pytest.mark.parametrize("T,lS,split_frac,scale,cap,seg_at,seed", [
    (8192, 7, 0.02, 3000, 1 << 30, 0, 1),            # cfg3-like: 8192 strips of 128 texels, a few split
    (8192, 7, 0.05, 40000, 1 << 30, 0, 2),           # crowded: the next map has to raise its target to fit the bins
    (1024, 9, 0.3, 800, 1 << 30, 0, 3),              # strips of 512 texels (2048^2 grids), many split
    (64, 7, 0.5, 50, 1 << 30, 0, 4),                 # a small draw: splits early
    (4096, 7, 0.1, 5000, 100000, 0, 5),              # does not fit: overflow, nothing may be scattered
    (8192, 7, 0.05, 20000, 1 << 30, 16384, 6),       # long split bins are folded in segments
    (300, 8, 0.2, 0, 1 << 30, 0, 7),                 # an empty draw
    (64, 9, 0.0, 300000, 1 << 30, 0, 8),             # strips of 512 texels with crowds that want 512 bins each: capped at 256
])
def test_single_gpu_plan(ph, T, lS, split_frac, scale, cap, seg_at, seed):
    rng = np.random.default_rng(seed)
    map_, n_bins_a, info, ls, first, n_bins = random_map(rng, T, lS, split_frac)
    counts = np.zeros(MAX_BINS, np.uint32)
    if scale:
        counts[:n_bins] = (rng.exponential(scale, n_bins) * (rng.random(n_bins) < 0.8)).astype(np.uint32)
        counts[rng.integers(0, n_bins, 5)] = rng.integers(8 * scale, 40 * scale, 5)          # a few crowds
    total = int(counts.astype(np.int64).sum())
    bin_off, items, tickets = np.zeros(MAX_BINS + 1, np.uint32), np.zeros(16 * MAX_BINS, np.uint32), u32([7, 7, 7, 7, 0, 7, 7, 0])
    map_next, info_next, n_next = np.zeros(T, np.uint32), np.zeros(MAX_BINS, np.uint32), np.zeros(1, np.uint32)
    seg_desc, seg_of = np.zeros(4 * MAX_BINS, np.uint32), np.zeros(MAX_BINS, np.uint32)
    out = Out()
    split_at, share_at, seg_len, seg_cap = 8192, 12288, 8192, 1 << 20
    ph.ph_plan(T, lS, ptr(map_), ptr(n_bins_a), ptr(info), ptr(counts), cap, split_at, share_at, seg_at, seg_len, 0, seg_cap,
               ptr(bin_off), ptr(items), ptr(tickets), ptr(map_next), ptr(info_next), ptr(n_next), ptr(seg_desc), ptr(seg_of), C.byref(out))
    ok = total <= cap
    assert (out.total, out.needed, out.overflow) == (total, total, 0 if ok else 1)
    assert tickets[:3].tolist() == [0, 0, 0] and tickets[3] == out.n_items and tickets[6] == 0
    off = np.concatenate([[0], np.cumsum(counts[:n_bins].astype(np.int64))])
    assert np.array_equal(bin_off[:n_bins + 1], off if ok else np.zeros(n_bins + 1))
    S = 1 << lS
    eff_share = min(max(total // 2048, 512), share_at)
    seg_len_eff = max(total // (2 * MAX_BINS), seg_len)

    def expect_lp(b, n, R):
        if seg_at and 2 * R <= S and n > seg_at:
            lp = 1
            while lp < 4 and (n >> lp) > seg_len_eff:
                lp += 1
            return lp, 1
        return lparts_shared(n, R, eff_share), 0

    if ok:
        check_items(items, out.n_items, counts, lambda b: S >> (int(info[b]) >> 24), range(n_bins), expect_lp)
        n_seg = int(tickets[5])
        segged = sorted(b for b in range(n_bins) if counts[b] and expect_lp(b, int(counts[b]), S >> (int(info[b]) >> 24))[1])
        desc = seg_desc[:4 * n_seg].reshape(n_seg, 4)
        assert sorted(desc[:, 0].tolist()) == segged and (n_seg > 0) == (seg_at > 0)
        spans = []
        for i, (b, lp, slot, cnt) in enumerate(desc.tolist()):
            assert seg_of[b] == i and lp == expect_lp(b, int(counts[b]), S >> (int(info[b]) >> 24))[0]
            spans.append((slot, slot + ((S >> (int(info[b]) >> 24)) << lp)))
        spans.sort()
        assert all(a[1] <= b[0] for a, b in zip(spans, spans[1:])) and (not spans or spans[-1][1] <= seg_cap)     # disjoint result slots
    else:
        assert out.n_items == 0
    strip_frags = [int(counts[first[t]:first[t] + (1 << ls[t])].astype(np.int64).sum()) for t in range(T)]
    check_next_map(strip_frags, total, split_at, lS, map_next, info_next, int(n_next[0]))


@pytest.mark.parametrize("P,T,lS,split_frac,scale,cap_scale,seg_at,seed", [
    (2, 8192, 7, 0.05, 4000, 4.0, 0, 11),
    (8, 8192, 7, 0.05, 3000, 4.0, 0, 12),            # the weak-scaling bench: 8 ranks on one grid
    (3, 500, 7, 0.3, 700, 4.0, 0, 13),               # a rank count that does not divide anything
    (4, 2048, 9, 0.2, 2000, 0.5, 0, 14),             # one owner's array is too small: every rank must refuse
    (8, 8192, 7, 0.05, 6000, 4.0, 16384, 15),        # owners fold their long bins in segments
    (16, 1024, 7, 0.5, 300, 4.0, 0, 16),
])
def test_sharded_plan_is_the_same_on_every_rank(ph, P, T, lS, split_frac, scale, cap_scale, seg_at, seed):
    rng = np.random.default_rng(seed)
    map_, n_bins_a, info, ls, first, n_bins = random_map(rng, T, lS, split_frac)
    totals = np.zeros((P, MAX_BINS), np.uint32)
    totals[:, :n_bins] = (rng.exponential(scale, (P, n_bins)) * (rng.random((P, n_bins)) < 0.7)).astype(np.uint32)
    totals[:, rng.integers(0, n_bins, 4)] = rng.integers(4 * scale, 20 * scale, (P, 4))
    bin_sum = totals.astype(np.int64).sum(0)
    owner_total = np.array([bin_sum[o:n_bins:P].sum() for o in range(P)])
    caps = u32(np.maximum((owner_total.mean() * cap_scale * rng.uniform(0.9, 1.1, P)), 1).astype(np.int64))
    ok = bool(np.all(owner_total <= caps))
    split_at, share_at, seg_len, seg_cap = 8192, 12288, 8192, 1 << 20
    S = 1 << lS
    res = []
    for me in range(P):
        scratch, items, tickets = np.zeros(4 * MAX_BINS, np.uint32), np.zeros(16 * MAX_BINS, np.uint32), u32([7, 7, 7, 7, 0, 7, 7, 0])
        map_next, info_next, n_next = np.zeros(T, np.uint32), np.zeros(MAX_BINS, np.uint32), np.zeros(1, np.uint32)
        seg_desc, seg_of = np.zeros(4 * MAX_BINS, np.uint32), np.zeros(MAX_BINS, np.uint32)
        out = Out()
        ph.ph_owners_plan(T, lS, ptr(map_), ptr(n_bins_a), ptr(info), ptr(totals), P, me, ptr(caps), split_at, share_at, seg_at, seg_len, seg_cap,
                          ptr(scratch), ptr(items), ptr(tickets), ptr(map_next), ptr(info_next), ptr(n_next), ptr(seg_desc), ptr(seg_of), C.byref(out))
        res.append(dict(scat=scratch[MAX_BINS:2 * MAX_BINS].copy(), begin=scratch[2 * MAX_BINS:3 * MAX_BINS].copy(),
                        count=scratch[3 * MAX_BINS:].copy(), items=items, tickets=tickets, map_next=map_next, info_next=info_next,
                        n_next=int(n_next[0]), out=(out.total, out.needed, out.overflow, out.n_items), bin_sum=scratch[:MAX_BINS].copy()))
    for me, r in enumerate(res):
        assert r["out"][:3] == (int(totals[me].astype(np.int64).sum()), int(owner_total[me]), 0 if ok else 1)
        assert np.array_equal(r["bin_sum"][:n_bins], bin_sum[:n_bins])
        assert np.array_equal(r["map_next"], res[0]["map_next"]) and r["n_next"] == res[0]["n_next"]
        assert np.array_equal(r["info_next"], res[0]["info_next"])
    strip_frags = [int(bin_sum[first[t]:first[t] + (1 << ls[t])].sum()) for t in range(T)]
    check_next_map(strip_frags, int(bin_sum.sum()), split_at, lS, res[0]["map_next"], res[0]["info_next"], res[0]["n_next"])
    if not ok:
        for r in res:                                                     # nobody scatters, nobody folds
            assert r["out"][3] == 0 and not r["scat"][:n_bins].any() and not r["count"][:n_bins].any()
        return
    for o in range(P):
        mine = np.arange(o, n_bins, P)
        if len(mine) == 0:                                                    # more ranks than bins: this one owns nothing
            assert res[o]["out"][3] == 0
            continue
        begin = np.concatenate([[0], np.cumsum(bin_sum[mine])[:-1]])          # the owner's bins, one after the other
        assert np.array_equal(res[o]["begin"][mine], begin) and np.array_equal(res[o]["count"][mine], bin_sum[mine])
        before = np.zeros(len(mine), np.int64)
        for r in range(P):                                                    # inside a bin: rank order = draw order
            assert np.array_equal(res[r]["scat"][mine], begin + before), (o, r)
            before += totals[r, mine]
        seg_len_eff = max(int(owner_total[o]) // (2 * MAX_BINS), seg_len)

        def expect_lp(b, n, R):
            if seg_at and 2 * R <= S and n > seg_at:
                lp = 1
                while lp < 4 and (n >> lp) > seg_len_eff:
                    lp += 1
                return lp, 1
            return lparts_shared(n, R, share_at), 0

        check_items(res[o]["items"], res[o]["out"][3], bin_sum, lambda b: S >> (int(info[b]) >> 24), mine.tolist(), expect_lp)
        assert seg_at == 0 or sum(int(r["tickets"][5]) for r in res) > 0
        assert (res[o]["tickets"][5] > 0) == any(expect_lp(b, int(bin_sum[b]), S >> (int(info[b]) >> 24))[1] for b in mine.tolist() if bin_sum[b])
