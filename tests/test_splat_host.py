"""Host-side tables of the flow splat, cut out of tendrils_b200/csrc/tb_api.cu unchanged and compiled for the CPU by this
test: the D6 pair table (build_pairs) against the oracle's vertex table, and the strip geometry (choose_geom) for its
invariants.  Nothing here is used by the product."""
import ctypes as C
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

HARNESS = r'''
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <vector>
namespace tb {
struct PairEntry { int32_t k, row_a, row_b, pad; };
constexpr int kMaxStrips = %(max_strips)s;
struct StripGeom { int W, H, sxl, syl, strips_x, strips_y, T; };
}
using namespace tb;
namespace {
%(pairs)s
%(geom)s
}
extern "C" int sh_pairs(int PH, int *out5, int cap) {
    const std::vector<PairEntry> pairs = build_pairs(PH);
    if ((int)pairs.size() > cap) return -1;
    for (size_t i = 0; i < pairs.size(); ++i) {
        out5[5 * i] = pairs[i].k; out5[5 * i + 1] = pairs[i].row_a & 0x7fffffff; out5[5 * i + 2] = pairs[i].row_a < 0;
        out5[5 * i + 3] = pairs[i].row_b & 0x7fffffff; out5[5 * i + 4] = pairs[i].row_b < 0;
    }
    return (int)pairs.size();
}
extern "C" void sh_geom(int W, int H, int *out6) {
    const StripGeom g = choose_geom(W, H);
    out6[0] = g.sxl; out6[1] = g.syl; out6[2] = g.strips_x; out6[3] = g.strips_y; out6[4] = g.T; out6[5] = kMaxStrips;
}
'''


@pytest.fixture(scope="module")
def sh(tmp_path_factory):
    d = tmp_path_factory.mktemp("sh")
    asrc = open(os.path.join(ROOT, "tendrils_b200", "csrc", "tb_api.cu")).read()
    ssrc = open(os.path.join(ROOT, "tendrils_b200", "csrc", "tb_splat.cuh")).read()
    pairs = asrc[asrc.index("int host_texel(float u, int size) {"):asrc.index("// column sampled by vertex column i")]
    geom = asrc[asrc.index("StripGeom choose_geom(int W, int H) {"):asrc.index("int tiles_release(tb_ctx *c);")]
    max_strips = ssrc[ssrc.index("constexpr int kMaxStrips = ") + len("constexpr int kMaxStrips = "):].split(";")[0]
    cpp = d / "splat_host.cpp"
    cpp.write_text(HARNESS % {"pairs": pairs, "geom": geom, "max_strips": max_strips})
    out = d / "libsplat_host.so"
    subprocess.run(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fno-fast-math", "-fPIC", "-shared", "-Wno-unused-function",
                    "-o", str(out), str(cpp)], check=True)
    L = C.CDLL(str(out))
    L.sh_pairs.restype = C.c_int
    L.sh_pairs.argtypes = [C.c_int, C.POINTER(C.c_int), C.c_int]
    L.sh_geom.restype = None
    L.sh_geom.argtypes = [C.c_int, C.c_int, C.POINTER(C.c_int)]
    return L


def test_pair_table_equals_oracle_vertex_table(sh, oracle):
    """build_pairs (tb_api.cu) against the oracle's vertex table (D6) for every height up to 3000 and the large ones."""
    for PH in list(range(1, 3001)) + [4096, 5000, 8192, 12288, 16384, 28672, 32768, 65536, 100003]:
        row, cur = oracle.vertex_table(PH)
        want = [(k, int(row[2 * k]), int(cur[2 * k]), int(row[2 * k + 1]), int(cur[2 * k + 1])) for k in range(PH)
                if not (row[2 * k] == row[2 * k + 1] and cur[2 * k] == cur[2 * k + 1])]
        buf = (C.c_int * (5 * PH))()
        n = sh.sh_pairs(PH, buf, PH)
        got = [tuple(buf[5 * i:5 * i + 5]) for i in range(n)]
        assert got == want, PH


@pytest.mark.parametrize("W,H", [(1, 1), (16, 8), (17, 9), (64, 64), (256, 256), (1024, 1024), (1000, 37), (2048, 2048), (2048, 1024),
                                 (4096, 4096), (32768, 1), (1, 32768), (5000, 3000)])
def test_strip_geometry(sh, W, H):
    out = (C.c_int * 6)()
    sh.sh_geom(W, H, out)
    sxl, syl, sx, sy, T, max_strips = list(out)
    assert sx == -(-W // (1 << sxl)) and sy == -(-H // (1 << syl)) and T == sx * sy
    assert 1 <= T <= max_strips                                     # one cursor per bin in shared memory
    assert sxl >= 4 and syl >= 3                                    # 16 x 8 texels unless the grid needs larger strips
    if W * H <= 1024 * 1024:
        assert (sxl, syl) == (4, 3)
