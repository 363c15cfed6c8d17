"""f4: pointer flow lines.  CPU tests: the host geometry mirror, the oracle against the reference's
shader text (tests/golden/glsl_flowline_v1.npz), the raster rules, and the CUDA kernels' __host__ __device__ core
compiled for the CPU against the oracle.  The GPU parity test is in tests/test_gpu_parity.py."""
import ctypes as C
import math
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_fp = C.POINTER(C.c_float)


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def same(a, b):
    a, b = np.asarray(a, np.float32), np.asarray(b, np.float32)
    return np.array_equal(np.isnan(a), np.isnan(b)) and np.array_equal(a[~np.isnan(a)], b[~np.isnan(b)])


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(ROOT, "tests", "golden", "glsl_flowline_v1.npz"))


# ---- host geometry ------------------------------------------------------------------------------------------------
def test_polyline_normals_straight_and_right_angle():
    from tendrils_b200.flow_line import polyline_normals
    out = polyline_normals([[0, 0], [1, 0], [2, 0]])
    assert [o[1] for o in out] == [1.0, 1.0, 1.0]
    assert all(o[0][1] == 1.0 and o[0][0] == 0.0 for o in out)           # normal (-0, 1)
    out = polyline_normals([[0, 0], [1, 0], [1, 1]])
    s = 1 / math.sqrt(2)
    assert out[1][0] == pytest.approx([-s, s]) and out[1][1] == pytest.approx(math.sqrt(2))
    assert out[2][0] == pytest.approx([-1.0, 0.0]) and out[2][1] == 1.0
    # a repeated point: gl-vec2 normalize leaves the zero direction alone, the miter length becomes Infinity / NaN
    out = polyline_normals([[0, 0], [0, 0], [1, 0]])
    assert len(out) == 3 and not math.isfinite(out[1][1]) or out[1][1] == 1.0
    assert polyline_normals([[0.3, 0.1]]) == [] and polyline_normals([]) == []


def test_polyline_normals_closed_loop():
    from tendrils_b200.flow_line import polyline_normals
    sq = [[0, 0], [1, 0], [1, 1], [0, 1]]
    out = polyline_normals(sq, closed=True)
    assert len(out) == 4                                                   # one per point (the duplicate is popped)
    for n, m in out:
        assert m == pytest.approx(math.sqrt(2)) and math.hypot(*n) == pytest.approx(1.0)


def test_flow_line_attributes_mirror_the_js():
    import tendrils_b200 as T
    fl = T.FlowLine(T.Device(8, 8))
    for k, p in enumerate([[0.0, 0.0], [0.5, 0.0], [0.5, 0.5]]):
        fl.add(100.0 + 16.0 * k, p)
    fl.update()
    a = {k: v["data"] for k, v in fl.line.attributes.items()}
    assert a["position"].dtype == np.float32 and a["position"].shape == (12,) and a["miter"].shape == (6,)
    assert list(a["position"]) == [0, 0, 0, 0, 0.5, 0, 0.5, 0, 0.5, 0.5, 0.5, 0.5]
    assert list(a["previous"]) == [0, 0, 0, 0, 0, 0, 0, 0, 0.5, 0, 0.5, 0]        # max(0, p-1)
    assert list(a["time"]) == [100, 100, 116, 116, 132, 132] and list(a["dt"]) == [0, 0, 16, 16, 16, 16]
    assert a["miter"][0] == -1 and a["miter"][1] == 1                              # even vertices flipped
    assert a["miter"][2] == pytest.approx(-math.sqrt(2)) and a["miter"][3] == pytest.approx(math.sqrt(2))
    assert fl.line.vertex_count() == 6
    assert fl.trim(20, 130) == 2 and fl.length == 2 and fl.line.path[0] == [0.5, 0.0]  # points older than now-ago go
    lines = T.FlowLines(T.Device(8, 8))
    assert lines.get(7) is lines.get(7)
    lines.get(7).add(1.0, [0, 0])
    assert lines.trim(5, 100) == 0 and lines.active == {}


# ---- oracle against the reference's shader text ------------------------------------------------------------------
def test_vertex_and_fragment_stage_match_reference_glsl(oracle, gold):
    u = gold["uniforms"]
    U = oracle.flow_line_uniforms(viewSize=(u[0], u[1]), rad=u[2], speed=u[3], speedLimit=u[4], crestShape=u[5])
    for vin, vout in zip(gold["vert_in"], gold["vert_out"]):
        got = oracle.flow_line_vertex(U, vin[0:2], vin[2:4], vin[4], vin[5:7], vin[7], vin[8])
        assert same(got, vout), (vin, got, vout)
    for fin, fout in zip(gold["frag_in"], gold["frag_out"]):
        got = oracle.flow_line_fragment(u[5], fin)
        assert same(got, fout), (fin, got, fout)


# ---- raster rules (spec/PARITY.md FL3-FL5) -----------------------------------------------------------------------
def quad_attributes(x0, y0, x1, y1, W, H, time=50.0):
    """A strip of 4 vertices whose expanded positions are exactly the window rectangle [x0,x1]x[y0,y1] (pixels):
    rad 0 would collapse the strip, so the corners are given as positions with miter 0 (no expansion)."""
    to_ndc = lambda x, y: (x / (W / 2) - 1.0, y / (H / 2) - 1.0)
    corners = [to_ndc(x0, y0), to_ndc(x0, y1), to_ndc(x1, y0), to_ndc(x1, y1)]
    pos = np.array(corners, np.float32).reshape(-1)
    prev = pos.copy(); prev[0::2] -= 0.1                                    # moving right: alpha = min(|vel|/limit, 1) = 1
    n = 4
    return {"position": pos, "normal": np.zeros(2 * n, np.float32), "miter": np.zeros(n, np.float32), "previous": prev,
            "time": np.full(n, time, np.float32), "dt": np.full(n, 16.0, np.float32)}


def test_shared_edges_draw_once_and_top_left_rule(oracle):
    W = H = 16
    U = oracle.flow_line_uniforms()
    flow = np.zeros((H, W, 4), np.float32)
    # pixel-aligned rectangle [3,11]x[2,7]: the centres x+.5 in (3,11), y+.5 in (2,7) -> 8*5 pixels, each hit ONCE
    # although the two triangles share the diagonal
    n = oracle.flow_line(U, quad_attributes(3, 2, 11, 7, W, H), flow)
    assert n == 8 * 5
    touched = flow[..., 2] != 0
    assert touched.sum() == 40 and touched[2:7, 3:11].all()
    # sdf = sign(0) = 0 -> d = 0: colour = (normalize(vel)*|vel|, time, alpha) with alpha = 1 -> an overwrite
    assert np.all(flow[2:7, 3:11, 2] == 50.0) and np.all(flow[2:7, 3:11, 3] == 1.0)
    # edges exactly through pixel centres: [3.5,10.5]x[2.5,6.5] keeps the left/bottom... exactly one side of each pair
    flow2 = np.zeros((H, W, 4), np.float32)
    n2 = oracle.flow_line(U, quad_attributes(3.5, 2.5, 10.5, 6.5, W, H), flow2)
    t2 = flow2[..., 2] != 0
    assert n2 == t2.sum() == 7 * 4                                          # 8x5 centres on or inside, one row and one column lost
    # two abutting rectangles tile: no pixel twice, none missing
    flow3 = np.zeros((H, W, 4), np.float32)
    n3 = oracle.flow_line(U, quad_attributes(3.5, 2.5, 7.5, 6.5, W, H), flow3) + \
        oracle.flow_line(U, quad_attributes(7.5, 2.5, 10.5, 6.5, W, H), flow3)
    assert n3 == n2 and np.array_equal(flow3[..., 2] != 0, t2)


def test_culls_and_degenerates(oracle):
    U = oracle.flow_line_uniforms()
    flow = np.zeros((8, 8, 4), np.float32)
    a = quad_attributes(1, 1, 6, 6, 8, 8)
    bad = {k: v.copy() for k, v in a.items()}
    bad["position"][0] = np.nan                                             # vertex 0 is in triangle 0 only
    assert oracle.flow_line(U, bad, flow) < 25 and not np.isnan(flow).any()
    flat = {k: v.copy() for k, v in a.items()}
    flat["position"][1::2] = flat["position"][1]                            # all on one row: zero area
    assert oracle.flow_line(U, flat, np.zeros((8, 8, 4), np.float32)) == 0
    two = {k: v[:2 * (2 if v.shape[0] == 8 else 1)] for k, v in a.items()}
    assert oracle.flow_line(U, two, np.zeros((8, 8, 4), np.float32)) == 0  # fewer than 3 vertices


# ---- the CUDA kernels' core, compiled for the host ---------------------------------------------------------------
@pytest.fixture(scope="module")
def host_core(tmp_path_factory):
    out = tmp_path_factory.mktemp("hh") / "libflowline_host.so"
    src = os.path.join(ROOT, "tests", "host_harness", "flowline_host.cpp")
    subprocess.run(["g++", "-O2", "-std=c++17", "-march=x86-64-v3", "-ffp-contract=off", "-fno-fast-math", "-fPIC", "-shared",
                    "-Wno-unknown-pragmas", "-o", str(out), src], check=True)
    L = C.CDLL(str(out))
    L.hh_flow_line.restype = C.c_longlong
    L.hh_flow_line.argtypes = [_fp, C.c_int, _fp, _fp, _fp, _fp, _fp, _fp, _fp, C.c_int, C.c_int]
    return L


def random_path(rng, n, step=0.22, jitter=0.9):
    p = [rng.uniform(-0.6, 0.6, 2)]
    ang = rng.uniform(0, 2 * np.pi)
    for _ in range(n - 1):
        ang += rng.normal(0, jitter)
        q = p[-1] + step * rng.uniform(0.2, 1.5) * np.array([np.cos(ang), np.sin(ang)])
        if np.abs(q).max() > 0.95:                                          # bounce off the edge of the view
            ang += np.pi
            q = np.clip(q, -0.95, 0.95)
        p.append(q)
    return [list(map(float, q)) for q in p]


@pytest.mark.parametrize("seed,closed,size", [(1, False, (48, 32)), (2, False, (40, 40)), (3, True, (33, 47)), (4, False, (64, 64))])
def test_kernel_core_on_host_equals_oracle(oracle, host_core, seed, closed, size):
    import tendrils_b200 as T
    rng = np.random.default_rng(seed)
    W, H = size
    fl = T.FlowLine(T.Device(W, H), {"closed": closed})
    t = 1000.0
    for p in random_path(rng, 14 + seed):
        t += float(rng.uniform(0.3, 30.0))
        fl.add(t, p)
    if seed == 4:                                                           # a repeated point: Inf / NaN miters -> culled triangles
        fl.add(t + 5.0, fl.line.path[-1])
        fl.add(t + 9.0, [0.1, -0.2])
    fl.line.uniforms.update({"speedLimit": 0.01, "viewSize": [1.0, W / H] if seed == 2 else [1, 1]})
    fl.update()
    a = {k: np.ascontiguousarray(v["data"], np.float32) for k, v in fl.line.attributes.items()}
    u = fl.line.uniforms
    U = oracle.flow_line_uniforms(viewSize=u["viewSize"], rad=u["rad"], speed=u["speed"], speedLimit=u["speedLimit"],
                                  crestShape=u["crestShape"])
    base = rng.normal(0, 0.004, (H, W, 4)).astype(np.float32)
    want = base.copy()
    with np.errstate(all="ignore"):
        frags = oracle.flow_line(U, a, want)
    got = base.copy()
    u6 = np.array([u["viewSize"][0], u["viewSize"][1], u["rad"], u["speed"], u["speedLimit"], u["crestShape"]], np.float32)
    p = lambda x: x.ctypes.data_as(_fp)
    host_core.hh_flow_line(p(u6), a["miter"].shape[0], p(a["position"]), p(a["normal"]), p(a["miter"]), p(a["previous"]),
                           p(a["time"]), p(a["dt"]), p(got), W, H)
    assert frags > 50
    assert same(got, want)
    assert not np.array_equal(bits(want), bits(base))


STRESS = int(os.environ.get("TB_STRESS", "0"))        # TB_STRESS=n: n extra random configurations (off in the normal run)


@pytest.mark.parametrize("seed", range(STRESS))
def test_stress_random_paths(oracle, host_core, seed):
    import tendrils_b200 as T
    rng = np.random.default_rng(50_000 + seed)
    W, H = int(rng.integers(1, 90)), int(rng.integers(1, 90))
    fl = T.FlowLine(T.Device(W, H), {"closed": bool(rng.random() < 0.3)})
    t = 1000.0
    for p in random_path(rng, int(rng.integers(2, 40)), step=float(rng.choice([0.01, 0.1, 0.4, 1.5])), jitter=float(rng.choice([0.1, 0.9, 3.0]))):
        t += float(rng.choice([0.0, 0.4, 16.0, 300.0]))
        fl.add(t, p)
        if rng.random() < 0.1:
            fl.add(t, p)                                                     # a repeated point
    fl.line.uniforms.update({"speedLimit": float(rng.choice([0.01, 0.2, 0.0])), "rad": float(rng.choice([0.1, 0.5, 0.01])),
                             "speed": float(rng.choice([3, 0.1, 100])), "crestShape": float(rng.choice([0.6, 0.0, 1.0, 5.0])),
                             "viewSize": [1.0, W / H] if rng.random() < 0.5 else [1, 1]})
    fl.update()
    a = {k: np.ascontiguousarray(v["data"], np.float32) for k, v in fl.line.attributes.items()}
    u = fl.line.uniforms
    U = oracle.flow_line_uniforms(viewSize=u["viewSize"], rad=u["rad"], speed=u["speed"], speedLimit=u["speedLimit"], crestShape=u["crestShape"])
    base = rng.normal(0, 0.004, (H, W, 4)).astype(np.float32)
    want, got = base.copy(), base.copy()
    with np.errstate(all="ignore"):
        oracle.flow_line(U, a, want)
    u6 = np.array([u["viewSize"][0], u["viewSize"][1], u["rad"], u["speed"], u["speedLimit"], u["crestShape"]], np.float32)
    p = lambda x: x.ctypes.data_as(_fp)
    host_core.hh_flow_line(p(u6), a["miter"].shape[0], p(a["position"]), p(a["normal"]), p(a["miter"]), p(a["previous"]),
                           p(a["time"]), p(a["dt"]), p(got), W, H)
    assert same(got, want), seed
