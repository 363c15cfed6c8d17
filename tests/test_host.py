"""CPU tests of the host layer: the JS-mirror classes, the C-ABI library's exported symbols, and the
fail-loudly behaviour without a GPU."""
import ctypes
import math
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    """The .so loads without a GPU and exports exactly what include/tendrils_b200.h declares."""
    from tendrils_b200 import _native as N
    from tendrils_b200 import build as B
    B.build()
    hdr = open(os.path.join(ROOT, "include", "tendrils_b200.h")).read()
    declared = set(re.findall(r"^(?:int64_t|int|const char \*)\s*(tb_\w+)\s*\(", hdr, re.M))
    assert declared == set(N.SYMBOLS), (declared ^ set(N.SYMBOLS))
    lib = ctypes.CDLL(N.lib_path())
    for name in declared:
        assert hasattr(lib, name), name
    assert N.load().tb_abi_version() == 2


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import tendrils_b200 as T
    t = T.Tendrils(T.Device(64, 64))
    with pytest.raises(T.TendrilsError, match="no CUDA device"):
        t.setup(32)


def test_product_never_touches_the_oracle():
    """The oracle is test infrastructure: nothing under tendrils_b200/ may import or load it."""
    pkg = os.path.join(ROOT, "tendrils_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in txt.lower(), os.path.join(dirpath, f)


def test_timer_mirrors_src_timer_js():
    from tendrils_b200 import Timer, defaults
    t = defaults()["timer"]
    assert t.step == 1000 / 60 and t.time == 0 and t.dt == 0
    for k in range(1, 6):
        t.tick()
        assert t.dt == 1000 / 60
        assert t.time == pytest.approx(k * 1000 / 60, rel=0, abs=1e-9)
    # f64 accumulation, as in JS: time after 3 ticks is ((0+s)+s)+s, not 3*s rounded differently
    s = 1000 / 60
    assert t.time == ((((s + s) + s) + s) + s)
    t.paused = True
    before = t.time
    t.tick()
    assert t.time == before and t.dt == 0
    # loop / end (src/timer.js:43-55; the demo sets end = 600000, loop = true)
    u = Timer(0, 0); u.step = 400.0; u.end = 1000.0; u.loop = True
    for _ in range(3):
        u.tick()
    assert u.time == math.fmod(1200.0, 1000.0)
    v = Timer(0, 0); v.step = 400.0; v.end = 1000.0
    for _ in range(3):
        v.tick()
    assert v.time == 1000.0 and v.paused
    w = Timer(now=5000.0)
    assert w.time == 0 and w.since == 5000.0
    w.tick(now=5100.0)
    assert w.time == 100.0 and w.dt == 100.0


def test_aspect_mirrors_gl_matrix_storage():
    from tendrils_b200.aspect import aspect, coverAspect, f32vec2
    vs = [0, 0]
    coverAspect(vs, [1000, 500])                 # plain Array: doubles (src/index.js:139,398)
    assert vs == [1.0, (1 / 500) * 1000]
    j = aspect(f32vec2(), [1000, 1000], 2)       # Float32Array: rounds at every store (spawn/pixels/index.js:43,55)
    assert j.dtype == np.float32 and j[0] == np.float32(np.float32(1 / 1000) * 2)


def test_shard_columns_partition():
    from tendrils_b200 import shard_columns
    for width, world in [(4096, 1), (4096, 8), (10, 3), (7, 7), (32768, 8)]:
        blocks = [shard_columns(width, r, world) for r in range(world)]
        assert blocks[0][0] == 0 and blocks[-1][1] == width
        assert all(a[1] == b[0] for a, b in zip(blocks, blocks[1:]))          # contiguous, in rank order
        sizes = [b - a for a, b in blocks]
        assert max(sizes) - min(sizes) <= 1


def test_defaults_match_reference_state():
    """src/index.js:29-57"""
    from tendrils_b200 import defaults
    s = defaults()["state"]
    want = dict(rootNum=512, damping=0.043, speedLimit=0.01, forceWeight=0.016, varyForce=-0.1, flowWeight=1,
                varyFlow=0.2, noiseWeight=0.002, varyNoise=0.3, flowDecay=0.005, flowWidth=5, noiseScale=2.125,
                varyNoiseScale=0.5, noiseSpeed=0.00025, varyNoiseSpeed=0.1, target=0, varyTarget=1)
    for k, v in want.items():
        assert s[k] == v, k


def test_custom_shaders_are_rejected():
    import tendrils_b200 as T
    with pytest.raises(T.TendrilsError, match="custom logic shaders"):
        T.Tendrils(T.Device(8, 8), {"logicShader": "void main(){}"})


def test_pixel_spawner_uniforms():
    from tendrils_b200.spawn import PixelSpawner
    from tendrils_b200.spawn import pixels as PX
    sp = PixelSpawner(None, {"shader": PX.bestSampleFrag, "buffer": np.zeros((2, 2, 4), np.float32)})
    u = sp.update({"viewRes": [1024, 512]})
    assert u["speed"] == 1 and u["bias"] == 1 and list(u["spawnSize"]) == [1, 1]
    assert u["jitter"][0] == np.float32(np.float32(1 / 1024) * 2) and u["jitter"][1] == np.float32(np.float32(1 / 512) * 2)
    m = PX.mat3_scale(PX.mat3_identity(), [-1, 1])
    assert list(m) == [-1, 0, 0, 0, 1, 0, 0, 0, 1]


def test_optical_flow_mirror_buffers():
    """src/optical-flow/index.js: two RGBA8 buffers rotated by step(); resize keeps matching shapes, zeroes others."""
    from tendrils_b200.optical_flow import OpticalFlow, defaults
    of = OpticalFlow(None, None, {"speed": 0.08, "offset": 0.1, "scaleUV": [-1, -1]})
    assert of.uniforms["lambda"] == defaults()["uniforms"]["lambda"] == 0.001 and of.uniforms["speed"] == 0.08
    of.resize([4, 3])
    assert all(b.shape == (3, 4, 4) and b.dtype == np.uint8 for b in of.buffers)
    a = np.full((3, 4, 4), 7, np.uint8)
    of.setPixels(a)
    of.step()                                  # utils.step: pop the last buffer to the front
    assert (of.buffers[1] == 7).all() and (of.buffers[0] == 0).all()
    of.setPixels(np.full((3, 4, 4), 9, np.uint8))
    of.step()
    assert (of.buffers[0] == 7).all() and (of.buffers[1] == 9).all()
    of.update({"time": 5.0})
    assert of._bound["time"] == 5.0 and of._bound["offset"] == 0.1


def test_env_knobs_are_documented():
    """every TB_* environment variable read by the library is listed in DESIGN.md"""
    import glob
    names = set()
    for f in glob.glob(os.path.join(ROOT, "tendrils_b200", "**", "*"), recursive=True):
        if f.endswith((".cu", ".cuh", ".py")):
            names |= set(re.findall(r'"(TB_[A-Z_0-9]+)"', open(f).read()))
    doc = open(os.path.join(ROOT, "DESIGN.md")).read()
    missing = sorted(n for n in names if n not in doc)
    assert not missing, missing


def test_default_transport_by_world_size(monkeypatch):
    """sharded runs share the flow blend over peer memory ("owners") unless told otherwise; TB_RING overrides."""
    import tendrils_b200 as T
    monkeypatch.delenv("TB_RING", raising=False)
    assert T.Device(4, 4, world_size=2, rank=1).ring == "owners"
    assert T.Device(4, 4, world_size=8, rank=3).ring == "owners"
    assert T.Device(4, 4, world_size=8, rank=3, ring="dist").ring == "dist"
    monkeypatch.setenv("TB_RING", "dist")
    assert T.Device(4, 4, world_size=8).ring == "dist"


def test_synthetic_video_loops_and_moves():
    from util import synthetic_video
    a, b = synthetic_video(40, 24, 4), synthetic_video(40, 24, 4)
    assert all(f.shape == (24, 40, 4) and f.dtype == np.uint8 for f in a)
    assert all(np.array_equal(x, y) for x, y in zip(a, b))                    # closed form: no RNG
    assert not np.array_equal(a[0], a[1]) and (a[0][..., 3] == 255).all() and a[0][..., :3].max() > 100


def test_bench_workloads_and_bytes():
    import bench
    assert set(bench.WORKLOADS) >= {"cfg1", "cfg2", "cfg3", "cfg5"}
    n, g = 4096 * 4096, 1024 * 1024
    ab = bench.algorithmic_bytes(n, g)
    assert ab["step"] == 587_202_560
    assert bench.algorithmic_bytes(n, g, optical=True)["step"] == ab["step"] + 40 * g
    w5 = bench.WORKLOADS["cfg5"]
    assert w5["R"] * w5["rows"] == 2 ** 27 and w5["G"] == 2048


def test_napi_shim_matches_the_c_header():
    """bindings/node cannot be built here (no Node), but it must at least type-check against include/tendrils_b200.h:
    g++ -fsyntax-only with a stub of node_api.h (tests/host_harness/node_api.h)."""
    import subprocess
    r = subprocess.run(["g++", "-std=c++17", "-fsyntax-only", "-Wall", "-Wextra", "-I", os.path.join(ROOT, "tests", "host_harness"),
                        os.path.join(ROOT, "bindings", "node", "tendrils_b200_napi.cc")], capture_output=True, text=True)
    assert r.returncode == 0 and not r.stderr.strip(), r.stderr
    shim = open(os.path.join(ROOT, "bindings", "node", "tendrils_b200_napi.cc")).read()
    for sym in ("tb_create", "tb_step", "tb_splat_flow", "tb_spawn_pixels", "tb_optical_flow", "tb_flow_line", "tb_blend_into_flow",
                "tb_step_streamed", "tb_sync", "tb_set_overlap", "tb_stats"):
        assert sym + "(" in shim, sym


def test_buffer_download_into_caller_memory(monkeypatch):
    """_Buffer.download(out=...) hands the caller's (e.g. pinned) array to tb_download: same call, no extra host copy."""
    import ctypes as C
    from tendrils_b200 import _native as N
    from tendrils_b200 import tendrils as TT
    calls = []

    class FakeLib:
        def tb_download(self, ctx, which, ptr, n):
            calls.append((which, C.addressof(ptr.contents), n))
            np.ctypeslib.as_array(ptr, shape=(n,))[:] = np.arange(n, dtype=np.float32)
            return 0

    class Owner:
        _ctx = C.c_void_p(1)
        flow_shape, shape, col0, col1 = [5, 3], [4, 6], 1, 3

    monkeypatch.setattr(N, "load", lambda: FakeLib())
    buf = TT._Buffer(Owner(), N.TB_BUF_CURRENT)
    fresh = buf.download()
    assert fresh.shape == (2, 6, 4) and fresh.dtype == np.float32 and fresh[0, 0, 1] == 1.0
    mine = np.zeros((2, 6, 4), np.float32)
    got = buf.download(out=mine)
    assert got is mine and calls[-1] == (N.TB_BUF_CURRENT, mine.ctypes.data, 48) and mine[1, 5, 3] == 47.0
    flow = TT._Buffer(Owner(), N.TB_BUF_FLOW).download(out=np.zeros((3, 5, 4), np.float32))
    assert flow.shape == (3, 5, 4) and calls[-1][2] == 60
    for bad in (np.zeros((2, 6, 4), np.float64), np.zeros((6, 2, 4), np.float32), np.zeros((2, 6, 8), np.float32)[..., ::2]):
        with pytest.raises(N.TendrilsError):
            buf.download(out=bad)


def test_scatter_claims_reconverge_between_rounds():
    """k_splat_scatter claims bin slots with one shared-memory atomic per round and relies on round r's atomics being issued
    before round r + 1's (draw order).  In the shipped SASS that holds because the warp reconverges after every round's atomic:
    between two consecutive ATOMS there must be a BSYNC.RECONVERGENT (or a WARPSYNC)."""
    import shutil
    import subprocess
    lib = os.path.join(ROOT, "tendrils_b200", "lib", "libtendrils_b200.so")
    tool = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not (os.path.exists(lib) and os.path.exists(tool)):
        pytest.skip("needs the built library and cuobjdump")
    sass = subprocess.run([tool, "-sass", "-fun", "_ZN2tb15k_splat_scatterENS_11ScatterArgsE", lib], capture_output=True, text=True).stdout
    lines = [l for l in sass.splitlines() if "/*" in l and ";" in l]
    atoms = [i for i, l in enumerate(lines) if "ATOMS" in l]
    assert len(atoms) >= 5, "the claim loop's atomics are gone?"
    for a, b in zip(atoms, atoms[1:]):
        between = lines[a + 1:b]
        assert any("BSYNC.RECONVERGENT" in l or "WARPSYNC" in l for l in between), "two claim atomics without a reconvergence point in between"
