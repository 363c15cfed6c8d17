"""GPU parity: the CUDA path (through the C ABI) against the CPU oracle, bit for bit.

Tolerance: NONE -- integrate, spawners and the ordered flow blend all have to match the oracle's 32-bit patterns exactly
(only the payload of a NaN is left open).  The whole module runs three times: as the library ships on one GPU, with the
opaque pruning that sharded runs may use (TB_PRUNE=1), and with the segmented fold forced onto small bins -- none may change a bit."""
import ctypes as C
import os

import numpy as np
import pytest

from util import assert_bits_equal, synthetic_image

pytestmark = pytest.mark.gpu

DT = 1000 / 60


@pytest.fixture(scope="module")
def T():
    import tendrils_b200
    tendrils_b200.load()
    return tendrils_b200


MODES = {"default": {}, "prune": {"TB_PRUNE": "1"}, "segments": {"TB_SEG_AT": "96", "TB_SEG_LEN": "64", "TB_SPLIT_AT": "512"}}


@pytest.fixture(scope="module", params=list(MODES), autouse=True)
def prune_mode(request):
    """The knobs are read when a context is created.  `segments`: every split bin above 96 fragments is folded in segments of
    64 (PARITY B4: bracketing chains, records, the join) -- the path sharded runs and crowded draws depend on."""
    keys = sorted({k for m in MODES.values() for k in m})
    old = {k: os.environ.get(k) for k in keys}
    for k in keys:
        os.environ.pop(k, None)
    os.environ.update(MODES[request.param])
    yield request.param
    for k in keys:
        os.environ.pop(k, None)
        if old[k] is not None:
            os.environ[k] = old[k]


def frags_ok(t, n):
    """fragments the last splat blended: all the oracle rasterised -- or, with pruning, at most as many"""
    got = t.particles.stats()["last_fragments"]
    return got <= n if os.environ.get("TB_PRUNE") == "1" else got == n


def make(T, R, G, state=None, view=None):
    w, h = (G, G) if view is None else view
    t = T.Tendrils(T.Device(w, h))
    if state:
        t.state.update(state)
    t.setup(R)
    t.resize()
    return t


def oracle_params(O, t):
    keys = O.DEFAULT_STATE.keys()
    return O.make_params(viewSize=tuple(np.float32(v) for v in t.viewSize), **{k: t.state[k] for k in keys})


class OracleSim:
    """The reference's call order restated on the oracle (src/demo.main.js:1024-1082)."""

    def __init__(self, O, R, W, H, P):
        self.O, self.P = O, P
        self.cur = O.spawn_init(R, R)
        self.prev = O.spawn_init(R, R)
        self.targets = np.zeros((R, R, 4), np.float32)
        self.flow = np.zeros((H, W, 4), np.float32)

    def spawn_ball(self, radius, speed):
        self.prev, self.cur = self.cur, self.O.spawn_ball(*self.cur.shape[:2], radius, speed)

    def step(self, time, dt):
        new = self.O.integrate(self.P, self.cur, self.targets, self.flow, time, dt)
        self.prev, self.cur = self.cur, new

    def draw(self, time):
        return self.O.splat(self.P, self.cur, self.prev, self.flow, time)


@pytest.mark.parametrize("R,G", [(64, 32), (96, 50)])
def test_ball_then_steps_bit_exact(T, oracle, R, G):
    from tendrils_b200.spawn import spawnBall
    t = make(T, R, G)
    sim = OracleSim(oracle, R, G, G, oracle_params(oracle, t))
    ball = spawnBall(t.gl, {"uniforms": {"radius": 0.3, "speed": 0.005}})
    ball.spawn(t)
    sim.spawn_ball(0.3, 0.005)
    assert_bits_equal(t.particles.buffers[0].download(), sim.cur, "ball spawn")
    for k in range(25):
        t.timer.tick()
        t.step().draw()
        sim.step(np.float32(t.timer.time), np.float32(t.timer.dt))
        n = sim.draw(np.float32(t.timer.time))
        assert frags_ok(t, n), f"fragment count at step {k}"
        assert_bits_equal(t.particles.buffers[0].download(), sim.cur, f"state after step {k}")
        assert_bits_equal(t.flow.download(), sim.flow, f"flow after step {k}")
    assert_bits_equal(t.particles.buffers[1].download(), sim.prev, "previous state")
    assert np.isfinite(sim.cur).all() and (sim.flow != 0).any()


@pytest.mark.parametrize("R,G,radius,steps", [(192, 1100, 0.9, 4), (160, 2048, 0.9, 3), (256, 64, 0.03, 12), (192, 2048, 0.01, 6)])
def test_strip_sizes_and_split_maps_bit_exact(T, oracle, prune_mode, R, G, radius, steps):
    """Grids beyond 1024^2 (strips of 256 / 512 texels), and balls so small that strips get crowded: the split map the plan
    derives from one draw (2, 4, ... 256 bins per strip) must not change the next draw's result."""
    from tendrils_b200.spawn import spawnBall
    t = make(T, R, G)
    sim = OracleSim(oracle, R, G, G, oracle_params(oracle, t))
    spawnBall(t.gl, {"uniforms": {"radius": radius, "speed": 0.005}}).spawn(t)
    sim.spawn_ball(radius, 0.005)
    for k in range(steps):
        t.timer.tick()
        t.step().draw()
        sim.step(np.float32(t.timer.time), np.float32(t.timer.dt))
        n = sim.draw(np.float32(t.timer.time))
        assert frags_ok(t, n), f"fragment count at step {k}"
        assert_bits_equal(t.flow.download(), sim.flow, f"flow after step {k}")
    assert_bits_equal(t.particles.buffers[0].download(), sim.cur, "state")
    if prune_mode == "segments" and radius < 0.1:
        seg = t.particles.segment_stats()
        assert seg["bins"] > 0 and seg["segments"] >= 2 * seg["bins"], seg      # the crowd was folded in segments, and joined


@pytest.mark.parametrize("R,G,chunks", [(96, 50, 16), (64, 32, 1), (40, 24, 64)])
def test_streamed_step_bit_exact(T, oracle, R, G, chunks):
    """Tendrils.stepStreamed: the state comes from and goes back to ONE host buffer every step (tb_step_streamed, chunked
    copies on two streams around the chunked logic pass) -- same results as upload + step + download, flow splat included."""
    from tendrils_b200.spawn import spawnBall
    t = make(T, R, G)
    sim = OracleSim(oracle, R, G, G, oracle_params(oracle, t))
    spawnBall(t.gl, {"uniforms": {"radius": 0.3, "speed": 0.005}}).spawn(t)
    sim.spawn_ball(0.3, 0.005)
    host = t.particles.buffers[0].download()
    for k in range(10):
        t.timer.tick()
        if k == 6:                                   # an ordinary step in between: the copies are joined, the buffers rotate as two
            t.particles.sync()
            t.particles.buffers[0].upload(host)
            t.step().draw()
            host = t.particles.buffers[0].download()
        else:
            t.stepStreamed(host, host, chunks).draw()
        sim.step(np.float32(t.timer.time), np.float32(t.timer.dt))
        sim.draw(np.float32(t.timer.time))
        if k in (0, 5, 6, 9):
            t.particles.sync()
            assert_bits_equal(host, sim.cur, f"host copy of the state after step {k}")
            assert_bits_equal(t.flow.download(), sim.flow, f"flow after step {k}")
            assert_bits_equal(t.particles.buffers[0].download(), sim.cur, f"device state after step {k}")
            assert_bits_equal(t.particles.buffers[1].download(), sim.prev, f"previous state after step {k}")


# ---------------------------------------------------------------------------------------------
# spawners
# ---------------------------------------------------------------------------------------------
def _spawner_uniforms(G):
    # jitter = aspect(Float32Array(2), viewRes, jitterRad) (src/spawn/pixels/index.js:55)
    j = np.float32(np.float32(1.0 / G) * 2.0)
    return j


@pytest.mark.parametrize("variant", ["direct", "best", "bright", "color", "data", "flow"])
def test_pixel_spawners_bit_exact(T, oracle, variant):
    from tendrils_b200.spawn import PixelSpawner, spawnBall
    from tendrils_b200.spawn import pixels as PX
    R, G = 48, 40
    t = make(T, R, G)
    O = oracle
    P = oracle_params(O, t)
    sim = OracleSim(O, R, G, G, P)
    # some history first so that `particles` and the flow grid are non-trivial
    spawnBall(t.gl, {"uniforms": {"radius": 0.6, "speed": 0.004}}).spawn(t)
    sim.spawn_ball(0.6, 0.004)
    for _ in range(4):
        t.timer.tick(); t.step().draw()
        sim.step(np.float32(t.timer.time), np.float32(t.timer.dt)); sim.draw(np.float32(t.timer.time))
    img = synthetic_image(37, 29)
    frag = {"direct": PX.pixelsFrag, "best": PX.bestSampleFrag, "bright": PX.brightSampleFrag,
            "color": PX.colorSampleFrag, "data": PX.dataSampleFrag, "flow": PX.flowSampleFrag}[variant]
    buf = img
    if variant == "flow":
        buf = t.flow
    if variant == "data":
        buf = t.particles.buffers[0]
    sp = PixelSpawner(t.gl, {"shader": frag, "buffer": buf, "speed": 0.7, "bias": 0.9, "jitterRad": 2,
                             "spawnSize": [0.9, 1.1]})
    sp.spawnMatrix = PX.mat3_scale(PX.mat3_identity(), [-1, 1])
    sp.spawn(t)
    time = np.float32(t.timer.time)
    j = _spawner_uniforms(G)
    S = O.make_spawn_pixels(spawnSize=(0.9, 1.1), jitter=(j, j), speed=0.7, bias=0.9,
                            spawnMatrix=(-1, 0, 0, 0, 1, 0, 0, 0, 1), flowDecay=t.state["flowDecay"])
    if variant == "flow":
        src = sim.flow
    elif variant == "data":
        src = np.ascontiguousarray(sim.cur.transpose(1, 0, 2))    # texture (x across, y rows) = [row y][col x]
    else:
        src = img
    if variant == "direct":
        want = O.spawn_pixels_direct(S, R, R, src, time)
    else:
        want = O.spawn_pixels_sample(S, variant, sim.cur, src, time)
    assert_bits_equal(t.particles.buffers[0].download(), want, f"{variant} spawn")
    assert_bits_equal(t.particles.buffers[1].download(), sim.cur, "previous state after spawn")


def test_spawn_into_targets_and_target_pull(T, oracle):
    """spawnImageTargets (src/demo.main.js:517-521): the direct spawn written into tendrils.targets with
    no ping-pong rotation, then steps with target != 0."""
    from tendrils_b200.spawn import PixelSpawner, spawnBall
    from tendrils_b200.spawn import pixels as PX
    R, G = 40, 32
    t = make(T, R, G, state={"target": 0.003, "varyTarget": 2.0})
    O = oracle
    sim = OracleSim(O, R, G, G, oracle_params(O, t))
    spawnBall(t.gl, {"uniforms": {"radius": 0.5, "speed": 0.002}}).spawn(t)
    sim.spawn_ball(0.5, 0.002)
    img = synthetic_image(64, 64)
    sp = PixelSpawner(t.gl, {"shader": PX.pixelsFrag, "buffer": img, "speed": 0.3})
    sp.spawn(t, None, t.targets)
    j = _spawner_uniforms(G)
    S = O.make_spawn_pixels(jitter=(j, j), speed=0.3)
    sim.targets = O.spawn_pixels_direct(S, R, R, img, np.float32(t.timer.time))
    assert_bits_equal(t.targets.download(), sim.targets, "targets")
    assert_bits_equal(t.particles.buffers[0].download(), sim.cur, "state untouched by a targets spawn")
    for k in range(6):
        t.timer.tick(); t.step().draw()
        sim.step(np.float32(t.timer.time), np.float32(t.timer.dt)); sim.draw(np.float32(t.timer.time))
        assert_bits_equal(t.particles.buffers[0].download(), sim.cur, f"state {k}")
        assert_bits_equal(t.flow.download(), sim.flow, f"flow {k}")


# ---------------------------------------------------------------------------------------------
# edge cases of the reference: inert texels, NaN (0/0 in the speed clamp), out of bounds,
# non-square view, odd sizes, custom CPU spawn upload
# ---------------------------------------------------------------------------------------------
def test_edge_states_bit_exact(T, oracle):
    R, W, H = 34, 48, 20
    t = make(T, R, 0, view=(W, H), state={"noiseScale": 40.0, "varyNoiseScale": -50.0, "flowWeight": -0.4,
                                          "speedLimit": 0.08, "flowDecay": 0.0005})
    O = oracle
    rng = np.random.default_rng(5)
    st = np.zeros((R, R, 4), np.float32)
    st[..., 0:2] = rng.uniform(-1.3, 1.3, (R, R, 2))
    st[..., 2:4] = rng.normal(0, 0.03, (R, R, 2))
    st[0, :, 0:2] = T.INERT                      # a column of inert texels: pass through unchanged
    st[0, :, 2:4] = 0
    st[1, 0:4] = (0.25, -0.5, 0.0, 0.0)          # at rest, zero force possible -> 0/0 = NaN (Q2)
    st[2, 0] = (np.nan, 0.1, 0.0, 0.01)
    st[2, 1] = (0.1, 0.1, np.inf, 0.0)
    st[2, 2] = (5.0e7, -3.0e9, 0.01, 0.0)        # far out of bounds; beyond the noise fast-path guard
    st[3, :, 0] = 0.999999                       # hugging the right edge
    st[4, :, 1] = -1.0

    def fill(data, x, y):
        data[:] = st[x, y]
    t.spawn(fill)
    assert t.viewSize[0] == pytest.approx(1.0) and t.viewSize[1] == pytest.approx(W / H)
    sim = OracleSim(O, R, W, H, oracle_params(O, t))
    sim.cur, sim.prev = st.copy(), st.copy()
    assert_bits_equal(t.particles.buffers[0].download(), st, "cpu spawn upload")
    assert_bits_equal(t.particles.buffers[1].download(), st, "cpu spawn upload (all buffers)")
    for k in range(12):
        t.timer.tick(); t.step().draw()
        sim.step(np.float32(t.timer.time), np.float32(t.timer.dt)); n = sim.draw(np.float32(t.timer.time))
        assert frags_ok(t, n)
        assert_bits_equal(t.particles.buffers[0].download(), sim.cur, f"state {k}")
        assert_bits_equal(t.flow.download(), sim.flow, f"flow {k}")
    assert np.isnan(sim.cur[2, 0]).any() and (sim.cur[0, :, 0] == T.INERT).all()


@pytest.mark.parametrize("weights", [dict(noiseWeight=0.0), dict(noiseWeight=0.0, flowWeight=0.0),
                                     dict(forceWeight=0.0), dict(damping=0.0)])
def test_fast_paths_are_exact(T, oracle, weights):
    """noiseWeight == 0 lets the kernel skip the simplex noise; the result must not change."""
    from tendrils_b200.spawn import spawnBall
    R, G = 64, 64
    t = make(T, R, G, state=weights)
    sim = OracleSim(oracle, R, G, G, oracle_params(oracle, t))
    spawnBall(t.gl, {"uniforms": {"radius": 0.8, "speed": 0.006}}).spawn(t)
    sim.spawn_ball(0.8, 0.006)
    for k in range(8):
        t.timer.tick(); t.step().draw()
        sim.step(np.float32(t.timer.time), np.float32(t.timer.dt)); sim.draw(np.float32(t.timer.time))
        assert_bits_equal(t.particles.buffers[0].download(), sim.cur, f"state {k}")
        assert_bits_equal(t.flow.download(), sim.flow, f"flow {k}")


def test_hot_texel_and_long_lines(T, oracle):
    """Maximum contention (every particle in one texel) and lines crossing the whole grid."""
    R, G = 64, 24
    t = make(T, R, G, state={"speedLimit": 3.0})
    O = oracle
    rng = np.random.default_rng(11)
    cur = np.zeros((R, R, 4), np.float32)
    cur[..., 0:2] = 0.3 + rng.uniform(0, 0.01, (R, R, 2))
    cur[..., 2:4] = rng.normal(0, 0.004, (R, R, 2))
    prev = cur.copy()
    prev[..., 0:2] -= cur[..., 2:4]
    prev[: R // 2, :, 0:2] = rng.uniform(-2.5, 2.5, (R // 2, R, 2))      # long lines, partly off-grid
    t.particles.buffers[0].upload(cur)
    t.particles.buffers[1].upload(prev)
    flow0 = rng.normal(0, 0.01, (G, G, 4)).astype(np.float32)
    t.flow.upload(flow0)
    sim = OracleSim(O, R, G, G, oracle_params(O, t))
    sim.cur, sim.prev, sim.flow = cur.copy(), prev.copy(), flow0.copy()
    t.timer.tick()
    t.draw()
    n = sim.draw(np.float32(t.timer.time))
    assert frags_ok(t, n) and n > 5000
    assert_bits_equal(t.flow.download(), sim.flow, "flow after a contended draw")


def test_sharded_contexts_equal_single(T, oracle):
    """Two column-sharded contexts folded in rank order onto one grid == one context (bit for bit)."""
    import ctypes as C
    from tendrils_b200 import _native as N
    from tendrils_b200.spawn import spawnBall
    R, G = 64, 48
    ts = [T.Tendrils(T.Device(G, G, rank=r, world_size=2)) for r in range(2)]
    for t in ts:
        t.setup(R); t.resize()
        assert (t.particles.col0, t.particles.col1) == ((0, 32), (32, 64))[t.gl.rank]
    one = make(T, R, G)
    sim = OracleSim(oracle, R, G, G, oracle_params(oracle, one))
    for t in ts + [one]:
        t.particles_world = 1
    L = N.load()
    ball = dict(uniforms={"radius": 0.4, "speed": 0.005})
    # spawn + step are purely local; run them through the public API with world_size forced to the ring-free path
    for t in ts + [one]:
        spawnBall(t.gl, ball).spawn(t)
    sim.spawn_ball(0.4, 0.005)
    for k in range(6):
        for t in ts + [one]:
            t.timer.tick(); t.step()
        one.draw()
        flow = None
        for r, t in enumerate(ts):                      # the ordered ring, by hand, on one GPU
            ctx = t.particles._ctx
            st = T.tendrils._state_struct({**t.state, "viewSize": t.viewSize})
            N.check(ctx, L.tb_set_state(ctx, C.byref(st)))
            N.check(ctx, L.tb_splat_collect(ctx, float(t.timer.time)))
            if flow is not None:
                t.flow.upload(flow)
            N.check(ctx, L.tb_splat_fold(ctx))
            flow = t.flow.download()
        ts[0].flow.upload(flow)                         # the "broadcast"
        sim.step(np.float32(one.timer.time), np.float32(one.timer.dt)); sim.draw(np.float32(one.timer.time))
        assert_bits_equal(one.flow.download(), sim.flow, f"single flow {k}")
        assert_bits_equal(flow, sim.flow, f"sharded flow {k}")
        got = np.concatenate([t.particles.buffers[0].download() for t in ts], 0)
        assert_bits_equal(got, sim.cur, f"sharded state {k}")


def test_opaque_cut_is_exact(T, oracle):
    """The fold starts at the last alpha == 1 fragment of a texel.  That must not change a single
    bit -- including for texels that already hold Inf / NaN (which the blend can never heal)."""
    R, G = 64, 16
    t = make(T, R, G, state={"speedLimit": 0.01})
    O = oracle
    rng = np.random.default_rng(3)
    cur = np.zeros((R, R, 4), np.float32)
    cur[..., 0:2] = rng.uniform(-0.9, 0.9, (R, R, 2))
    speed = rng.choice([0.002, 0.01, 0.0100001, 0.02, 0.05], size=(R, R)).astype(np.float32)   # many at/over the limit
    ang = rng.uniform(0, 2 * np.pi, (R, R))
    cur[..., 2] = speed * np.cos(ang)
    cur[..., 3] = speed * np.sin(ang)
    prev = cur.copy()
    prev[..., 0:2] -= 8 * cur[..., 2:4]              # lines of 0..3 texels
    prev[..., 2:4] = cur[..., 2:4]                   # both vertices get the same alpha
    flow0 = rng.normal(0, 0.01, (G, G, 4)).astype(np.float32)
    flow0[2, 3] = np.inf
    flow0[5, 5, 0] = -np.inf
    flow0[7, 1] = np.nan
    flow0[9, 9, 2] = np.nan
    flow0[11:13] = 0.0
    t.particles.buffers[0].upload(cur)
    t.particles.buffers[1].upload(prev)
    t.flow.upload(flow0)
    sim = OracleSim(O, R, G, G, oracle_params(O, t))
    sim.cur, sim.prev, sim.flow = cur.copy(), prev.copy(), flow0.copy()
    for k in range(3):
        t.timer.tick()
        t.draw()
        n = sim.draw(np.float32(t.timer.time))
        assert frags_ok(t, n)
        got = t.flow.download()
        assert_bits_equal(got, sim.flow, f"flow after draw {k}")
    assert (sim.flow[..., 3] == 1.0).sum() > 20, "the case must actually contain opaque fragments"
    assert np.isnan(sim.flow).any()


def test_optical_flow_bit_exact(T, oracle):
    """f1: the optical-flow pass drawn into the flow grid after the particle splat (src/demo.main.js:1131-1159)."""
    from tendrils_b200.optical_flow import OpticalFlow
    from tendrils_b200.spawn import spawnBall
    R, W, H = 32, 40, 24
    t = make(T, R, 0, view=(W, H))
    O = oracle
    sim = OracleSim(O, R, W, H, oracle_params(O, t))
    spawnBall(t.gl, {"uniforms": {"radius": 0.5, "speed": 0.004}}).spawn(t)
    sim.spawn_ball(0.5, 0.004)
    rng = np.random.default_rng(8)
    of = OpticalFlow(t.gl, None, {"speed": 0.08, "offset": 0.1, "scaleUV": [-1, -1]})
    frames = [rng.integers(0, 256, (18, 30, 4), dtype=np.uint8)]
    for _ in range(4):
        frames.append(np.clip(frames[-1].astype(np.int32) + rng.integers(-30, 31, frames[-1].shape), 0, 255).astype(np.uint8))
    of.resize([30, 18])
    last = np.zeros_like(frames[0])
    for k, frame in enumerate(frames):
        t.timer.tick(); t.step().draw()
        sim.step(np.float32(t.timer.time), np.float32(t.timer.dt)); sim.draw(np.float32(t.timer.time))
        of.setPixels(frame)
        of.update({"speedLimit": t.state["speedLimit"], "time": t.timer.time, "viewSize": t.viewSize}).render(t)
        O.optical_flow(sim.flow, frame, last, viewSize=tuple(np.float32(v) for v in t.viewSize), scaleUV=(-1, -1), offset=0.1,
                       lambda_=0.001, speed=0.08, speedLimit=t.state["speedLimit"], time=np.float32(t.timer.time))
        of.step(); last = frame
        assert_bits_equal(t.flow.download(), sim.flow, f"flow after optical-flow pass {k}")
        assert_bits_equal(t.particles.buffers[0].download(), sim.cur, f"state {k}")


# ---------------------------------------------------------------------------------------------
# degenerate and extreme shapes
# ---------------------------------------------------------------------------------------------
def test_all_inert_and_tiny_shapes(T, oracle):
    """Nothing to integrate, nothing to draw; the smallest texture (2x2) and a 1x1 flow grid."""
    t = make(T, 2, 1)
    O = oracle
    sim = OracleSim(O, 2, 1, 1, oracle_params(O, t))
    for _ in range(3):                                   # everything inert after setup(): state and flow stay put
        t.timer.tick(); t.step().draw()
        sim.step(np.float32(t.timer.time), np.float32(t.timer.dt)); n = sim.draw(np.float32(t.timer.time))
        assert n == 0 and t.particles.stats()["last_fragments"] == 0
    assert_bits_equal(t.particles.buffers[0].download(), sim.cur, "inert state")
    assert (t.flow.download() == 0).all()
    from tendrils_b200.spawn import spawnBall
    spawnBall(t.gl, {"uniforms": {"radius": 0.9, "speed": 0.05}}).spawn(t)
    sim.spawn_ball(0.9, 0.05)
    for k in range(5):
        t.timer.tick(); t.step().draw()
        sim.step(np.float32(t.timer.time), np.float32(t.timer.dt)); sim.draw(np.float32(t.timer.time))
        assert_bits_equal(t.particles.buffers[0].download(), sim.cur, f"2x2 state {k}")
        assert_bits_equal(t.flow.download(), sim.flow, f"1x1 flow {k}")


def test_non_square_particle_texture(T, oracle):
    """Particles accepts any [w, h] (src/particles.js:56); the sharded runs use tall textures."""
    from tendrils_b200.spawn import spawnBall
    W, H = 40, 24
    t = T.Tendrils(T.Device(W, H))
    t.setup([12, 50]); t.resize()
    O = oracle
    P = oracle_params(O, t)
    cur, prev = O.spawn_ball(12, 50, 0.6, 0.006), O.spawn_init(12, 50)
    spawnBall(t.gl, {"uniforms": {"radius": 0.6, "speed": 0.006}}).spawn(t)
    targets, flow = np.zeros((12, 50, 4), np.float32), np.zeros((H, W, 4), np.float32)
    for k in range(6):
        t.timer.tick(); t.step().draw()
        new = O.integrate(P, cur, targets, flow, np.float32(t.timer.time), np.float32(t.timer.dt))
        prev, cur = cur, new
        O.splat(P, cur, prev, flow, np.float32(t.timer.time))
        assert_bits_equal(t.particles.buffers[0].download(), cur, f"state {k}")
        assert_bits_equal(t.flow.download(), flow, f"flow {k}")


def test_redraw_after_parameter_change_recounts(T, oracle):
    """The fragment count rides in the integrate kernel; a draw whose viewSize / grid differs from the step's
    must not use it."""
    from tendrils_b200.spawn import spawnBall
    R = 48
    t = make(T, R, 32)
    O = oracle
    spawnBall(t.gl, {"uniforms": {"radius": 0.7, "speed": 0.008}}).spawn(t)
    cur, prev = O.spawn_ball(R, R, 0.7, 0.008), O.spawn_init(R, R)
    targets = np.zeros((R, R, 4), np.float32)
    t.timer.tick(); t.step()
    P = oracle_params(O, t)
    new = O.integrate(P, cur, targets, np.zeros((32, 32, 4), np.float32), np.float32(t.timer.time), np.float32(t.timer.dt))
    prev, cur = cur, new
    t.gl.drawingBufferWidth, t.gl.drawingBufferHeight = 56, 24          # the canvas was resized between step and draw
    t.resize()
    t.draw()
    P2 = oracle_params(O, t)
    flow = np.zeros((24, 56, 4), np.float32)
    n = O.splat(P2, cur, prev, flow, np.float32(t.timer.time))
    assert frags_ok(t, n)
    assert_bits_equal(t.flow.download(), flow, "flow after a resized draw")


@pytest.mark.parametrize("steps", [2])
def test_full_size_cfg3_bit_exact(T, oracle, steps):
    """BASELINE.json configs[2] at FULL size (4096^2 particles, 1024^2 flow grid): direct image spawn, then every
    step compared with the oracle bit for bit -- state (16.8 M particles) and flow grid (1 M texels)."""
    from tendrils_b200.spawn import PixelSpawner
    from tendrils_b200.spawn import pixels as PX
    R, G = 4096, 1024
    t = make(T, R, G)
    O = oracle
    img = synthetic_image(G, G)
    sp = PixelSpawner(t.gl, {"shader": PX.pixelsFrag, "buffer": img, "speed": 0.3, "jitterRad": 2, "spawnSize": [1, 1]})
    sp.spawnMatrix = PX.mat3_scale(PX.mat3_identity(), [-1, 1])
    sp.spawn(t)
    j = np.float32(np.float32(1.0 / G) * 2.0)
    S = O.make_spawn_pixels(jitter=(j, j), speed=0.3, spawnMatrix=(-1, 0, 0, 0, 1, 0, 0, 0, 1))
    cur = O.spawn_pixels_direct(S, R, R, img, np.float32(t.timer.time))
    prev = O.spawn_init(R, R)
    assert_bits_equal(t.particles.buffers[0].download(), cur, "direct spawn at full size")
    P = oracle_params(O, t)
    targets, flow = np.zeros((R, R, 4), np.float32), np.zeros((G, G, 4), np.float32)
    for k in range(steps):
        t.timer.tick(); t.step().draw()
        new = O.integrate(P, cur, targets, flow, np.float32(t.timer.time), np.float32(t.timer.dt))
        prev, cur = cur, new
        n = O.splat(P, cur, prev, flow, np.float32(t.timer.time), mt=True)
        assert frags_ok(t, n) and n > 10_000_000
        assert_bits_equal(t.particles.buffers[0].download(), cur, f"full-size state {k}")
        assert_bits_equal(t.flow.download(), flow, f"full-size flow {k}")
    # size-independent properties (what the reference guarantees by construction)
    sp_ = np.hypot(cur[..., 2], cur[..., 3])
    assert sp_.max() <= np.float32(0.01) * (1 + 1e-6)                 # speedLimit clamp
    assert (flow[..., 3] >= 0).all() and (flow[..., 3] <= 1).all()    # alpha is a convex combination
    assert flow[..., 2].max() <= np.float32(t.timer.time)             # time stamps never exceed `time`
    t.dispose()


def test_determinism_full_size(T):
    """Run-to-run determinism at cfg2 size: no atomics decide any value, so two runs agree bit for bit."""
    from tendrils_b200.spawn import spawnBall
    outs = []
    for _ in range(2):
        t = make(T, 2048, 512)
        spawnBall(t.gl, {"uniforms": {"radius": 0.3, "speed": 0.005}}).spawn(t)
        for _ in range(12):
            t.timer.tick(); t.step().draw()
        outs.append((t.particles.buffers[0].download(), t.flow.download()))
        t.dispose()
    assert np.array_equal(outs[0][0], outs[1][0]) and np.array_equal(outs[0][1], outs[1][1])


def test_device_resident_inputs_equal_host_inputs(T):
    """Frames and spawn images may already live on the device (unified addressing): same bits as host inputs."""
    import torch
    from tendrils_b200.optical_flow import OpticalFlow
    from tendrils_b200.spawn import PixelSpawner
    from tendrils_b200.spawn import pixels as PX
    from util import synthetic_video
    R, G = 64, 48
    frames = synthetic_video(G, G, 4)
    outs = []
    for on_device in (False, True):
        conv = (lambda a: torch.as_tensor(a, device="cuda")) if on_device else (lambda a: a)
        t = make(T, R, G)
        of = OpticalFlow(t.gl, None, {"speed": 0.08, "offset": 0.1, "scaleUV": [-1, -1]})
        of.resize([G, G])
        sp = PixelSpawner(t.gl, {"shader": PX.pixelsFrag, "buffer": conv(frames[0].astype(np.float32) / np.float32(255)),
                                 "speed": 0.3, "jitterRad": 2, "spawnSize": [1, 1]})
        best = PixelSpawner(t.gl, {"shader": PX.bestSampleFrag, "buffer": None, "speed": 1, "bias": 1, "jitterRad": 2,
                                   "spawnSize": [1, 1]})
        sp.spawn(t)
        for k in range(6):
            if k == 3:
                best.setPixels(conv(frames[k % 4].astype(np.float32) / np.float32(255)))
                best.spawn(t)
            t.timer.tick(); t.step().draw()
            of.setPixels(conv(frames[k % 4]))
            of.update({"speedLimit": t.state["speedLimit"], "time": t.timer.time, "viewSize": t.viewSize}).render(t)
            of.step()
        outs.append((t.particles.buffers[0].download(), t.flow.download()))
        assert np.abs(outs[-1][1]).max() > 0
        t.dispose()
    assert_bits_equal(outs[1][0], outs[0][0], "state: device-resident vs host inputs")
    assert_bits_equal(outs[1][1], outs[0][1], "flow: device-resident vs host inputs")


@pytest.mark.parametrize("seed,closed,view", [(11, False, (48, 32)), (12, True, (40, 56)), (13, False, (64, 64))])
def test_flow_lines_bit_exact(T, oracle, seed, closed, view):
    """f4: FlowLine.update().draw() into the flow grid after the particle splat (src/demo.main.js:1107-1121)."""
    from test_flow_line import random_path
    from tendrils_b200.spawn import spawnBall
    R, (W, H) = 32, view
    t = make(T, R, 0, view=(W, H))
    O = oracle
    sim = OracleSim(O, R, W, H, oracle_params(O, t))
    spawnBall(t.gl, {"uniforms": {"radius": 0.5, "speed": 0.004}}).spawn(t)
    sim.spawn_ball(0.5, 0.004)
    rng = np.random.default_rng(seed)
    lines = T.FlowLines(t.gl)
    fl = lines.get(1, {"closed": closed})
    path = random_path(rng, 18)
    if seed == 13:
        path.insert(7, path[6])                                           # a repeated point: Inf / NaN miters, culled triangles
    clock = 2000.0
    for k, p in enumerate(path):
        clock += float(rng.uniform(0.3, 30.0))
        fl.add(clock, p)
        if k < 3 or k % 5:
            continue
        t.timer.tick(); t.step().draw()
        sim.step(np.float32(t.timer.time), np.float32(t.timer.dt)); sim.draw(np.float32(t.timer.time))
        fl.line.uniforms.update(t.state)                                  # Object.assign(flowLine.line.uniforms, tendrils.state)
        fl.update().draw(t)
        u = fl.line.uniforms
        U = O.flow_line_uniforms(viewSize=u["viewSize"], rad=u["rad"], speed=u["speed"], speedLimit=u["speedLimit"],
                                 crestShape=u["crestShape"])
        a = {name: np.ascontiguousarray(v["data"], np.float32) for name, v in fl.line.attributes.items()}
        with np.errstate(all="ignore"):
            n = O.flow_line(U, a, sim.flow)
        assert n > 20
        assert_bits_equal(t.flow.download(), sim.flow, f"flow after the flow line, {k + 1} points")
        assert_bits_equal(t.particles.buffers[0].download(), sim.cur, f"state, {k + 1} points")
    assert lines.trim(1 / t.state["flowDecay"], clock) > 0


@pytest.mark.parametrize("PW,PH,G", [(24, 47, 40), (16, 83, 32)])
def test_rows_drawn_twice_bit_exact(T, oracle, PW, PH, G):
    """Heights whose D6 vertex table draws some texel rows TWICE (two line pairs land on one row; most
    non-power-of-two heights do): the count cannot ride in k_integrate there.  Found by tests/test_splat_host.py."""
    from tendrils_b200.spawn import spawnBall
    row, cur_of = oracle.vertex_table(PH)
    drawn = [int(row[2 * k]) for k in range(PH) if not (row[2 * k] == row[2 * k + 1] and cur_of[2 * k] == cur_of[2 * k + 1])]
    assert len(drawn) > len(set(drawn))                                       # the premise: a row is drawn twice
    t = T.Tendrils(T.Device(G, G))
    t.setup([PW, PH]); t.resize()
    t.state["speedLimit"] = 0.2
    O = oracle
    P = oracle_params(O, t)
    cur, prev = O.spawn_ball(PW, PH, 0.6, 0.15), O.spawn_init(PW, PH)
    spawnBall(t.gl, {"uniforms": {"radius": 0.6, "speed": 0.15}}).spawn(t)
    targets, flow = np.zeros((PW, PH, 4), np.float32), np.zeros((G, G, 4), np.float32)
    for k in range(5):
        t.timer.tick(); t.step().draw()
        new = O.integrate(P, cur, targets, flow, np.float32(t.timer.time), np.float32(t.timer.dt))
        prev, cur = cur, new
        n = O.splat(P, cur, prev, flow, np.float32(t.timer.time))
        assert frags_ok(t, n) and n > 100
        assert_bits_equal(t.particles.buffers[0].download(), cur, f"state {k}")
        assert_bits_equal(t.flow.download(), flow, f"flow {k}")
