"""GPU parity: the CUDA path (through the C ABI) against the CPU oracle, bit for bit.

Tolerance: NONE for finite values -- integrate, spawners and the ordered flow blend all have to
match the oracle exactly (any NaN equals any NaN, -0 equals +0)."""
import ctypes as C

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
        assert t.particles.stats()["last_fragments"] == n, f"fragment count at step {k}"
        assert_bits_equal(t.particles.buffers[0].download(), sim.cur, f"state after step {k}")
        assert_bits_equal(t.flow.download(), sim.flow, f"flow after step {k}")
    assert_bits_equal(t.particles.buffers[1].download(), sim.prev, "previous state")
    assert np.isfinite(sim.cur).all() and (sim.flow != 0).any()
