"""The oracle against the REFERENCE'S OWN SHADER TEXT.

tests/golden/glsl_v1.npz holds inputs and outputs of the reference's glslified shaders (logic.frag,
flow/index.vert, spawn/{init,ball}, spawn/pixels/{index,best,bright,data,flow}-sample) executed by
tools/glsl_interp.py -- an interpreter that shares no code with the oracle (tools/make_glsl_golden.py,
run where /root/reference exists).  The C oracle has to reproduce every value bit for bit."""
import os

import numpy as np
import pytest

from util import assert_bits_equal

GOLD = os.path.join(os.path.dirname(__file__), "golden", "glsl_v1.npz")
STATE = dict(damping=0.043, speedLimit=0.01, forceWeight=0.016, varyForce=-0.1, flowWeight=1.0, varyFlow=0.2,
             noiseWeight=0.002, varyNoise=0.3, flowDecay=0.005, noiseScale=2.125, varyNoiseScale=0.5,
             noiseSpeed=0.00025, varyNoiseSpeed=0.1, target=0.002, varyTarget=1.5)


@pytest.fixture(scope="module")
def g():
    return np.load(GOLD)


def test_logic_frag(oracle, g):
    O = oracle
    P = O.make_params(viewSize=tuple(g["logic_viewSize"]), **STATE)
    time, dt = g["logic_time_dt"]
    out = O.integrate(P, g["logic_state"], g["logic_targets"], g["logic_flow"], time, dt)
    assert_bits_equal(out, g["logic_out"], "logic.frag")
    assert np.isfinite(g["logic_out"][1:]).all() and (g["logic_out"][0, 1, 0] == -1e6)
    P2 = O.make_params(viewSize=tuple(g["logic_viewSize"]), **{**STATE, "noiseWeight": 0.0, "target": 0.0})
    out2 = O.integrate(P2, g["logic_state"], g["logic_targets"], g["logic_flow"], time, dt)
    assert_bits_equal(out2, g["logic_out_nonoise"], "logic.frag without noise/target")
    assert not np.array_equal(g["logic_out"], g["logic_out_nonoise"])


def test_flow_vert(oracle, g):
    O = oracle
    P = O.make_params(viewSize=tuple(g["logic_viewSize"]), **STATE)
    cur, prev, want = g["logic_state"], g["vert_prev"], g["vert_out"]
    time = g["logic_time_dt"][0]
    R = cur.shape[0]
    n_written = 0
    for i in range(R):
        for j in range(2 * R):
            ok, out = O.flow_vertex(P, cur, prev, i, j, time)
            assert ok == bool(want[i, j, 0]), (i, j)
            if ok:
                n_written += 1
                assert_bits_equal(out, want[i, j, 1:7], f"flow vertex ({i},{j})")
    assert n_written > R * 2 * R - 8


@pytest.mark.parametrize("name", ["init", "ball", "direct", "best", "bright", "data", "flow"])
def test_spawn_frags(oracle, g, name):
    O = oracle
    R = g["logic_state"].shape[0]
    want = g["spawn_" + name]
    if name == "init":
        got = O.spawn_init(R, R)
    elif name == "ball":
        got = O.spawn_ball(R, R, 0.3, 0.005)
    else:
        sx, sy, jx, jy, time, speed, bias = g["spawn_uniforms"]
        S = O.make_spawn_pixels(spawnSize=(sx, sy), jitter=(jx, jy), speed=speed, bias=bias,
                                spawnMatrix=(-1, 0, 0, 0, 1, 0, 0, 0, 1), flowDecay=STATE["flowDecay"])
        st = g["logic_state"]
        src = {"direct": g["spawn_image"], "best": g["spawn_image"], "bright": g["spawn_image"],
               "data": np.ascontiguousarray(g["vert_prev"].transpose(1, 0, 2)), "flow": g["logic_flow"]}[name]
        if name == "direct":
            got = O.spawn_pixels_direct(S, R, R, src, time)
        else:
            got = O.spawn_pixels_sample(S, name, st, src, time)
    assert_bits_equal(got, want, f"spawn {name}")
    if name not in ("init",):
        assert np.isfinite(want).all() and len(np.unique(want[..., 0])) > 3


def test_optical_flow_frag(oracle, g):
    """optical-flow/index.frag: the interpreter gives the fragment colour; blended over a zero grid the oracle must
    give colour*alpha per channel (dst = src*a + 0*(1-a))."""
    O = oracle
    sx, sy, offset, lam, speed, limit, time = g["of_uniforms"]
    H, W = g["of_frag"].shape[:2]
    flow = np.zeros((H, W, 4), np.float32)
    O.optical_flow(flow, g["of_view"], g["of_last"], viewSize=tuple(g["logic_viewSize"]), scaleUV=(sx, sy), offset=offset,
                   lambda_=lam, speed=speed, speedLimit=limit, time=time)
    frag = g["of_frag"]
    want = (frag * frag[..., 3:4]).astype(np.float32)         # one rounding per channel, as the blend does
    assert_bits_equal(flow, want, "optical flow blended onto zero")
    assert (frag[..., 3] > 0).mean() > 0.5 and np.isfinite(frag).all()
