"""Run the REFERENCE'S OWN shader text through tools/glsl_interp.py and commit inputs + outputs as golden
vectors (tests/golden/glsl_v1.npz).  Needs /root/reference (this container only); the test that consumes the
vectors (tests/test_glsl_golden.py) does not.

Shader text: the glslified sources embedded in the reference's built source maps
  docs/js/index.js.map : ./src/logic.frag, ./src/flow/index.vert
  docs/js/demo.js.map  : ./src/spawn/{init,ball}/index.frag, ./src/spawn/pixels/{index,best-sample,
                         bright-sample,data-sample,flow-sample}.frag
(the third-party glsl-noise / glsl-random / glsl-map code is inlined there, node_modules being absent).
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from tools.glsl_interp import Shader, nearest_sampler  # noqa: E402
from util import synthetic_image  # noqa: E402

REF = "/root/reference"
f32 = np.float32
DT = 1000 / 60


def shader_source(map_file, suffix):
    m = json.load(open(os.path.join(REF, "docs", "js", map_file)))
    for name, content in zip(m["sources"], m["sourcesContent"]):
        if name.split("?")[0].endswith(suffix):
            body = content.split("module.exports = ", 1)[1]
            end = body.index('"\n')                      # the JSON string literal ends at the first bare quote + newline
            return json.loads(body[:end + 1])
    raise KeyError(suffix)


STATE = dict(damping=0.043, speedLimit=0.01, forceWeight=0.016, varyForce=-0.1, flowWeight=1.0, varyFlow=0.2,
             noiseWeight=0.002, varyNoise=0.3, flowDecay=0.005, noiseScale=2.125, varyNoiseScale=0.5,
             noiseSpeed=0.00025, varyNoiseSpeed=0.1, target=0.002, varyTarget=1.5)


def state_texture(R, rng):
    st = np.zeros((R, R, 4), f32)                          # x-major [x][y]
    st[..., 0:2] = rng.uniform(-1.1, 1.1, (R, R, 2))
    st[..., 2:4] = rng.normal(0, 0.006, (R, R, 2))
    st[0, 1] = (-1e6, -1e6, 0.0, 0.0)                      # inert
    st[1, 0, 2:4] = (0.02, -0.03)                          # over the speed limit
    return st


def tex_of_state(st):
    """the GL texture holding an x-major state array: [row y][col x]"""
    return np.ascontiguousarray(st.transpose(1, 0, 2))


def run_fragments(sh, R, uniforms):
    out = np.zeros((R, R, 4), f32)
    for x in range(R):
        for y in range(R):
            g = sh.run({**uniforms, "gl_FragCoord": [x + 0.5, y + 0.5, 0.0, 1.0]})
            out[x, y] = g["gl_FragColor"]
    return out


def main():
    rng = np.random.default_rng(2024)
    out = {}
    R, W, H = 5, 7, 4
    view_size = np.array([1.0, f32(W / H)], f32)

    # ---- logic.frag ---------------------------------------------------------------------------
    logic = Shader(shader_source("index.js.map", "/src/logic.frag"))
    st = state_texture(R, rng)
    targets = np.zeros((R, R, 4), f32); targets[..., 0:2] = rng.uniform(-0.5, 0.5, (R, R, 2))
    flow = rng.normal(0, 0.01, (H, W, 4)).astype(f32); flow[..., 2] = rng.uniform(60, 130, (H, W)); flow[1, 2] = 0
    time, dt = f32(7 * DT), f32(DT)
    uni = {**{k: f32(v) for k, v in STATE.items()}, "particles": nearest_sampler(tex_of_state(st)),
           "flow": nearest_sampler(flow), "targets": nearest_sampler(tex_of_state(targets)),
           "dataRes": [R, R], "viewSize": view_size, "time": time, "dt": dt}
    out["logic_state"], out["logic_targets"], out["logic_flow"] = st, targets, flow
    out["logic_time_dt"] = np.array([time, dt], f32)
    out["logic_viewSize"] = view_size
    out["logic_out"] = run_fragments(logic, R, uni)
    # noise off / no target: the paths the CUDA kernel shortcuts
    uni2 = dict(uni); uni2["noiseWeight"] = f32(0); uni2["target"] = f32(0)
    out["logic_out_nonoise"] = run_fragments(logic, R, uni2)

    # ---- flow/index.vert ----------------------------------------------------------------------
    vert = Shader(shader_source("index.js.map", "/src/flow/index.vert"))
    prev = state_texture(R, rng)
    vout = np.zeros((R, 2 * R, 7), f32)
    for i in range(R):
        for j in range(2 * R):
            uv = [f32(i * (1 / (R - 1))), f32(j * (1 / (2 * R - 1)))]        # Particles.generateLUT, Float32Array
            g = vert.run({"previous": nearest_sampler(tex_of_state(prev)), "data": nearest_sampler(tex_of_state(st)),
                          "dataRes": [R, R], "viewSize": view_size, "time": time, "speedLimit": f32(STATE["speedLimit"]),
                          "flowDecay": f32(STATE["flowDecay"]), "uv": uv})
            written = "gl_Position" in g
            vout[i, j, 0] = 1.0 if written else 0.0
            if written:
                vout[i, j, 1:3] = g["gl_Position"][0:2]
                vout[i, j, 3:7] = g["color"]
    out["vert_prev"], out["vert_out"] = prev, vout

    # ---- spawners -----------------------------------------------------------------------------
    init = Shader(shader_source("demo.js.map", "/src/spawn/init/index.frag"))
    out["spawn_init"] = run_fragments(init, R, {})
    ball = Shader(shader_source("demo.js.map", "/src/spawn/ball/index.frag"))
    out["spawn_ball"] = run_fragments(ball, R, {"radius": f32(0.3), "speed": f32(0.005)})
    img = synthetic_image(6, 4)
    out["spawn_image"] = img
    sp = {"dataRes": [R, R], "geomRes": [R, 2 * R], "spawnSize": [0.9, 1.1],
          "jitter": [f32(f32(1 / W) * 2), f32(f32(1 / H) * 2)], "time": f32(11 * DT), "speed": f32(0.7),
          "spawnMatrix": [-1, 0, 0, 0, 1, 0, 0, 0, 1], "bias": f32(0.9), "flowDecay": f32(STATE["flowDecay"]),
          "particles": nearest_sampler(tex_of_state(st))}
    out["spawn_uniforms"] = np.array([0.9, 1.1, sp["jitter"][0], sp["jitter"][1], sp["time"], 0.7, 0.9], f32)
    for name, suffix, data in [("direct", "/src/spawn/pixels/index.frag", img),
                               ("best", "/src/spawn/pixels/best-sample.frag", img),
                               ("bright", "/src/spawn/pixels/bright-sample.frag", img),
                               ("data", "/src/spawn/pixels/data-sample.frag", tex_of_state(prev)),
                               ("flow", "/src/spawn/pixels/flow-sample.frag", flow)]:
        sh = Shader(shader_source("demo.js.map", suffix))
        out["spawn_" + name] = run_fragments(sh, R, {**sp, "spawnData": nearest_sampler(data)})
    # ---- optical-flow/index.frag (f1), drawn over a W x H flow grid with the big triangle ----------------
    of = Shader(shader_source("demo.js.map", "/src/optical-flow/index.frag"))
    iw, ih = 9, 6
    view = rng.integers(0, 256, (ih, iw, 4), dtype=np.uint8)
    last = np.clip(view.astype(np.int32) + rng.integers(-40, 41, (ih, iw, 4)), 0, 255).astype(np.uint8)
    as_float = lambda img: (img.astype(f32) / f32(255.0)).astype(f32)           # RGBA8 sampler: byte/255
    ofu = {"view": nearest_sampler(as_float(view)), "last": nearest_sampler(as_float(last)), "viewSize": view_size,
           "scaleUV": [-1.0, -1.0], "offset": f32(0.1), "lambda": f32(0.001), "time": f32(9 * DT), "speed": f32(0.08),
           "speedLimit": f32(STATE["speedLimit"])}
    frag = np.zeros((H, W, 4), f32)
    for gy in range(H):
        for gx in range(W):
            uv = [f32(f32(f32(gx + 0.5) / f32(W)) * f32(2.0)) - f32(1.0), f32(f32(f32(gy + 0.5) / f32(H)) * f32(2.0)) - f32(1.0)]
            frag[gy, gx] = of.run({**ofu, "uv": uv})["gl_FragColor"]
    out["of_view"], out["of_last"], out["of_frag"] = view, last, frag
    out["of_uniforms"] = np.array([-1.0, -1.0, 0.1, 0.001, 0.08, STATE["speedLimit"], 9 * DT], f32)

    path = os.path.join(ROOT, "tests", "golden", "glsl_v1.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


def flow_line():
    """f4: ./src/flow-line/index.vert and index.frag (docs/js/demo.js.map) on random attribute / varying values ->
    tests/golden/glsl_flowline_v1.npz (a file of its own, so that glsl_v1.npz never has to be regenerated)."""
    rng = np.random.default_rng(4242)
    vert = Shader(shader_source("demo.js.map", "/src/flow-line/index.vert"))
    frag = Shader(shader_source("demo.js.map", "/src/flow-line/index.frag"))
    uni = {"viewSize": [1.0, f32(4 / 3)], "rad": f32(0.1), "speed": f32(3.0), "speedLimit": f32(0.01)}
    n = 64
    vin = np.zeros((n, 9), f32)                      # position.xy, normal.xy, miter, previous.xy, time, dt
    vin[:, 0:2] = rng.uniform(-1, 1, (n, 2))
    ang = rng.uniform(0, 2 * np.pi, n)
    vin[:, 2], vin[:, 3] = np.cos(ang), np.sin(ang)
    vin[:, 4] = rng.uniform(0.8, 3.0, n) * rng.choice([-1.0, 1.0], n)
    vin[:, 5:7] = vin[:, 0:2] + rng.normal(0, 0.02, (n, 2))
    vin[:, 7] = rng.uniform(1000, 90000, n)
    vin[:, 8] = rng.uniform(0.0, 40.0, n)
    vin[0, 5:7] = vin[0, 0:2]                        # first path point: previous = itself, dt = 0
    vin[0, 8] = 0.0
    vin[1, 4] = 0.0                                  # miter 0: sign() = 0
    vin[2, 8] = 0.25                                 # dt below 1 ms: max(dt, 1.0)
    vout = np.zeros((n, 9), f32)                     # gl_Position.xy, values.rgba, crest.xy, sdf
    with np.errstate(all="ignore"):
        for i in range(n):
            g = vert.run({**uni, "position": vin[i, 0:2], "normal": vin[i, 2:4], "miter": vin[i, 4], "previous": vin[i, 5:7],
                          "time": vin[i, 7], "dt": vin[i, 8]})
            vout[i, 0:2], vout[i, 2:6], vout[i, 6:8], vout[i, 8] = g["gl_Position"][0:2], g["values"], g["crest"], g["sdf"]
        fin = np.zeros((n, 7), f32)                  # values.rgba, crest.xy, sdf
        fin[:, 0:2] = rng.normal(0, 0.006, (n, 2))
        fin[:, 2] = rng.uniform(1000, 90000, n)
        fin[:, 3] = rng.uniform(0, 1, n)
        fin[:, 4:6] = rng.normal(0, 1.2, (n, 2))
        fin[:, 6] = rng.uniform(-1, 1, n)
        fin[0, 6] = 0.0                              # on the path: d = 0
        fin[1, 0:2] = 0.0                            # no velocity: direction comes from the crest alone
        fout = np.zeros((n, 4), f32)
        for i in range(n):
            fout[i] = frag.run({"crestShape": f32(0.6), "values": fin[i, 0:4], "crest": fin[i, 4:6], "sdf": fin[i, 6]})["gl_FragColor"]
    path = os.path.join(ROOT, "tests", "golden", "glsl_flowline_v1.npz")
    np.savez_compressed(path, uniforms=np.array([1.0, f32(4 / 3), 0.1, 3.0, 0.01, 0.6], f32), vert_in=vin, vert_out=vout,
                        frag_in=fin, frag_out=fout)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "flow-line":
        flow_line()
    else:
        main()
