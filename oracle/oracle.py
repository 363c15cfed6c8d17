"""ctypes wrapper around the CPU ORACLE (oracle/tendrils_oracle.c).

TEST INFRASTRUCTURE ONLY.  May be imported from tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py -- never from tendrils_b200/.

PARITY UNPINNED by the reference's own tests (there are none); see tendrils_oracle.h.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libtendrils_oracle.so")

f32 = np.float32
_fp = C.POINTER(C.c_float)
_ip = C.POINTER(C.c_int)


class Params(C.Structure):
    """or_params -- uniforms of logic.frag (reference src/index.js:29-57 defaults)."""
    _fields_ = [(n, C.c_float) for n in (
        "damping", "speedLimit", "forceWeight", "varyForce", "flowWeight", "varyFlow",
        "noiseWeight", "varyNoise", "flowDecay", "flowWidth", "noiseScale", "varyNoiseScale",
        "noiseSpeed", "varyNoiseSpeed", "target", "varyTarget")] + [("viewSize", C.c_float * 2)]


class SpawnPixels(C.Structure):
    _fields_ = [("spawnSize", C.c_float * 2), ("jitter", C.c_float * 2), ("speed", C.c_float),
                ("bias", C.c_float), ("spawnMatrix", C.c_float * 9), ("flowDecay", C.c_float)]


class FlowLineUniforms(C.Structure):
    """or_flow_line_uniforms (reference src/flow-line/index.js:19-22 + src/geom/line/index.js:16-20)."""
    _fields_ = [("viewSize", C.c_float * 2), ("rad", C.c_float), ("speed", C.c_float), ("speedLimit", C.c_float),
                ("crestShape", C.c_float)]


APPLY_COLOR, APPLY_BRIGHTEST, APPLY_IDENTITY, APPLY_FLOW = 0, 1, 2, 3

DEFAULT_STATE = dict(damping=0.043, speedLimit=0.01, forceWeight=0.016, varyForce=-0.1,
                     flowWeight=1.0, varyFlow=0.2, noiseWeight=0.002, varyNoise=0.3,
                     flowDecay=0.005, flowWidth=5.0, noiseScale=2.125, varyNoiseScale=0.5,
                     noiseSpeed=0.00025, varyNoiseSpeed=0.1, target=0.0, varyTarget=1.0)


def make_params(viewSize=(1.0, 1.0), **over) -> Params:
    d = dict(DEFAULT_STATE)
    d.update(over)
    p = Params()
    for k, v in d.items():
        setattr(p, k, float(v))
    p.viewSize[0], p.viewSize[1] = float(viewSize[0]), float(viewSize[1])
    return p


def make_spawn_pixels(spawnSize=(1.0, 1.0), jitter=(0.0, 0.0), speed=1.0, bias=1.0,
                      spawnMatrix=(1, 0, 0, 0, 1, 0, 0, 0, 1), flowDecay=0.005) -> SpawnPixels:
    s = SpawnPixels()
    s.spawnSize[0], s.spawnSize[1] = map(float, spawnSize)
    s.jitter[0], s.jitter[1] = map(float, jitter)
    s.speed, s.bias, s.flowDecay = float(speed), float(bias), float(flowDecay)
    for i, v in enumerate(spawnMatrix):
        s.spawnMatrix[i] = float(v)
    return s


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "tendrils_oracle.c")
    hdr = os.path.join(_HERE, "tendrils_oracle.h")
    stale = (not os.path.exists(_LIB_PATH)) or any(
        os.path.exists(s) and os.path.getmtime(s) > os.path.getmtime(_LIB_PATH) for s in (src, hdr))
    if force or stale:
        subprocess.run(["make", "-C", _HERE, "-s"] + (["-B"] if force else []), check=True)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        L = C.CDLL(_LIB_PATH)
        L.or_sin.restype = L.or_cos.restype = C.c_float
        L.or_sin.argtypes = L.or_cos.argtypes = [C.c_float]
        L.or_random.restype = C.c_float
        L.or_random.argtypes = [C.c_float, C.c_float]
        L.or_snoise3.restype = C.c_float
        L.or_snoise3.argtypes = [C.c_float] * 3
        L.or_integrate.restype = None
        L.or_integrate.argtypes = [C.POINTER(Params), C.c_int, C.c_int, C.c_int, C.c_int,
                                   _fp, _fp, _fp, _fp, C.c_int, C.c_int, C.c_float, C.c_float]
        L.or_vertex_table.restype = None
        L.or_vertex_table.argtypes = [C.c_int, _ip, _ip]
        L.or_column_table.restype = None
        L.or_column_table.argtypes = [C.c_int, _ip]
        L.or_splat.restype = C.c_longlong
        L.or_splat.argtypes = [C.POINTER(Params), C.c_int, C.c_int, C.c_int, C.c_int,
                               _fp, _fp, _fp, C.c_int, C.c_int, C.c_float]
        L.or_flow_vertex.restype = C.c_int
        L.or_flow_vertex.argtypes = [C.POINTER(Params), C.c_int, C.c_int, C.c_int, C.c_int, _fp, _fp, C.c_float, _fp]
        L.or_splat_mt.restype = C.c_longlong
        L.or_splat_mt.argtypes = L.or_splat.argtypes
        L.or_spawn_init.restype = None
        L.or_spawn_init.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, _fp]
        L.or_spawn_ball.restype = None
        L.or_spawn_ball.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, _fp]
        L.or_spawn_pixels_direct.restype = None
        L.or_spawn_pixels_direct.argtypes = [C.POINTER(SpawnPixels), C.c_int, C.c_int, C.c_int, C.c_int,
                                             _fp, C.c_int, C.c_int, C.c_float, _fp]
        L.or_spawn_pixels_sample.restype = None
        L.or_spawn_pixels_sample.argtypes = [C.POINTER(SpawnPixels), C.c_int, C.c_int, C.c_int,
                                             C.c_int, C.c_int, C.c_int, C.c_int, _fp,
                                             _fp, C.c_int, C.c_int, C.c_float, _fp]
        L.or_optical_flow.restype = None
        L.or_optical_flow.argtypes = [_fp, _fp, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float,
                                      C.c_void_p, C.c_void_p, C.c_int, C.c_int, _fp, C.c_int, C.c_int]
        L.or_flow_line_vertex.restype = None
        L.or_flow_line_vertex.argtypes = [C.POINTER(FlowLineUniforms), _fp, _fp, C.c_float, _fp, C.c_float, C.c_float, _fp]
        L.or_flow_line_fragment.restype = None
        L.or_flow_line_fragment.argtypes = [C.c_float, _fp, _fp]
        L.or_flow_line.restype = C.c_longlong
        L.or_flow_line.argtypes = [C.POINTER(FlowLineUniforms), C.c_int, _fp, _fp, _fp, _fp, _fp, _fp, _fp, C.c_int, C.c_int]
        L.or_num_threads.restype = C.c_int
        L.or_set_threads.restype = None
        L.or_set_threads.argtypes = [C.c_int]
        _lib = L
    return _lib


def _p(a: np.ndarray):
    assert a.dtype == np.float32 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(_fp)


def sin(x): return lib().or_sin(float(x))
def cos(x): return lib().or_cos(float(x))
def random(cx, cy): return lib().or_random(float(cx), float(cy))
def snoise3(x, y, z): return lib().or_snoise3(float(x), float(y), float(z))
def num_threads(): return lib().or_num_threads()


def set_threads(n: int):
    """OpenMP threads of the parallel entry points (a launcher such as torchrun exports OMP_NUM_THREADS=1)."""
    lib().or_set_threads(int(n))


def integrate(P, state, targets, flow, time, dt, cols=None):
    """state/targets: [PW,PH,4] x-major; flow: [H,W,4].  Returns the new state."""
    PW, PH = state.shape[:2]
    H, W = flow.shape[:2]
    x0, x1 = cols or (0, PW)
    out = np.array(state, copy=True) if cols is None else np.empty_like(state)
    lib().or_integrate(C.byref(P), PW, PH, x0, x1, _p(state), _p(out), _p(targets), _p(flow),
                       W, H, f32(time), f32(dt))
    return out


def vertex_table(PH):
    row = np.zeros(2 * PH, np.int32)
    cur = np.zeros(2 * PH, np.int32)
    lib().or_vertex_table(PH, row.ctypes.data_as(_ip), cur.ctypes.data_as(_ip))
    return row, cur


def column_table(PW):
    col = np.zeros(PW, np.int32)
    lib().or_column_table(PW, col.ctypes.data_as(_ip))
    return col


def splat(P, cur, prev, flow, time, cols=None, mt=False):
    """Blends in place into flow ([H,W,4]); returns the fragment count."""
    PW, PH = cur.shape[:2]
    H, W = flow.shape[:2]
    x0, x1 = cols or (0, PW)
    fn = lib().or_splat_mt if mt else lib().or_splat
    return fn(C.byref(P), PW, PH, x0, x1, _p(cur), _p(prev), _p(flow), W, H, f32(time))


def flow_vertex(P, cur, prev, i, j, time):
    """(written, [gl_Position.x, gl_Position.y, r, g, b, a]) of vertex (column i, row j) of the flow draw."""
    PW, PH = cur.shape[:2]
    out = np.zeros(6, np.float32)
    ok = lib().or_flow_vertex(C.byref(P), PW, PH, int(i), int(j), _p(cur), _p(prev), f32(time), _p(out))
    return bool(ok), out


def spawn_init(PW, PH, cols=None):
    out = np.empty((PW, PH, 4), np.float32)
    x0, x1 = cols or (0, PW)
    lib().or_spawn_init(PW, PH, x0, x1, _p(out))
    return out


def spawn_ball(PW, PH, radius=1.0, speed=0.0, cols=None):
    out = np.empty((PW, PH, 4), np.float32)
    x0, x1 = cols or (0, PW)
    lib().or_spawn_ball(PW, PH, x0, x1, f32(radius), f32(speed), _p(out))
    return out


def spawn_pixels_direct(S, PW, PH, image, time, cols=None):
    out = np.empty((PW, PH, 4), np.float32)
    IH, IW = image.shape[:2]
    x0, x1 = cols or (0, PW)
    lib().or_spawn_pixels_direct(C.byref(S), PW, PH, x0, x1, _p(image), IW, IH, f32(time), _p(out))
    return out


SAMPLE_VARIANTS = {            # name: (apply, vignette, samples)  -- src/spawn/pixels/*-sample.frag
    "best": (APPLY_COLOR, 1, 6),
    "bright": (APPLY_BRIGHTEST, 0, 6),
    "color": (APPLY_COLOR, 0, 3),
    "data": (APPLY_IDENTITY, 1, 2),
    "flow": (APPLY_FLOW, 0, 5),
}


def spawn_pixels_sample(S, variant, state, image, time, cols=None):
    PW, PH = state.shape[:2]
    IH, IW = image.shape[:2]
    apply, vig, samples = SAMPLE_VARIANTS[variant]
    out = np.empty((PW, PH, 4), np.float32)
    x0, x1 = cols or (0, PW)
    lib().or_spawn_pixels_sample(C.byref(S), apply, vig, samples, PW, PH, x0, x1, _p(state),
                                 _p(image), IW, IH, f32(time), _p(out))
    return out


def optical_flow(flow, view, last, viewSize=(1.0, 1.0), scaleUV=(1.0, -1.0), offset=1.0, lambda_=0.001, speed=1.0,
                 speedLimit=1.0, time=1.0):
    """Blends the optical flow of two RGBA8 frames ([h,w,4] uint8) into flow ([H,W,4] float32), in place."""
    H, W = flow.shape[:2]
    ih, iw = view.shape[:2]
    assert view.dtype == np.uint8 and last.dtype == np.uint8 and view.shape == last.shape
    vs = np.asarray(viewSize, np.float32)
    sc = np.asarray(scaleUV, np.float32)
    view, last = np.ascontiguousarray(view), np.ascontiguousarray(last)
    lib().or_optical_flow(_p(vs), _p(sc), f32(offset), f32(lambda_), f32(speed), f32(speedLimit), f32(time),
                          view.ctypes.data, last.ctypes.data, iw, ih, _p(flow), W, H)
    return flow


def flow_line_uniforms(viewSize=(1.0, 1.0), rad=0.1, speed=3.0, speedLimit=0.01, crestShape=0.6):
    u = FlowLineUniforms()
    u.viewSize[0], u.viewSize[1] = float(viewSize[0]), float(viewSize[1])
    u.rad, u.speed, u.speedLimit, u.crestShape = float(rad), float(speed), float(speedLimit), float(crestShape)
    return u


def flow_line_vertex(U, position, normal, miter, previous, time, dt):
    """src/flow-line/index.vert for one vertex: (gl_Position.xy, values.rgba, crest.xy, sdf)."""
    out = np.zeros(9, np.float32)
    pos, nor, prv = (np.ascontiguousarray(v, np.float32) for v in (position, normal, previous))
    lib().or_flow_line_vertex(C.byref(U), _p(pos), _p(nor), f32(miter), _p(prv), f32(time), f32(dt), _p(out))
    return out


def flow_line_fragment(crestShape, in7):
    """src/flow-line/index.frag for one fragment: in7 = (values.rgba, crest.xy, sdf)."""
    out = np.zeros(4, np.float32)
    a = np.ascontiguousarray(in7, np.float32)
    lib().or_flow_line_fragment(f32(crestShape), _p(a), _p(out))
    return out


def flow_line(U, attributes, flow):
    """Draws the strip into flow ([H,W,4], in place).  attributes: dict of float32 arrays as gl-geometry holds
    them (position [n,2], normal [n,2], miter [n], previous [n,2], time [n], dt [n]).  Returns the fragment count."""
    H, W = flow.shape[:2]
    a = {k: np.ascontiguousarray(attributes[k], np.float32) for k in ("position", "normal", "miter", "previous", "time", "dt")}
    n = a["miter"].shape[0]
    return lib().or_flow_line(C.byref(U), n, _p(a["position"]), _p(a["normal"]), _p(a["miter"]), _p(a["previous"]),
                              _p(a["time"]), _p(a["dt"]), _p(flow), W, H)
