"""FlowLine / FlowLines / Line -- mirror of src/flow-line/index.js, src/flow-line/multi.js and src/geom/line/index.js:
pointer paths drawn as mitred triangle strips INTO the flow FBO (src/demo.main.js:378-395, 1107-1121).

The geometry is host work in the reference too (JS doubles, stored to Float32Array attributes); it is restated
here operation for operation, including its third-party pieces, whose text is in docs/js/demo.js.map:
polyline-normals@2.0.2, polyline-miter-util@1.0.1 and the gl-vec2 helpers they call.  The shaders and the
rasterisation run on the GPU behind `tb_flow_line` (include/tendrils_b200.h).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _native as N

_f64 = np.float64


# ---- gl-vec2 / polyline-miter-util, in doubles like the JS ------------------------------------------------------
def _normalize(v):
    """gl-vec2/normalize: a zero vector is left untouched."""
    x, y = v
    ln = x * x + y * y
    if ln > 0:
        ln = 1 / np.sqrt(_f64(ln))
        return [float(x * ln), float(y * ln)]
    return [x, y]


def _direction(a, b):                                                   # polyline-miter-util direction(out, a, b)
    return _normalize([a[0] - b[0], a[1] - b[1]])


def _normal(d):                                                         # polyline-miter-util normal(out, dir)
    return [-d[1], d[0]]


def _compute_miter(line_a, line_b, half_thick):                          # polyline-miter-util computeMiter
    tangent = _normalize([line_a[0] + line_b[0], line_a[1] + line_b[1]])
    miter = [-tangent[1], tangent[0]]
    tmp = [-line_a[1], line_a[0]]
    with np.errstate(all="ignore"):                                     # JS: x/0 is +-Infinity or NaN, never an exception
        length = float(_f64(half_thick) / _f64(miter[0] * tmp[0] + miter[1] * tmp[1]))
    return miter, length


def polyline_normals(points, closed=False):
    """polyline-normals@2.0.2 index.js: [[normal, miterLength], ...] for a 2-D polyline."""
    cur_normal = None
    out = []
    points = [list(map(float, p)) for p in points]
    if closed:
        points = points + [points[0]]
    total = len(points)
    for i in range(1, total):
        last, cur = points[i - 1], points[i]
        nxt = points[i + 1] if i < len(points) - 1 else None
        line_a = _direction(cur, last)
        if cur_normal is None:
            cur_normal = _normal(line_a)
        if i == 1:                                                      # add initial normals
            out.append([list(cur_normal), 1.0])
        if nxt is None:                                                 # no miter, simple segment
            cur_normal = _normal(line_a)
            out.append([list(cur_normal), 1.0])
        else:                                                           # miter with last
            line_b = _direction(nxt, cur)
            miter, miter_len = _compute_miter(line_a, line_b, 1)
            out.append([list(miter), miter_len])
    if len(points) > 2 and closed:                                      # clean up the last normal of a closed loop
        last2, cur2, next2 = points[total - 2], points[0], points[1]
        line_a = _direction(cur2, last2)
        line_b = _direction(next2, cur2)
        miter, miter_len2 = _compute_miter(line_a, line_b, 1)
        out[0][0] = list(miter)
        out[total - 1][0] = list(miter)
        out[0][1] = miter_len2
        out[total - 1][1] = miter_len2
        out.pop()
    return out


# ---- src/geom/line/index.js -------------------------------------------------------------------------------------
def defaults():                                                         # src/geom/line/index.js:14-27
    return {"shader": None, "uniforms": {"color": [1, 1, 1, 1], "rad": 0.1, "viewSize": [1, 1]}, "attributes": None,
            "vertNum": 2, "vertSize": 2, "path": [], "closed": False}


class Line:
    """Two vertices per path point with attributes position / normal / miter (+ any the owner adds)."""

    def __init__(self, gl, options=None):
        params = defaults()
        params.update(options or {})
        self.gl = gl
        self.uniforms = params["uniforms"]
        self.shader = params["shader"]
        self.vertNum, self.vertSize = params["vertNum"], params["vertSize"]
        if (self.vertNum, self.vertSize) != (2, 2):
            raise N.TendrilsError("tendrils-b200: Line supports vertNum = vertSize = 2 (what FlowLine uses)")
        self.path = params["path"] if params["path"] is not None else []
        self.closed = params["closed"]
        self.attributes = {"position": {"data": None, "getSize": lambda line: line.vertSize},
                           "normal": {"data": None, "getSize": lambda line: line.vertSize},
                           "miter": {"data": None, "size": 1}}
        self.attributes.update(params["attributes"] or {})
        self.drawnPath = self.drawnNormals = None

    def update(self, setAttributes=None):                               # :73-117
        setAttributes = setAttributes or self.setAttributes
        self.drawnPath = self.path
        self.drawnNormals = polyline_normals(self.drawnPath, self.closed)
        if self.closed and len(self.path):
            self.drawnPath = list(self.drawnPath) + [self.drawnPath[0]]
            self.drawnNormals.append(self.drawnNormals[0])
        self.initAttributes()
        values, index = {}, {}
        for p in range(len(self.drawnNormals)):
            point_normal = self.drawnNormals[p]
            values["point"], values["normal"], values["miter"] = self.drawnPath[p], point_normal[0], point_normal[1]
            index["path"], index["point"] = p, p * self.vertNum
            for v in range(self.vertNum):
                index["vert"], index["data"] = v, index["point"] + v
                setAttributes(values, index, self.attributes, self)
        return self

    def initAttributes(self):                                           # :131-149
        num = len(self.drawnPath) * self.vertNum
        for attribute in self.attributes.values():
            if attribute.get("getSize"):
                attribute["size"] = attribute["getSize"](self)
            length = num * attribute["size"]
            if attribute.get("data") is None or attribute["data"].shape[0] != length:
                attribute["data"] = np.zeros(length, np.float32)       # new Float32Array(length)
        return self

    def setAttributes(self, values, index, attributes, line=None):      # :151-160
        i = index["data"]
        s = attributes["position"]["size"]
        attributes["position"]["data"][i * s:i * s + s] = values["point"]
        s = attributes["normal"]["size"]
        attributes["normal"]["data"][i * s:i * s + s] = values["normal"]
        attributes["miter"]["data"][i] = values["miter"] * (((i % 2) * 2) - 1)     # flip odd miters

    def vertex_count(self):
        """What gl-geometry draws: one vertex per `miter` entry filled by update()."""
        return 0 if self.drawnNormals is None else len(self.drawnNormals) * self.vertNum


# ---- src/flow-line/index.js --------------------------------------------------------------------------------------
def _wrap_index(i, n):
    return n + i if i < 0 else i % n


class FlowLine:
    def __init__(self, gl, options=None):
        options = dict(options or {})
        self.times = options.pop("times", None) or []
        uniforms = defaults()["uniforms"]
        uniforms.update({"speed": 3, "speedLimit": 0.01, "rad": 0.1, "crestShape": 0.6})
        params = {"shader": "flow-line", "uniforms": uniforms,
                  "attributes": {"previous": {"data": None, "getSize": lambda line: line.vertSize},
                                 "time": {"data": None, "size": 1}, "dt": {"data": None, "size": 1}}}
        params.update(options)
        self.line = Line(gl, params)

    def update(self, setAttributes=None):                               # :39-48
        setAttributes = setAttributes or self.setAttributes
        closed_times = self.line.closed and len(self.line.path)
        drawn_times = (list(self.times) + [self.times[0]]) if closed_times else self.times
        self.line.update(lambda *rest: setAttributes(drawn_times, *rest))
        return self

    def setAttributes(self, times, values, index, attributes, line):    # :56-69
        line.setAttributes(values, index, attributes, line)
        n = len(line.path)
        prev = _wrap_index(index["path"] - 1, n) if line.closed else max(0, index["path"] - 1)
        s = attributes["previous"]["size"]
        i = index["data"]
        attributes["previous"]["data"][i * s:i * s + s] = line.path[prev]
        time = times[index["path"]]
        attributes["time"]["data"][i] = time
        attributes["dt"]["data"][i] = time - times[prev]

    def draw(self, tendrils):
        """`line.draw()` with the flow FBO bound (src/demo.main.js:1107-1121): TRIANGLE_STRIP of all vertices."""
        line = self.line
        n = line.vertex_count()
        if len(line.path) == 0 or n == 0:
            return self
        if line.shader != "flow-line":
            raise N.TendrilsError("tendrils-b200: custom line shaders are not supported")
        u = line.uniforms
        p = N.TbFlowLineParams()
        p.viewSize[0], p.viewSize[1] = float(u["viewSize"][0]), float(u["viewSize"][1])
        p.rad, p.speed, p.speedLimit, p.crestShape = float(u["rad"]), float(u["speed"]), float(u["speedLimit"]), float(u["crestShape"])
        a = {k: np.ascontiguousarray(v["data"], dtype=np.float32) for k, v in line.attributes.items()}
        ptr = lambda name: a[name].ctypes.data_as(N._fp)
        ctx = tendrils.particles._ctx
        N.check(ctx, N.load().tb_flow_line(ctx, C.byref(p), n, ptr("position"), ptr("normal"), ptr("miter"), ptr("previous"),
                                           ptr("time"), ptr("dt")))
        return self

    def add(self, time, point):                                         # :71-76
        self.times.append(time)
        self.line.path.append(list(point))
        return self

    def insert(self, time, point):                                      # :78-85
        i = self.findIndex(time)
        self.times.insert(i, time)
        self.line.path.insert(i, list(point))
        return self

    def at(self, index, out=None):                                      # :87-92
        out = {} if out is None else out
        out["time"], out["point"] = self.times[index], self.line.path[index]
        return out

    def findIndex(self, time):                                          # :94-98
        for i, other in enumerate(self.times):
            if other > time:
                return i
        return len(self.times)

    def trim(self, ago, now):                                           # :107-117
        oldest = now - ago
        while self.times and self.times[0] < oldest:
            self.times.pop(0)
            self.line.path.pop(0)
        return self.length

    @property
    def length(self):
        return len(self.times)


class FlowLines:                                                         # src/flow-line/multi.js
    def __init__(self, gl):
        self.gl = gl
        self.active = {}

    def get(self, id, options=None):
        if id not in self.active:
            self.active[id] = FlowLine(self.gl, options)
        return self.active[id]

    def trim(self, *times):
        remaining = 0
        for id in list(self.active):
            if self.active[id].trim(*times) == 0:
                del self.active[id]
            else:
                remaining += 1
        return remaining
