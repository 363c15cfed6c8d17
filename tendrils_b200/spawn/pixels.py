"""src/spawn/pixels/index.js:25-67 -- PixelSpawner: wraps an image (or the flow grid, or the
particle texture) and a built-in pixel spawn shader for Tendrils.spawnShader."""
from __future__ import annotations

import numpy as np

from .. import _native as N
from ..aspect import aspect, f32vec2
from ..tendrils import Shader

# the reference's fragment shaders (src/spawn/pixels/*.frag)
pixelsFrag = Shader("spawn-pixels", N.TB_SPAWN_DIRECT)
bestSampleFrag = Shader("spawn-pixels", N.TB_SPAWN_BEST_SAMPLE)
brightSampleFrag = Shader("spawn-pixels", N.TB_SPAWN_BRIGHT_SAMPLE)
colorSampleFrag = Shader("spawn-pixels", N.TB_SPAWN_COLOR_SAMPLE)
dataSampleFrag = Shader("spawn-pixels", N.TB_SPAWN_DATA_SAMPLE)
flowSampleFrag = Shader("spawn-pixels", N.TB_SPAWN_FLOW_SAMPLE)


def defaults():
    return {"shader": pixelsFrag, "buffer": None, "spawnSize": [1, 1], "jitterRad": 2, "speed": 1, "bias": 1}


def mat3_identity():
    return np.array([1, 0, 0, 0, 1, 0, 0, 0, 1], dtype=np.float32)     # gl-matrix mat3.create()


def mat3_scale(m, v):
    """gl-matrix mat3.scale(out, a, v): scales the first two columns (Float32Array storage)."""
    out = np.array(m, dtype=np.float32)
    out[0:3] = (out[0:3].astype(np.float64) * float(v[0])).astype(np.float32)
    out[3:6] = (out[3:6].astype(np.float64) * float(v[1])).astype(np.float32)
    return out


class PixelSpawner:
    def __init__(self, gl, options=None):
        params = defaults()
        params.update(options or {})
        self.gl = gl
        self.shader = params["shader"]
        self.buffer = params["buffer"]          # [h,w,4] float image, tendrils.flow or particles.buffers[0]
        self.speed = params["speed"]
        self.bias = params["bias"]
        self.jitterRad = params["jitterRad"]
        self.jitter = f32vec2()                 # vec2.create() -> Float32Array
        self.spawnSize = params["spawnSize"]
        self.spawnMatrix = mat3_identity()

    def update(self, uniforms):                                          # :49-58
        uniforms.update({
            "spawnData": self.buffer,
            "spawnSize": self.spawnSize,
            "spawnMatrix": self.spawnMatrix,
            "speed": self.speed,
            "jitter": aspect(self.jitter, uniforms["viewRes"], self.jitterRad),
            "bias": self.bias,
        })
        return uniforms

    def spawn(self, tendrils, update=None, *rest):                       # :61-63
        return tendrils.spawnShader(self.shader, update or self.update, *rest)

    def setPixels(self, pixels):                                         # :65-67
        from .. import _native as N
        self.buffer = pixels if N.is_device_array(pixels) else np.ascontiguousarray(pixels, dtype=np.float32)
        return self.buffer
