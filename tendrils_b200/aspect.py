"""aspect helpers -- mirror of src/utils/aspect.js (gl-matrix vec2.inverse + vec2.scale).

`out` decides the rounding, as in JS: a plain list keeps doubles (Tendrils.viewSize is a plain
Array, src/index.js:139), a float32 numpy array rounds at every store (PixelSpawner.jitter is a
gl-matrix Float32Array, src/spawn/pixels/index.js:43).
"""
from __future__ import annotations

import numpy as np


def aspect(out, size, scale):
    inv = [1.0 / float(size[0]), 1.0 / float(size[1])]       # vec2.inverse(out, size)
    out[0], out[1] = inv[0], inv[1]
    a, b = float(out[0]), float(out[1])                        # re-read: float32 storage rounds here
    out[0], out[1] = a * scale, b * scale                      # vec2.scale(out, out, scale)
    return out


def containAspect(out, size):
    return aspect(out, size, min(size[0], size[1]))


def coverAspect(out, size):
    return aspect(out, size, max(size[0], size[1]))


def f32vec2():
    return np.zeros(2, dtype=np.float32)
