"""OpticalFlow -- mirror of src/optical-flow/index.js: two RGBA8 frame buffers (current, previous) and the
uniforms of the gradient optical-flow shader that the app draws INTO the flow FBO after the particle splat
(src/demo.main.js:1131-1159)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _native as N


def defaults():                                                          # src/optical-flow/index.js:16-31
    return {"uniforms": {"viewSize": [1, 1], "scaleUV": [1, -1], "offset": 1, "lambda": 0.001, "speed": 1,
                         "speedLimit": 1, "time": 1}}


class OpticalFlow:
    def __init__(self, gl, options=None, uniforms=None):
        self.gl = gl
        self.buffers = [np.zeros((1, 1, 4), np.uint8), np.zeros((1, 1, 4), np.uint8)]   # FBO(gl, [1,1]) x 2, RGBA8
        self.uniforms = defaults()["uniforms"]
        self.uniforms.update(uniforms or {})
        self._bound = dict(self.uniforms)

    def update(self, uniforms=None):                                     # :50-58
        self._bound = dict(self.uniforms)
        self._bound.update(uniforms or {})
        return self

    def render(self, tendrils):
        """`screen.render()` with the flow FBO bound (src/demo.main.js:1107-1109,1155)."""
        u = self._bound
        p = N.TbOpticalFlowParams()
        p.viewSize[0], p.viewSize[1] = float(u["viewSize"][0]), float(u["viewSize"][1])
        p.scaleUV[0], p.scaleUV[1] = float(u["scaleUV"][0]), float(u["scaleUV"][1])
        p.offset, p.lambda_, p.speed = float(u["offset"]), float(u["lambda"]), float(u["speed"])
        p.speedLimit, p.time = float(u["speedLimit"]), float(u["time"])
        view, last = self.buffers
        if N.is_device_array(view) or N.is_device_array(last):            # frames already on the device stay there
            import torch
            dev = view.device if N.is_device_array(view) else last.device
            view, last = (b if N.is_device_array(b) else torch.as_tensor(np.ascontiguousarray(b, dtype=np.uint8), device=dev)
                          for b in (view, last))
        else:
            view, last = (np.ascontiguousarray(b, dtype=np.uint8) for b in (view, last))
        if tuple(view.shape) != tuple(last.shape):
            raise N.TendrilsError("tendrils-b200: optical-flow buffers differ in shape (call resize)")
        ctx = tendrils.particles._ctx
        N.wait_for_producer(ctx, view)                 # device frames: ordered after the stream that wrote them
        N.check(ctx, N.load().tb_optical_flow(ctx, C.byref(p), N.array_pointer(view, "uint8"), N.array_pointer(last, "uint8"),
                                              view.shape[1], view.shape[0]))
        self._keep = (view, last)          # device frames are read in stream order: keep them alive until the next render
        return self

    def step(self):                                                      # :60-62 utils.step(buffers)
        self.buffers.insert(0, self.buffers.pop())

    def setPixels(self, pixels):                                         # :64-66
        """`pixels`: an [h,w,4] uint8 array, or a CUDA tensor of that shape (a decoded frame already on the device)."""
        self.buffers[0] = pixels if N.is_device_array(pixels) else np.ascontiguousarray(pixels, dtype=np.uint8)
        return self.buffers[0]

    def resize(self, size):                                              # :68-70 (gl-fbo reshape zeroes)
        w, h = int(size[0]), int(size[1])
        self.buffers = [b if tuple(b.shape[:2]) == (h, w) else np.zeros((h, w, 4), np.uint8) for b in self.buffers]
