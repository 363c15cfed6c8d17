"""src/spawn/init/index.js:11-26 -- a spawner object {gl, uniforms, shader, spawn(tendrils, ...rest)}."""
from __future__ import annotations

from ..tendrils import Shader

frag = Shader("spawn-init")


def defaults():
    return {"shader": frag, "uniforms": None}


class _Spawner:
    def __init__(self, gl, params):
        self.gl = gl
        self.uniforms = params["uniforms"]
        self.shader = params["shader"]

    def spawn(self, tendrils, *rest):
        tendrils.spawnShader(self.shader, self.uniforms, *rest)


def spawner(gl, options=None):
    params = defaults()
    params.update(options or {})
    return _Spawner(gl, params)
