"""src/spawn/ball/index.js:5-16 -- the deterministic (hash of gl_FragCoord) ball spawner.
The reference's CPU ball spawner (spawn/ball/cpu.js) uses unseeded Math.random and is not mirrored."""
from __future__ import annotations

from ..tendrils import Shader
from . import init

frag = Shader("spawn-ball")


def defaults():
    return {"shader": frag, "uniforms": {"radius": 1, "speed": 0}}


def spawnBall(gl, options=None):
    params = defaults()
    params.update(options or {})
    return init.spawner(gl, params)
