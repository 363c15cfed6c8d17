"""Timer -- mirror of the reference's src/timer.js (same fields, same tick() arithmetic).

All values are Python floats (IEEE doubles, as in JS); they become float32 only when handed to
the GPU, exactly like `gl.uniform1f`.
"""
from __future__ import annotations

import math
import time as _time


def _now_ms() -> float:
    return _time.time() * 1000.0


class Timer:
    def __init__(self, now=None, since=None):
        self.time = 0.0
        self.since = 0.0
        self.offset = 0.0
        self.rate = 1.0
        self.step = -1.0
        self.dt = 0.0
        self.paused = False
        self.end = -1.0
        self.loop = False
        self.reset(now, since)

    def now(self, now=None):                      # src/timer.js:20-22
        if now is None:
            now = _now_ms()
        return (now - self.offset) * self.rate

    def tick(self, now=None):                     # src/timer.js:24-60
        time = self.time
        dt = 0.0
        if self.step >= 0:
            dt = self.step * self.rate
            time += dt
        else:
            past = time
            time = self.now(now)
            dt = time - past
        if self.paused:
            self.offset += dt
            dt = 0.0
        elif self.end < 0:
            self.time = time
        elif self.loop:
            self.time = math.fmod(time, self.end)  # JS % is fmod
        else:
            self.time = (min if self.rate > 0 else max)(time, self.end)
            if self.time != time:
                self.paused = True
        self.dt = dt
        return self

    def seek(self, to):
        self.offset = -to
        return self

    def scrub(self, by):
        self.offset -= by
        return self

    def reset(self, now=None, since=None):        # src/timer.js:74-79
        if now is None:
            now = _now_ms()
        if since is None:
            since = now
        self.since = self.offset = since
        self.time = self.now(now)
        return self
