"""Tendrils / Particles -- host-side mirror of the reference's JS facade for the hot path.

Same names, argument meaning and call order as `src/index.js` (class Tendrils) and
`src/particles.js` (class Particles) of keeffEoghan/tendrils, with the WebGL command stream
replaced by the C ABI of include/tendrils_b200.h.  Only the step, the flow half of draw() and
the spawn passes exist here; view drawing stays in WebGL and is out of scope.

The `gl` argument of the reference (a WebGLRenderingContext) is replaced by a `Device`: the
CUDA device ordinal plus the size of the drawing buffer the flow FBO follows.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import _native as N
from .aspect import coverAspect
from .timer import Timer

INERT = -1000000.0                                   # src/const/inert.js:2


class Device:
    """Stand-in for the WebGL context: where to run and how big the drawing buffer is.

    rank/world_size describe a column-sharded multi-GPU run (one process per GPU); group is the
    torch.distributed process group used for the ordered flow exchange."""

    def __init__(self, width=1, height=1, device=0, rank=0, world_size=1, group=None, ring=None):
        self.drawingBufferWidth = int(width)
        self.drawingBufferHeight = int(height)
        self.device = int(device)
        self.rank = int(rank)
        self.world_size = int(world_size)
        self.group = group
        # how the ordered flow blend is shared between ranks:
        #   "owners" = over CUDA-IPC peer memory (default): the fragment bins are owned round-robin by the ranks, every rank
        #              rasterises straight into the owners' bins over NVLink, folds its own bins and stores the finished
        #              texels into every grid; three all-rank barriers per draw, all inside the CUDA stream
        #   "dist"   = the grid travels rank 0 -> 1 -> ... over torch.distributed send/recv/broadcast, every rank folding
        #              its own bins onto what it received (serial; needs no peer mapping)
        self.ring = ring or os.environ.get("TB_RING") or "owners"


class Shader:
    """A built-in shader of the path.  The reference accepts arbitrary GLSL here
    (src/index.js:69,112; src/spawn/init/index.js:18-20); this implementation only the
    built-ins, and raises for anything else (no fallback by design)."""

    def __init__(self, kind, variant=None):
        self.kind = kind            # 'logic' | 'flow' | 'spawn-init' | 'spawn-ball' | 'spawn-pixels'
        self.variant = variant
        self.uniforms = {}

    def __repr__(self):
        return f"Shader({self.kind!r}, {self.variant!r})"


logicFrag = Shader("logic")
flowShader = Shader("flow")


def defaults():
    """src/index.js:28-75 (step-relevant part; display-only options are accepted and ignored)."""
    timer = Timer()
    timer.step = 1000 / 60
    return {
        "state": {
            "rootNum": 2 ** 9,
            "autoClearView": False, "autoFade": True,
            "damping": 0.043, "speedLimit": 0.01,
            "forceWeight": 0.016, "varyForce": -0.1,
            "flowWeight": 1, "varyFlow": 0.2,
            "noiseWeight": 0.002, "varyNoise": 0.3,
            "flowDecay": 0.005, "flowWidth": 5,
            "noiseScale": 2.125, "varyNoiseScale": 0.5,
            "noiseSpeed": 0.00025, "varyNoiseSpeed": 0.1,
            "target": 0, "varyTarget": 1,
            "lineWidth": 1, "speedAlpha": 0.000001, "colorMapAlpha": 0.4,
            "baseColor": [1, 1, 1, 0.5], "flowColor": [1, 1, 1, 0.04],
            "fadeColor": [0.1333, 0.1333, 0.1333, 0],
        },
        "timer": timer,
        "numBuffers": 0,
        "logicShader": None,
        "flowShader": flowShader,
    }


def shard_columns(width, rank, world_size):
    """Contiguous column block of `rank`: contiguous in draw order p = x*PH + y, which the
    ordered flow blend needs."""
    base, rem = divmod(int(width), int(world_size))
    c0 = rank * base + min(rank, rem)
    return c0, c0 + base + (1 if rank < rem else 0)


def _state_struct(uniforms) -> N.TbState:
    s = N.TbState()
    for name in N.STATE_FIELDS:
        setattr(s, name, float(uniforms[name]))        # JS double -> f32, as gl.uniform1f
    vs = uniforms.get("viewSize", (1.0, 1.0))
    s.viewSize[0], s.viewSize[1] = float(vs[0]), float(vs[1])
    return s


class _Buffer:
    """A device-resident RGBA32F texture of the path (stands in for a gl-fbo)."""

    def __init__(self, owner, which):
        self._owner, self._which = owner, which

    def _ctx(self):
        return self._owner._ctx

    def download(self, out=None) -> np.ndarray:
        """Reads the texture back (gl.readPixels).  `out`: an existing C-contiguous float32 array of the texture's
        shape to read into -- e.g. a view of pinned host memory, which makes the copy a single DMA."""
        p = self._owner
        if self._which == N.TB_BUF_FLOW:
            shape = (p.flow_shape[1], p.flow_shape[0], 4)
        else:
            shape = (p.col1 - p.col0, p.shape[1], 4)
        if out is None:
            out = np.empty(shape, np.float32)
        elif not (isinstance(out, np.ndarray) and out.dtype == np.float32 and out.shape == shape
                  and out.flags["C_CONTIGUOUS"] and out.flags["WRITEABLE"]):
            raise N.TendrilsError(f"tendrils-b200: download(out=...) needs a writable C-contiguous float32 array of shape {shape}")
        N.check(self._ctx(), N.load().tb_download(self._ctx(), self._which, out.ctypes.data_as(N._fp), out.size))
        return out

    def upload(self, data):
        a = np.ascontiguousarray(data, dtype=np.float32)
        N.check(self._ctx(), N.load().tb_upload(self._ctx(), self._which, a.ctypes.data_as(N._fp), a.size))

    def device_ptr(self):
        ptr, n = C.c_void_p(), C.c_int64()
        N.check(self._ctx(), N.load().tb_device_ptr(self._ctx(), self._which, C.byref(ptr), C.byref(n)))
        return ptr.value, n.value


class _FlowBuffer(_Buffer):
    """tendrils.flow: assigning .shape reallocates and zeroes, like gl-fbo (src/index.js:405)."""

    @property
    def shape(self):
        return list(self._owner.flow_shape)

    @shape.setter
    def shape(self, wh):
        self._owner._resize_flow(int(wh[0]), int(wh[1]))


class Particles:
    """src/particles.js:43-196 -- ping-pong state buffers, logic pass, line draw, CPU spawn."""

    def __init__(self, gl, options):
        params = {"shape": [64, 64], "geomShape": None, "logic": None, "logicFrag": None, "render": None}
        params.update(options or {})
        self.gl = gl
        self.shape = [int(params["shape"][0]), int(params["shape"][1])]
        self.geomShape = list(params["geomShape"] or self.shape)
        if self.geomShape != [self.shape[0], self.shape[1] * 2]:
            raise N.TendrilsError("tendrils-b200: geomShape must be (w, 2h) as Tendrils.setupParticles sets it")
        self.logic = params["logic"] or params["logicFrag"]
        self.render = params["render"]
        self.col0, self.col1 = shard_columns(self.shape[0], gl.rank, gl.world_size)
        self.flow_shape = [1, 1]
        self._tiles_ready = False
        self._L = N.load()
        cfg = N.TbConfig(self.shape[0], self.shape[1], self.col0, self.col1, 1, 1, gl.device, 0)
        ctx = C.c_void_p()
        N.check(None, self._L.tb_create(C.byref(cfg), C.byref(ctx)))
        self._ctx = ctx
        self.buffers = []
        # CPU mirror, x-major [w,h,4] (src/particles.js:76-78)
        self.pixels = np.zeros((self.col1 - self.col0, self.shape[1], 4), np.float32)
        self.flow = _FlowBuffer(self, N.TB_BUF_FLOW)
        self.targets = _Buffer(self, N.TB_BUF_TARGETS)

    def setup(self, numBuffers=1):                                      # src/particles.js:81-92
        if numBuffers != 2:
            raise N.TendrilsError("tendrils-b200: the step needs exactly 2 state buffers (ping-pong)")
        self.buffers = [_Buffer(self, N.TB_BUF_CURRENT), _Buffer(self, N.TB_BUF_PREVIOUS)]

    def spawn(self, map, pixels=None, offset=(0, 0)):                   # src/particles.js:94-117
        """CPU spawn: map(data, x, y) per texel, x outer / y inner, uploaded to ALL buffers."""
        if tuple(offset) != (0, 0):
            raise N.TendrilsError("tendrils-b200: spawn offset is not supported")
        pixels = self.pixels if pixels is None else pixels
        data = np.zeros(4, np.float32)
        for xl in range(pixels.shape[0]):
            for y in range(pixels.shape[1]):
                data[:] = 0
                map(data, self.col0 + xl, y)
                pixels[xl, y] = data
        for b in self.buffers:
            b.upload(pixels)

    def _resize_flow(self, w, h):
        N.check(self._ctx, self._L.tb_resize_flow(self._ctx, w, h))
        self.flow_shape = [w, h]
        self._tiles_ready = False         # new allocation: the IPC handles must be exchanged again

    def step(self, update, buffer=None):                                # src/particles.js:123-145
        """Runs `self.logic` once over the state texture.  buffer=None rotates the ping-pong
        pair and writes buffers[0]; an explicit buffer (tendrils.targets) is written in place
        with no rotation."""
        u = dict(update)
        u["dataRes"], u["geomRes"] = self.shape, self.geomShape
        L, ctx, sh = self._L, self._ctx, self.logic
        if buffer is not None and buffer is not self.targets:
            raise N.TendrilsError("tendrils-b200: spawn target must be None or tendrils.targets")
        target = N.TB_TARGET_STATE if buffer is None else N.TB_TARGET_TARGETS
        if not isinstance(sh, Shader):
            raise N.TendrilsError("tendrils-b200: custom logic/spawn shaders are not supported")
        if sh.kind == "logic":
            if buffer is not None:
                raise N.TendrilsError("tendrils-b200: the logic pass cannot target an explicit buffer")
            st = _state_struct(u)
            N.check(ctx, L.tb_set_state(ctx, C.byref(st)))
            N.check(ctx, L.tb_step(ctx, float(u["time"]), float(u["dt"])))
        elif sh.kind == "spawn-init":
            N.check(ctx, L.tb_spawn_init(ctx, target))
        elif sh.kind == "spawn-ball":
            N.check(ctx, L.tb_spawn_ball(ctx, float(u["radius"]), float(u["speed"]), target))
        elif sh.kind == "spawn-pixels":
            st = _state_struct(u)
            N.check(ctx, L.tb_set_state(ctx, C.byref(st)))
            ps = N.TbPixelSpawner()
            ps.spawnSize[0], ps.spawnSize[1] = float(u["spawnSize"][0]), float(u["spawnSize"][1])
            ps.jitter[0], ps.jitter[1] = float(u["jitter"][0]), float(u["jitter"][1])
            ps.speed, ps.bias = float(u["speed"]), float(u["bias"])
            for i in range(9):
                ps.spawnMatrix[i] = float(u["spawnMatrix"][i])
            src = u["spawnData"]
            if src is self.flow:
                source = N.TB_SOURCE_FLOW
            elif self.buffers and src is self.buffers[0]:
                source = N.TB_SOURCE_PARTICLES
            else:
                img = src if N.is_device_array(src) else np.ascontiguousarray(src, dtype=np.float32)
                if img.ndim != 3 or img.shape[2] != 4:
                    raise N.TendrilsError("tendrils-b200: spawnData must be an [h,w,4] float image")
                N.wait_for_producer(ctx, img)          # a device image: ordered after the stream that wrote it
                ptr = C.cast(C.c_void_p(N.array_pointer(img, "float32")), N._fp)
                N.check(ctx, L.tb_set_spawn_image(ctx, ptr, img.shape[1], img.shape[0]))
                self._spawn_image = img       # a device image is copied in stream order: keep it alive
                source = N.TB_SOURCE_IMAGE
            N.check(ctx, L.tb_spawn_pixels(ctx, C.byref(ps), sh.variant, source, float(u["time"]), target))
        else:
            raise N.TendrilsError(f"tendrils-b200: shader {sh!r} cannot run as a logic pass")

    def step_streamed(self, update, host_in, host_out, chunks=16):
        """The logic pass for callers that keep the state on the host: `host_in` ([columns, rows, 4] float32, ideally pinned) is
        uploaded, stepped and read back into `host_out` chunk by chunk, PCIe busy in both directions (tb_step_streamed).
        Asynchronous: `host_out` is complete after sync()."""
        u = dict(update)
        shape = (self.col1 - self.col0, self.shape[1], 4)
        for a in (host_in, host_out):
            if not (isinstance(a, np.ndarray) and a.dtype == np.float32 and a.shape == shape and a.flags["C_CONTIGUOUS"]):
                raise N.TendrilsError(f"tendrils-b200: step_streamed needs C-contiguous float32 arrays of shape {shape}")
        st = _state_struct(u)
        N.check(self._ctx, self._L.tb_set_state(self._ctx, C.byref(st)))
        N.check(self._ctx, self._L.tb_step_streamed(self._ctx, float(u["time"]), float(u["dt"]), host_in.ctypes.data_as(N._fp),
                                                    host_out.ctypes.data_as(N._fp), int(chunks)))

    def draw(self, update, mode="LINES"):                               # src/particles.js:147-158
        """Line draw of all particles with `self.render`; only the flow shader is in scope."""
        if mode != "LINES" or not (isinstance(self.render, Shader) and self.render.kind == "flow"):
            raise N.TendrilsError("tendrils-b200: only the flow-shader GL_LINES draw is in scope")
        u = dict(update)
        st = _state_struct(u)
        L, ctx, gl = self._L, self._ctx, self.gl
        N.check(ctx, L.tb_set_state(ctx, C.byref(st)))
        if gl.world_size == 1:
            N.check(ctx, L.tb_splat_flow(ctx, float(u["time"])))
            return
        if gl.ring == "owners":
            self._ensure_owners()
            N.check(ctx, L.tb_splat_flow_owners(ctx, float(u["time"])))
            return
        N.check(ctx, L.tb_splat_collect(ctx, float(u["time"])))
        from .multi_gpu import ordered_ring_fold
        ordered_ring_fold(gl.rank, gl.world_size, gl.group,
                          fold=lambda: N.check(ctx, L.tb_splat_fold(ctx)),
                          flow_tensor=self._flow_tensor, stream=self.stream_handle())

    def _ensure_owners(self):
        """Exchange CUDA IPC handles of (fragment bins, flow grid, totals table, flags) once per flow-grid allocation and map
        every rank's.  The bin array is fixed at `TB_OWNERS_RESERVE` (default 8) fragments per local particle while mapped.
        torch.distributed is only the courier of the handle blobs."""
        if self._tiles_ready:
            return
        from .multi_gpu import gather_handles
        L, ctx, gl = self._L, self._ctx, self.gl
        per = float(os.environ.get("TB_OWNERS_RESERVE", "8"))
        reserve = min(int(per * (self.col1 - self.col0) * self.shape[1]) + (1 << 16), (1 << 31) - 1)
        nbytes = L.tb_owners_handle_bytes()
        mine = (C.c_ubyte * nbytes)()
        N.check(ctx, L.tb_owners_export(ctx, reserve, mine, nbytes))
        blobs = gather_handles(bytes(mine), gl.world_size, gl.group, gl.device)
        buf = (C.c_ubyte * (nbytes * gl.world_size)).from_buffer_copy(b"".join(blobs))
        N.check(ctx, L.tb_owners_connect(ctx, gl.rank, gl.world_size, buf, nbytes * gl.world_size))
        self._tiles_ready = True

    # -- plumbing ------------------------------------------------------------------------
    def stream_handle(self) -> int:
        s = C.c_void_p()
        N.check(self._ctx, self._L.tb_stream(self._ctx, C.byref(s)))
        return s.value or 0

    def _flow_tensor(self):
        from .multi_gpu import wrap_device_buffer
        ptr, n = self.flow.device_ptr()
        return wrap_device_buffer(ptr, n, self.gl.device)

    def sync(self):
        N.check(self._ctx, self._L.tb_sync(self._ctx))

    def set_overlap(self, on: bool):
        """scheduling only: run the next step's noise under the flow splat (default) or fuse the whole logic pass"""
        N.check(self._ctx, self._L.tb_set_overlap(self._ctx, 1 if on else 0))

    def stats(self):
        a, b = C.c_int64(), C.c_int64()
        N.check(self._ctx, self._L.tb_stats(self._ctx, C.byref(a), C.byref(b)))
        return {"kernel_launches": a.value, "last_fragments": b.value}

    def segment_stats(self):
        """Diagnostics of the last flow splat: bins folded in segments, segments, fragments left on record for the join."""
        nb, ns, nr = C.c_int32(), C.c_int32(), C.c_int64()
        N.check(self._ctx, self._L.tb_debug_segments(self._ctx, C.byref(nb), C.byref(ns), C.byref(nr)))
        return {"bins": nb.value, "segments": ns.value, "records": nr.value}

    def timing(self, reset=False):
        """Summed CUDA-event time of the integrate launches, the flow splats and the side-stream noise
        launches since the last reset."""
        ni, ns, nn = C.c_int64(), C.c_int64(), C.c_int64()
        mi, ms, mn = C.c_float(), C.c_float(), C.c_float()
        N.check(self._ctx, self._L.tb_timing(self._ctx, int(reset), C.byref(ni), C.byref(mi), C.byref(ns), C.byref(ms),
                                             C.byref(nn), C.byref(mn)))
        return {"n_integrate": ni.value, "integrate_ms": mi.value, "n_splat": ns.value, "splat_ms": ms.value,
                "n_noise": nn.value, "noise_ms": mn.value}

    def dispose(self):                                                  # src/particles.js:168-169 (@todo there)
        if getattr(self, "_ctx", None):
            self._L.tb_destroy(self._ctx)
            self._ctx = None

    def __del__(self):
        try:
            self.dispose()
        except Exception:
            pass

    @staticmethod
    def applyUpdate(state, update):                                     # src/particles.js:192-195
        if callable(update):
            return update(state)
        if update:
            state.update(update)
        return state


def initSpawner(data, x=0, y=0):                                        # src/spawn/init/cpu.js:3-8
    data[0] = data[1] = INERT
    data[2] = data[3] = 0
    return data


class Tendrils:
    """src/index.js:83-457, hot-path methods only."""

    def __init__(self, gl, options=None):
        params = defaults()
        params.update(options or {})
        self.gl = gl
        self.state = params["state"]                 # held by reference and mutated by callers
        self.timer = params["timer"]
        self.flowShader = params["flowShader"]
        self.logicShader = None
        self.uniforms = {"render": {}, "update": {}}
        self.particles = None
        self.flow = None                             # becomes particles.flow once set up
        self.targets = None
        self.viewRes = [0, 0]
        self.viewSize = [0, 0]
        if params.get("logicShader") not in (None, logicFrag):
            raise N.TendrilsError("tendrils-b200: custom logic shaders are not supported")

    def setup(self, *rest):                                             # :149-154
        self.setupParticles(*rest)
        self.reset()
        return self

    def reset(self):                                                    # :156-160
        self.spawn()
        return self

    def dispose(self):                                                  # :162-169
        if self.particles:
            self.particles.dispose()
        self.particles = None
        return self

    def setupParticles(self, rootNum=None, numBuffers=2):               # :186-210
        rootNum = self.state["rootNum"] if rootNum is None else rootNum
        self.state["rootNum"] = rootNum
        shape = [rootNum, rootNum] if np.isscalar(rootNum) else [int(rootNum[0]), int(rootNum[1])]
        old = self.particles
        self.particles = Particles(self.gl, {"shape": shape, "geomShape": [shape[0], shape[1] * 2],
                                             "logicFrag": logicFrag, "render": self.flowShader})
        self.logicShader = self.particles.logic
        self.particles.setup(numBuffers)
        self.flow, self.targets = self.particles.flow, self.particles.targets
        if old is not None:                         # the flow FBO outlives a particle re-setup
            self.flow.shape = old.flow_shape
            self.flow.upload(old.flow.download())
            old.dispose()
        elif self.viewRes[0] > 0:
            self.flow.shape = self.viewRes
        return self

    def clearFlow(self):                                                # :234-239
        N.check(self.particles._ctx, N.load().tb_clear_flow(self.particles._ctx))
        return self

    def clear(self):                                                    # :220-225 (view clear is display-only)
        return self.clearFlow()

    def restart(self):                                                  # :241-246
        self.clear()
        self.reset()
        return self

    def step(self):                                                     # :248-272
        if not self.timer.paused:
            self.particles.logic = self.logicShader
            self.uniforms["update"].update(self.state)
            self.uniforms["update"].update({
                "dt": self.timer.dt, "time": self.timer.time, "start": self.timer.since,
                "flow": self.flow, "targets": self.targets,
                "viewSize": self.viewSize, "viewRes": self.viewRes})
            self.particles.step(self.uniforms["update"])
        return self

    def stepStreamed(self, host_in, host_out, chunks=16):
        """step() with the particle state coming from and going back to host memory (Particles.step_streamed)."""
        if not self.timer.paused:
            self.particles.logic = self.logicShader
            self.uniforms["update"].update(self.state)
            self.uniforms["update"].update({
                "dt": self.timer.dt, "time": self.timer.time, "start": self.timer.since,
                "flow": self.flow, "targets": self.targets,
                "viewSize": self.viewSize, "viewRes": self.viewRes})
            self.particles.step_streamed(self.uniforms["update"], host_in, host_out, chunks)
        return self

    def draw(self):                                                     # :278-303 (flow half)
        self.uniforms["render"].update(self.state)
        self.uniforms["render"].update({
            "time": self.timer.time, "previous": self.particles.buffers[1],
            "viewSize": self.viewSize, "viewRes": self.viewRes})
        self.particles.render = self.flowShader
        self.particles.draw(self.uniforms["render"], "LINES")
        return self

    def resize(self):                                                   # :393-408
        self.viewRes[0] = self.gl.drawingBufferWidth
        self.viewRes[1] = self.gl.drawingBufferHeight
        coverAspect(self.viewSize, self.viewRes)
        if self.flow is not None:
            self.flow.shape = self.viewRes
        return self

    def spawn(self, spawner=initSpawner):                               # :425-429
        if spawner is initSpawner:
            N.check(self.particles._ctx, N.load().tb_reset(self.particles._ctx))
            self.particles.pixels[..., 0:2] = INERT
            self.particles.pixels[..., 2:4] = 0
        else:
            self.particles.spawn(spawner)
        return self

    def spawnShader(self, shader, update=None, *rest):                  # :432-457
        self.timer.tick()
        self.particles.logic = shader
        base = dict(self.state)
        base.update({"time": self.timer.time, "viewSize": self.viewSize, "viewRes": self.viewRes})
        self.particles.step(Particles.applyUpdate(base, update), *rest)
        self.particles.logic = self.logicShader
        return self


__all__ = ["Tendrils", "Particles", "Device", "Shader", "Timer", "defaults", "initSpawner",
           "logicFrag", "flowShader", "shard_columns", "INERT"]
