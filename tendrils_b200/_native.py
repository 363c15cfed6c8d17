"""ctypes binding of include/tendrils_b200.h.  Fails loudly when the CUDA library is missing:
there is no CPU fallback anywhere in this package."""
from __future__ import annotations

import ctypes as C
import os

from . import build as _build

_fp = C.POINTER(C.c_float)


class TbConfig(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "particles_w", "particles_h", "col0", "col1", "flow_w", "flow_h", "device", "flags")]


STATE_FIELDS = ("damping", "speedLimit", "forceWeight", "varyForce", "flowWeight", "varyFlow",
                "noiseWeight", "varyNoise", "flowDecay", "flowWidth", "noiseScale", "varyNoiseScale",
                "noiseSpeed", "varyNoiseSpeed", "target", "varyTarget")


class TbState(C.Structure):
    _fields_ = [(n, C.c_float) for n in STATE_FIELDS] + [("viewSize", C.c_float * 2)]


class TbPixelSpawner(C.Structure):
    _fields_ = [("spawnSize", C.c_float * 2), ("jitter", C.c_float * 2), ("speed", C.c_float),
                ("bias", C.c_float), ("spawnMatrix", C.c_float * 9)]


class TbOpticalFlowParams(C.Structure):
    _fields_ = [("viewSize", C.c_float * 2), ("scaleUV", C.c_float * 2), ("offset", C.c_float), ("lambda_", C.c_float),
                ("speed", C.c_float), ("speedLimit", C.c_float), ("time", C.c_float)]


class TbFlowLineParams(C.Structure):
    _fields_ = [("viewSize", C.c_float * 2), ("rad", C.c_float), ("speed", C.c_float), ("speedLimit", C.c_float),
                ("crestShape", C.c_float)]


TB_TARGET_STATE, TB_TARGET_TARGETS = 0, 1
TB_SPAWN_DIRECT, TB_SPAWN_BEST_SAMPLE, TB_SPAWN_BRIGHT_SAMPLE = 0, 1, 2
TB_SPAWN_COLOR_SAMPLE, TB_SPAWN_DATA_SAMPLE, TB_SPAWN_FLOW_SAMPLE = 3, 4, 5
TB_SOURCE_IMAGE, TB_SOURCE_FLOW, TB_SOURCE_PARTICLES = 0, 1, 2
TB_BUF_CURRENT, TB_BUF_PREVIOUS, TB_BUF_TARGETS, TB_BUF_FLOW = 0, 1, 2, 3

# every symbol include/tendrils_b200.h declares: name -> (restype, argtypes)
_ctx = C.c_void_p
SYMBOLS = {
    "tb_abi_version": (C.c_int, []),
    "tb_last_error": (C.c_char_p, [_ctx]),
    "tb_create": (C.c_int, [C.POINTER(TbConfig), C.POINTER(_ctx)]),
    "tb_destroy": (C.c_int, [_ctx]),
    "tb_set_state": (C.c_int, [_ctx, C.POINTER(TbState)]),
    "tb_set_overlap": (C.c_int, [_ctx, C.c_int32]),
    "tb_resize_flow": (C.c_int, [_ctx, C.c_int32, C.c_int32]),
    "tb_clear_flow": (C.c_int, [_ctx]),
    "tb_step": (C.c_int, [_ctx, C.c_float, C.c_float]),
    "tb_step_streamed": (C.c_int, [_ctx, C.c_float, C.c_float, _fp, _fp, C.c_int32]),
    "tb_splat_flow": (C.c_int, [_ctx, C.c_float]),
    "tb_splat_collect": (C.c_int, [_ctx, C.c_float]),
    "tb_splat_fold": (C.c_int, [_ctx]),
    "tb_owners_handle_bytes": (C.c_int64, []),
    "tb_owners_export": (C.c_int, [_ctx, C.c_int64, C.c_void_p, C.c_int64]),
    "tb_owners_connect": (C.c_int, [_ctx, C.c_int32, C.c_int32, C.c_void_p, C.c_int64]),
    "tb_splat_flow_owners": (C.c_int, [_ctx, C.c_float]),
    "tb_reset": (C.c_int, [_ctx]),
    "tb_spawn_init": (C.c_int, [_ctx, C.c_int]),
    "tb_spawn_ball": (C.c_int, [_ctx, C.c_float, C.c_float, C.c_int]),
    "tb_set_spawn_image": (C.c_int, [_ctx, _fp, C.c_int32, C.c_int32]),
    "tb_spawn_pixels": (C.c_int, [_ctx, C.POINTER(TbPixelSpawner), C.c_int, C.c_int, C.c_float, C.c_int]),
    "tb_upload": (C.c_int, [_ctx, C.c_int, _fp, C.c_int64]),
    "tb_download": (C.c_int, [_ctx, C.c_int, _fp, C.c_int64]),
    "tb_blend_into_flow": (C.c_int, [_ctx, _fp, C.c_int32, C.c_int32]),
    "tb_debug_max_bins": (C.c_int, []),
    "tb_debug_segments": (C.c_int, [_ctx, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int64)]),
    "tb_debug_bins": (C.c_int, [_ctx, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_int32), C.POINTER(C.c_int32),
                                C.POINTER(C.c_int32)]),
    "tb_flow_line": (C.c_int, [_ctx, C.POINTER(TbFlowLineParams), C.c_int32, _fp, _fp, _fp, _fp, _fp, _fp]),
    "tb_optical_flow": (C.c_int, [_ctx, C.POINTER(TbOpticalFlowParams), C.c_void_p, C.c_void_p, C.c_int32, C.c_int32]),
    "tb_device_ptr": (C.c_int, [_ctx, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_int64)]),
    "tb_stream": (C.c_int, [_ctx, C.POINTER(C.c_void_p)]),
    "tb_wait_stream": (C.c_int, [_ctx, C.c_void_p]),
    "tb_sync": (C.c_int, [_ctx]),
    "tb_stats": (C.c_int, [_ctx, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "tb_timing": (C.c_int, [_ctx, C.c_int, C.POINTER(C.c_int64), _fp, C.POINTER(C.c_int64), _fp,
                            C.POINTER(C.c_int64), _fp]),
}

_lib = None


class TendrilsError(RuntimeError):
    """Raised where the reference would throw a stack.gl Error (status != 0)."""


def lib_path() -> str:
    return _build.LIB_PATH


def load():
    """Load libtendrils_b200.so; raise if it has not been built (no fallback)."""
    global _lib
    if _lib is None:
        path = lib_path()
        if not os.path.exists(path):
            raise TendrilsError(
                f"tendrils_b200: CUDA library {path} is missing -- run `python -m tendrils_b200.build` "
                "(there is no CPU fallback)")
        L = C.CDLL(path)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(L, name)       # AttributeError if the header and the library diverge
            fn.restype = res
            fn.argtypes = args
        if L.tb_abi_version() != 2:
            raise TendrilsError("tendrils_b200: ABI version mismatch")
        _lib = L
    return _lib


def check(ctx, status: int):
    if status != 0:
        msg = load().tb_last_error(ctx)
        raise TendrilsError((msg or b"unknown error").decode() + f" (status {status})")


def is_device_array(a) -> bool:
    """A torch CUDA tensor (anything with .is_cuda set): its memory is handed to the library as a device pointer."""
    return bool(getattr(a, "is_cuda", False))


def wait_for_producer(ctx, a):
    """A torch CUDA tensor was (possibly) written on torch's current stream: order the library's stream after it."""
    if is_device_array(a):
        import torch
        check(ctx, load().tb_wait_stream(ctx, C.c_void_p(torch.cuda.current_stream(a.device).cuda_stream)))


def array_pointer(a, dtype: str) -> int:
    """Address of a contiguous numpy array or torch CUDA tensor of the given dtype name."""
    if is_device_array(a):
        if str(a.dtype) != f"torch.{dtype}" or not a.is_contiguous():
            raise TendrilsError(f"tendrils-b200: device arrays must be contiguous {dtype}")
        return a.data_ptr()
    return a.ctypes.data
