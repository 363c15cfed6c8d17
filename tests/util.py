"""Shared helpers for the parity tests."""
import numpy as np


def bits_equal(a, b):
    """Bit-exact float32 comparison: the raw 32-bit patterns must agree (so -0 differs from +0); only the payload of a
    NaN is left open (any NaN equals any NaN)."""
    a = np.ascontiguousarray(a, np.float32)
    b = np.ascontiguousarray(b, np.float32)
    assert a.shape == b.shape, (a.shape, b.shape)
    both_nan = np.isnan(a) & np.isnan(b)
    return (a.view(np.uint32) == b.view(np.uint32)) | both_nan


def assert_bits_equal(a, b, what=""):
    ok = bits_equal(a, b)
    if not ok.all():
        a = np.asarray(a, np.float32); b = np.asarray(b, np.float32)
        idx = np.argwhere(~ok)
        first = tuple(idx[0])
        raise AssertionError(
            f"{what}: {len(idx)} of {ok.size} values differ; first at {first}: {a[first]!r} vs {b[first]!r}")


def synthetic_image(w, h, seed=7):
    """Smooth, closed-form RGBA float image in [0,1] (the bench's spawn image)."""
    y, x = np.mgrid[0:h, 0:w].astype(np.float64)
    u, v = (x + 0.5) / w, (y + 0.5) / h
    r = 0.5 + 0.5 * np.sin(6.0 * u + 1.3 * seed) * np.cos(4.0 * v)
    g = 0.5 + 0.5 * np.sin(5.0 * v - 0.7 * seed + 3.0 * u)
    b = 0.5 + 0.5 * np.cos(7.0 * (u - 0.5) * (v - 0.5) * 4 + seed)
    a = 0.25 + 0.75 * (0.5 + 0.5 * np.sin(3.0 * (u + v)))
    return np.stack([r, g, b, a], -1).astype(np.float32)


def synthetic_video(w, h, count=8):
    """`count` RGBA8 frames [h,w,4] of a closed-form looping video: five gaussian blobs circling the centre
    ("moving gaussian blobs").  Frame k+count equals frame k."""
    ys, xs = np.meshgrid((np.arange(h, dtype=np.float32) + 0.5) / h, (np.arange(w, dtype=np.float32) + 0.5) / w, indexing="ij")
    blobs = [  # orbit radius, phase, sigma, rgb
        (0.30, 0.00, 0.060, (1.0, 0.3, 0.2)), (0.22, 0.21, 0.045, (0.2, 1.0, 0.4)), (0.36, 0.47, 0.080, (0.3, 0.4, 1.0)),
        (0.12, 0.63, 0.035, (1.0, 1.0, 0.5)), (0.41, 0.82, 0.050, (0.9, 0.5, 1.0))]
    frames = []
    for k in range(count):
        rgb = np.zeros((h, w, 3), np.float32)
        for orbit, phase, sigma, colour in blobs:
            a = 2.0 * np.pi * (k / count + phase)
            cx, cy = 0.5 + orbit * np.cos(a), 0.5 + orbit * np.sin(a)
            g = np.exp(-((xs - np.float32(cx)) ** 2 + (ys - np.float32(cy)) ** 2) / np.float32(2.0 * sigma * sigma))
            rgb += g[..., None] * np.asarray(colour, np.float32)
        frame = np.empty((h, w, 4), np.uint8)
        frame[..., :3] = np.clip(np.rint(np.clip(rgb, 0.0, 1.0) * 255.0), 0, 255).astype(np.uint8)
        frame[..., 3] = 255
        frames.append(frame)
    return frames
