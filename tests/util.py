"""Shared helpers for the parity tests."""
import numpy as np


def bits_equal(a, b):
    """Bit-exact float32 comparison that treats any NaN as equal to any NaN and -0 == +0."""
    a = np.asarray(a, np.float32)
    b = np.asarray(b, np.float32)
    assert a.shape == b.shape, (a.shape, b.shape)
    both_nan = np.isnan(a) & np.isnan(b)
    return (a == b) | both_nan


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
