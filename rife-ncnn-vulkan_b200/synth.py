"""Deterministic synthetic frame pairs: a smooth pattern plus a coordinate-hashed
texture, both translating by (dx, dy) pixels per frame, so flows are non-trivial but well-posed.
Pure numpy, bit-identical on every host (integer hash + float32 math done in float64 then rounded)."""
import numpy as np


def _hash32(c, y, x):
    v = (x.astype(np.uint64) * np.uint64(73856093)) ^ (y.astype(np.uint64) * np.uint64(19349663)) ^ np.uint64((c + 1) * 83492791)
    v &= np.uint64(0xFFFFFFFF)
    v ^= v >> np.uint64(16)
    v = (v * np.uint64(0x7FEB352D)) & np.uint64(0xFFFFFFFF)
    v ^= v >> np.uint64(15)
    v = (v * np.uint64(0x846CA68B)) & np.uint64(0xFFFFFFFF)
    v ^= v >> np.uint64(16)
    return v


def frame(k, w, h, dx=3, dy=2, seed=0):
    """Frame k of the stream as packed RGB u8, shape (h, w, 3)."""
    ys, xs = np.meshgrid(np.arange(h, dtype=np.int64), np.arange(w, dtype=np.int64), indexing="ij")
    xs = xs - dx * k + 100003 * (seed + 1)
    ys = ys - dy * k + 100019 * (seed + 1)
    out = np.empty((h, w, 3), np.uint8)
    for c in range(3):
        base = 0.5 + 0.4 * np.sin(0.05 * xs + c) * np.cos(0.07 * ys)
        # texture at 4x4 blocks so that it survives the 1/8 .. 1/2 pyramid
        n = ((_hash32(c, ys >> 2, xs >> 2) >> np.uint64(9)) & np.uint64(255)).astype(np.float64) / 255.0 * 0.10 - 0.05
        v = np.clip(base + n, 0.0, 1.0)
        out[:, :, c] = np.floor(v * 255.0 + 0.5).astype(np.uint8)
    return out


def pair(w, h, k=0, dx=3, dy=2, seed=0):
    return frame(k, w, h, dx, dy, seed), frame(k + 1, w, h, dx, dy, seed)


def stream(k0, n, w, h, dx=3, dy=2, seed=0):
    """Frames k0 .. k0 + n - 1 of the stream, bit-identical to frame(k, ...): the stream is one pattern translating by whole
    pixels, so every frame is a window of ONE canvas (computed once: n frames for the price of about one and a half)."""
    if n <= 0:
        return []
    kmax = k0 + n - 1
    adx, ady = abs(dx) * (n - 1), abs(dy) * (n - 1)
    # canvas[Y, X] = pattern(X - X0, Y - Y0) with (X0, Y0) chosen so that every frame's window starts at a non-negative offset
    ys, xs = np.meshgrid(np.arange(h + ady, dtype=np.int64), np.arange(w + adx, dtype=np.int64), indexing="ij")
    ox = dx * kmax if dx >= 0 else dx * k0
    oy = dy * kmax if dy >= 0 else dy * k0
    xs = xs - ox + 100003 * (seed + 1)
    ys = ys - oy + 100019 * (seed + 1)
    canvas = np.empty((h + ady, w + adx, 3), np.uint8)
    for c in range(3):
        base = 0.5 + 0.4 * np.sin(0.05 * xs + c) * np.cos(0.07 * ys)
        n_ = ((_hash32(c, ys >> 2, xs >> 2) >> np.uint64(9)) & np.uint64(255)).astype(np.float64) / 255.0 * 0.10 - 0.05
        v = np.clip(base + n_, 0.0, 1.0)
        canvas[:, :, c] = np.floor(v * 255.0 + 0.5).astype(np.uint8)
    out = []
    for k in range(k0, k0 + n):
        x0, y0 = ox - dx * k, oy - dy * k  # frame k reads the pattern at (x - dx k, y - dy k) = canvas column x + (ox - dx k)
        out.append(np.ascontiguousarray(canvas[y0:y0 + h, x0:x0 + w]))
    return out
