"""Seeded synthetic microscopy frames: one blurred disc ("nucleus") per cell of
a jittered 26-px grid on a noisy background, uint16.  Pure NumPy so it runs on the GPU box.

``nuclei(H, W, seed)`` -> (H, W) uint16 with (H//26)*(W//26) non-touching nuclei
(8192^2 -> 99 225, 4096^2 -> 24 649, 2048^2 -> 6 084).
"""
from __future__ import annotations

import numpy as np

PITCH = 26


def _blur_axis(a: np.ndarray, taps: np.ndarray, axis: int) -> np.ndarray:
    r = len(taps) // 2
    pad = [(0, 0)] * a.ndim
    pad[axis] = (r, r)
    p = np.pad(a, pad, mode="reflect")
    out = np.zeros_like(a)
    n = a.shape[axis]
    for i, t in enumerate(taps):
        sl = [slice(None)] * a.ndim
        sl[axis] = slice(i, i + n)
        out += np.float32(t) * p[tuple(sl)]
    return out


def nuclei(height: int, width: int, seed: int = 1, pitch: int = PITCH) -> np.ndarray:
    rng = np.random.default_rng(seed)
    ny, nx = height // pitch, width // pitch
    canvas = np.zeros((height, width), np.float32)
    if ny and nx:
        cy = (pitch // 2 + rng.integers(-4, 5, (ny, nx))).astype(np.float32)
        cx = (pitch // 2 + rng.integers(-4, 5, (ny, nx))).astype(np.float32)
        rad = rng.integers(4, 8, (ny, nx)).astype(np.float32)
        amp = rng.uniform(0.3, 1.0, (ny, nx)).astype(np.float32)
        yy = np.arange(pitch, dtype=np.float32)
        # cells[y, i, x, j]: local pixel (i, j) of grid cell (y, x)
        d2 = (yy[None, :, None, None] - cy[:, None, :, None]) ** 2 + (yy[None, None, None, :] - cx[:, None, :, None]) ** 2
        cells = np.where(d2 <= (rad * rad)[:, None, :, None], amp[:, None, :, None], np.float32(0))
        canvas[: ny * pitch, : nx * pitch] = cells.reshape(ny * pitch, nx * pitch)
    sigma = 1.5
    r = 5
    x = np.arange(-r, r + 1, dtype=np.float64)
    taps = np.exp(-0.5 * (x / sigma) ** 2)
    taps /= taps.sum()
    canvas = _blur_axis(_blur_axis(canvas, taps, 0), taps, 1)
    noise = rng.standard_normal((height, width), dtype=np.float32) * np.float32(300.0)
    img = canvas * np.float32(40000.0) + np.float32(2000.0) + noise
    return np.clip(img, 0, 65535).astype(np.uint16)


def nuclei_bgr(height: int, width: int, seed: int = 1) -> np.ndarray:
    """3-channel variant with +-2 % channel gains (exercises K1)."""
    g = nuclei(height, width, seed).astype(np.float32)
    chans = [np.clip(g * np.float32(k), 0, 65535).astype(np.uint16) for k in (0.98, 1.0, 1.02)]
    return np.ascontiguousarray(np.stack(chans, axis=-1))


def expected_nuclei(height: int, width: int, pitch: int = PITCH) -> int:
    return (height // pitch) * (width // pitch)


__all__ = ["nuclei", "nuclei_bgr", "expected_nuclei", "PITCH"]
