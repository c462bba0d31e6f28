"""Seeded synthetic grayscale frames (SURVEY.md section 8d generator specification).

``synth(W, H, idx)`` is the workload generator for bench.py and the parity tests: a flat canvas,
random filled rectangles / triangles, a 3x3 box blur and Gaussian pixel noise.  The draw order of
the random stream (level; per shape gray then geometry; noise field) is part of the specification,
so frames are reproducible across machines with the same numpy / cv2.
"""
from __future__ import annotations

import numpy as np

SEED = 20261018


def synth(width: int, height: int, idx: int, seed: int = SEED) -> np.ndarray:
    """Return one ``(height, width)`` uint8 frame."""
    import cv2

    rng = np.random.default_rng([seed, width, height, idx])
    img = np.full((height, width), int(rng.integers(32, 224)), np.uint8)
    n_shapes = max(8, (width * height) // 6000)
    for k in range(n_shapes):
        gray = int(rng.integers(0, 256))
        if k % 3 != 2:
            cx, cy = int(rng.integers(0, width)), int(rng.integers(0, height))
            w = int(rng.integers(8, max(9, width // 6)))
            h = int(rng.integers(8, max(9, height // 6)))
            cv2.rectangle(img, (cx - w // 2, cy - h // 2), (cx + w // 2, cy + h // 2), gray, -1)
        else:
            centre = np.array([rng.integers(0, width), rng.integers(0, height)])
            r = max(9, min(width, height) // 8)
            pts = (centre + rng.integers(-r, r + 1, size=(3, 2))).astype(np.int32)
            cv2.fillPoly(img, [pts], gray)
    img = cv2.blur(img, (3, 3), borderType=cv2.BORDER_REPLICATE)
    noise = rng.normal(0.0, 2.0, size=(height, width))
    return np.clip(np.rint(img.astype(np.float64) + noise), 0, 255).astype(np.uint8)


def synth_batch(width: int, height: int, n: int, start: int = 0, seed: int = SEED, out: np.ndarray | None = None) -> np.ndarray:
    """Frames ``start .. start+n-1`` stacked as ``(n, height, width)`` uint8."""
    if out is None:
        out = np.empty((n, height, width), np.uint8)
    for i in range(n):
        out[i] = synth(width, height, start + i, seed)
    return out


def tiled_batch(width: int, height: int, n: int, n_unique: int = 64, start: int = 0, seed: int = SEED) -> np.ndarray:
    """A batch of ``n`` frames built from ``n_unique`` generated frames (cyclic), each copy shifted by a
    distinct circular row offset so no two frames are byte-identical.  Used by bench.py to build the
    1024-frame workload without spending minutes in the Python generator."""
    base = synth_batch(width, height, min(n, n_unique), start, seed)
    if n <= n_unique:
        return base
    out = np.empty((n, height, width), np.uint8)
    for i in range(n):
        out[i] = np.roll(base[i % n_unique], shift=(i // n_unique) * 7, axis=0)
    return out


def synth_heatmap(width: int, height: int, idx: int, quantum: float = 0.0, seed: int = SEED) -> np.ndarray:
    """A keypoint heat map of the kind a SuperPoint / DISK head emits, ``(height, width)`` float32 in [0, 1): a low noise
    floor plus a few hundred narrow peaks.  ``quantum`` > 0 rounds the values to multiples of it, which produces many
    exactly equal responses (the tie rule of the reference's multimap walk is part of the parity contract)."""
    rng = np.random.default_rng([seed, width, height, idx, 77])
    hm = (rng.random((height, width)) ** 6 * 0.12).astype(np.float32)   # a few per cent of the floor clears the default threshold 0.1
    n_peaks = max(4, (width * height) // 900)
    ys, xs = rng.integers(0, height, n_peaks), rng.integers(0, width, n_peaks)
    amp = rng.random(n_peaks).astype(np.float32)
    for y, x, a in zip(ys, xs, amp):
        y0, y1, x0, x1 = max(y - 2, 0), min(y + 3, height), max(x - 2, 0), min(x + 3, width)
        yy, xx = np.mgrid[y0:y1, x0:x1]
        blob = a * np.exp(-0.5 * ((yy - y) ** 2 + (xx - x) ** 2)).astype(np.float32)
        hm[y0:y1, x0:x1] = np.maximum(hm[y0:y1, x0:x1], blob)
    if quantum > 0.0:
        hm = (np.round(hm / np.float32(quantum)) * np.float32(quantum)).astype(np.float32)
    return np.minimum(hm, np.float32(0.999)).astype(np.float32)


def synth_descriptor_volume(channels: int, map_rows: int, map_cols: int, idx: int, seed: int = SEED) -> np.ndarray:
    """A dense descriptor volume ``(channels, map_rows, map_cols)`` float32 with unit-norm columns, as the descriptor head
    of a SuperPoint-style model produces at 1/8 resolution."""
    rng = np.random.default_rng([seed, channels, map_rows, map_cols, idx, 78])
    vol = rng.standard_normal((channels, map_rows, map_cols)).astype(np.float32)
    vol /= np.sqrt((vol.astype(np.float64) ** 2).sum(0, keepdims=True)).astype(np.float32)
    return vol
