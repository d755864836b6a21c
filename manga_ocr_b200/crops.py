"""Synthetic manga-style text-bubble crops for the benchmark and the tests.

Shapes and seeds follow SURVEY.md section 8(d): white background, a few
vertical columns of dark glyph-like blobs, mild Gaussian noise, stored as
``uint8 [H, W, 3]`` RGB exactly like the arrays the app hands to the engine
(reference/src/ui/main_window.py:9800).  numpy PCG64 only, so the same crops
are produced on every machine.
"""
from __future__ import annotations

from typing import List, Tuple

import numpy as np


def make_crop(rng: np.random.Generator, h: int, w: int, tint: bool = False) -> np.ndarray:
    img = np.full((h, w), 255.0, dtype=np.float32)
    ncol = int(rng.integers(1, max(2, min(10, w // 24)) + 1))
    colw = max(2, w // (ncol + 1))
    for c in range(ncol):
        cx = int((c + 0.5) * w / ncol)
        gh = max(3, min(colw, 28))
        y = int(rng.integers(2, 8))
        while y + gh < h - 2:
            # one glyph = a handful of strokes inside a gh x gh cell
            for _ in range(int(rng.integers(2, 6))):
                x0 = cx - gh // 2 + int(rng.integers(0, gh))
                y0 = y + int(rng.integers(0, gh))
                ln = int(rng.integers(2, gh + 1))
                th = int(rng.integers(1, 3))
                val = float(rng.integers(0, 61))
                if rng.random() < 0.5:
                    img[max(0, y0):min(h, y0 + th), max(0, x0):min(w, x0 + ln)] = val
                else:
                    img[max(0, y0):min(h, y0 + ln), max(0, x0):min(w, x0 + th)] = val
            y += gh + int(rng.integers(1, 6))
    img += rng.normal(0.0, 3.0, size=img.shape).astype(np.float32)
    g = np.clip(np.rint(img), 0, 255).astype(np.uint8)
    rgb = np.repeat(g[:, :, None], 3, axis=2)
    if tint:
        scale = rng.uniform(0.6, 1.0, size=3).astype(np.float32)
        rgb = np.clip(np.rint(rgb.astype(np.float32) * scale), 0, 255).astype(np.uint8)
    return np.ascontiguousarray(rgb)


def _sizes_bubble(rng, n) -> List[Tuple[int, int]]:
    return [(int(rng.integers(64, 481)), int(rng.integers(48, 321))) for _ in range(n)]


def _loguniform(rng, lo, hi) -> int:
    return int(round(float(np.exp(rng.uniform(np.log(lo), np.log(hi))))))


def single_224(seed: int = 1001) -> List[np.ndarray]:
    """Config 1: one 224x224 grayscale crop."""
    rng = np.random.default_rng(seed)
    return [make_crop(rng, 224, 224)]


def bubble_batch(n: int = 64, seed: int = 1002) -> List[np.ndarray]:
    """Config 2: n bubble-like crops, W in [48,320], H in [64,480]."""
    rng = np.random.default_rng(seed)
    return [make_crop(rng, h, w) for h, w in _sizes_bubble(rng, n)]


def page_batch(n: int = 512, seed: int = 1003) -> List[np.ndarray]:
    """Config 3 / 5: mixed-size crops, W,H log-uniform in [32,1024]; 10 % tinted,
    5 % exactly 224x224, 5 % with one 2-8 px dimension."""
    rng = np.random.default_rng(seed)
    out = []
    for _ in range(n):
        u = rng.random()
        if u < 0.05:
            h, w = 224, 224
        elif u < 0.10:
            h, w = _loguniform(rng, 32, 1024), int(rng.integers(2, 9))
            if rng.random() < 0.5:
                h, w = w, h
        else:
            h, w = _loguniform(rng, 32, 1024), _loguniform(rng, 32, 1024)
        out.append(make_crop(rng, h, w, tint=bool(rng.random() < 0.10)))
    return out


def tall_batch(n: int = 64, seed: int = 1004) -> List[np.ndarray]:
    """Config 4: long vertical-text crops, W in [40,120], H in [600,1600]."""
    rng = np.random.default_rng(seed)
    return [make_crop(rng, int(rng.integers(600, 1601)), int(rng.integers(40, 121))) for _ in range(n)]
