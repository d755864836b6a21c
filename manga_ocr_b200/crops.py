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


def page_with_selections(n: int = 24, seed: int = 1006, height: int = 1400, width: int = 1000):
    """A synthetic page and ``n`` selections on it, for the region-staging path:
    speech-bubble-like polygons (12-40 points, the way the app's lasso / bubble detector produce them) around
    glyph-like blobs; every fourth selection is a plain rectangle, every fifth has its text orientation set
    against its aspect (so it is rotated), a few reach over the page border.
    Returns (page uint8 [H, W, 3], list of (rect_xywh, polygon-or-None [k, 2] int32, orientation-or-None))."""
    rng = np.random.default_rng(seed)
    page = np.full((height, width, 3), 255, np.uint8)
    page = np.clip(page.astype(np.float32) + rng.normal(0, 3, page.shape), 0, 255).astype(np.uint8)
    sels = []
    for i in range(n):
        w, h = int(rng.integers(48, 320)), int(rng.integers(64, 480))
        x, y = int(rng.integers(-10, width - w + 10)), int(rng.integers(-10, height - h + 10))
        crop = make_crop(rng, h, w, tint=(i % 10 == 0))
        y0, y1, x0, x1 = max(y, 0), min(y + h, height), max(x, 0), min(x + w, width)
        page[y0:y1, x0:x1] = crop[y0 - y:y1 - y, x0 - x:x1 - x]
        poly = None
        if i % 4 != 3:
            k = int(rng.integers(12, 41))
            ang = np.sort(rng.uniform(0, 2 * np.pi, k))
            rad = rng.uniform(0.8, 1.0, k)
            poly = np.stack([x + w / 2 + (w / 2) * rad * np.cos(ang), y + h / 2 + (h / 2) * rad * np.sin(ang)], 1)
            poly = np.round(poly).astype(np.int32)
            bx, by = int(poly[:, 0].min()), int(poly[:, 1].min())
            rect = (bx, by, int(poly[:, 0].max()) - bx + 1, int(poly[:, 1].max()) - by + 1)     # QPolygon.boundingRect()
        else:
            rect = (x, y, w, h)
        orientation = None
        if i % 5 == 4:
            orientation = "Vertical" if rect[2] > rect[3] else "Horizontal"
        sels.append((rect, poly, orientation))
    return page, sels
