"""ORACLE (test infrastructure) - golden fixtures for the crop-staging row.

Runs the library calls the reference makes (reference/src/ui/main_window.py:6497-6506, 9789-9800:
``PIL.Image.crop``, ``cv2.cvtColor``, ``cv2.fillPoly``, ``cv2.bitwise_and`` / ``bitwise_not`` / ``add``,
``cv2.rotate``) with opencv-python 4.13.0 and Pillow 12.2.0 in the build container, on a seeded page and
seeded selections, and stores

  masks    the fillPoly masks (bit-packed)                     -> pins oracle.crop_staging_np.fill_poly_mask
  digests  SHA-256 + shape of the RGB image handed to MangaOcr -> pins oracle.crop_staging_np.stage_region

Run:  python -m oracle.make_golden_staging      (writes tests/golden/staging_kat.npz, a few KB)
"""
from __future__ import annotations

import hashlib
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "tests", "golden", "staging_kat.npz")
PAGE_SEED, PAGE_H, PAGE_W = 909, 420, 360


def page_rgb() -> np.ndarray:
    rng = np.random.default_rng(PAGE_SEED)
    yy, xx = np.mgrid[0:PAGE_H, 0:PAGE_W]
    base = (127 + 90 * np.sin(xx / 17.0) * np.cos(yy / 23.0))[..., None] + rng.normal(0, 25, (PAGE_H, PAGE_W, 3))
    return np.clip(base + np.array([10, -20, 30]), 0, 255).astype(np.uint8)


def selections():
    """(box, polygon-or-None in page coordinates, rotation code) - seeded.  Boxes follow the reference's Qt
    convention (right = x + w - 1 handed to PIL as the EXCLUSIVE end, so the polygon's last column/row falls
    outside the crop), some reach beyond the page, polygons include concave and self-intersecting ones."""
    rng = np.random.default_rng(910)
    out = []
    for i in range(28):
        n = int(rng.integers(3, 14))
        cx, cy = int(rng.integers(20, PAGE_W - 20)), int(rng.integers(20, PAGE_H - 20))
        rx, ry = int(rng.integers(8, 120)), int(rng.integers(8, 150))
        if i % 4 == 3:      # random (self-intersecting) vertex order
            pts = np.stack([cx + rng.integers(-rx, rx + 1, n), cy + rng.integers(-ry, ry + 1, n)], 1)
        else:               # star-shaped, concave
            ang = np.sort(rng.uniform(0, 2 * np.pi, n))
            rad = rng.uniform(0.35, 1.0, n)
            pts = np.stack([cx + rx * rad * np.cos(ang), cy + ry * rad * np.sin(ang)], 1)
        pts = np.round(pts).astype(np.int32)
        x0, y0 = int(pts[:, 0].min()), int(pts[:, 1].min())
        x1, y1 = int(pts[:, 0].max()), int(pts[:, 1].max())
        box = (x0, y0, x1, y1)                                  # QRect.right()/bottom() = max coordinate
        if i % 7 == 5:
            box = (x0 - 9, y0 - 4, x1 + 13, y1 + 6)             # a box that is not the bounding rectangle
        rot = int(i % 3)
        out.append((box, pts if i % 5 != 4 else None, rot))
    out.append(((-7, -5, 40, 33), None, 1))                      # rectangle selections reaching outside the page
    out.append(((PAGE_W - 30, PAGE_H - 25, PAGE_W + 12, PAGE_H + 9), np.array([[PAGE_W - 28, PAGE_H - 20], [PAGE_W + 10, PAGE_H - 2], [PAGE_W - 10, PAGE_H + 8]], np.int32), 2))
    return out


def reference_stage(page: np.ndarray, box, polygon, rot: int):
    """The reference's own sequence of library calls; returns (mask or None, RGB array given to MangaOcr)."""
    import cv2
    from PIL import Image
    pil = Image.fromarray(page)
    cropped = pil.crop(box)
    bgr = cv2.cvtColor(np.array(cropped), cv2.COLOR_RGB2BGR)
    mask = None
    if polygon is not None:
        mask = np.zeros(bgr.shape[:2], dtype=np.uint8)
        rel = np.array([[int(p[0]) - box[0], int(p[1]) - box[1]] for p in polygon], dtype=np.int32)
        cv2.fillPoly(mask, [rel], 255)
        white = np.full(bgr.shape, 255, dtype=np.uint8)
        fg = cv2.bitwise_and(bgr, bgr, mask=mask)
        bg = cv2.bitwise_and(white, white, mask=cv2.bitwise_not(mask))
        bgr = cv2.add(fg, bg)
    if rot == 1:
        bgr = cv2.rotate(bgr, cv2.ROTATE_90_CLOCKWISE)
    elif rot == 2:
        bgr = cv2.rotate(bgr, cv2.ROTATE_90_COUNTERCLOCKWISE)
    return mask, np.array(Image.fromarray(cv2.cvtColor(bgr, cv2.COLOR_BGR2RGB)))


def main() -> None:
    page = page_rgb()
    data = {}
    for i, (box, poly, rot) in enumerate(selections()):
        mask, rgb = reference_stage(page, box, poly, rot)
        data[f"box{i}"] = np.array(box, np.int32)
        data[f"rot{i}"] = np.int32(rot)
        data[f"poly{i}"] = np.zeros((0, 2), np.int32) if poly is None else np.asarray(poly, np.int32)
        data[f"has_poly{i}"] = np.int32(poly is not None)
        data[f"mask{i}"] = np.zeros((0,), np.uint8) if mask is None else np.packbits(mask != 0)
        data[f"shape{i}"] = np.array(rgb.shape, np.int32)
        data[f"sha{i}"] = np.frombuffer(hashlib.sha256(np.ascontiguousarray(rgb).tobytes()).digest(), np.uint8)
    data["n"] = np.int32(len(selections()))
    np.savez_compressed(OUT, **data)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
