"""ORACLE - TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

CPU restatement of the reference's crop staging in front of the Manga-OCR engine (SURVEY.md
section 8f, row N2):

  reference/src/ui/main_window.py:6497-6506   page crop (PIL ``crop`` on the bounding box), polygon
                                              mask with ``cv2.fillPoly``, composite on white
  reference/src/ui/main_window.py:6429-6430   the same crop without a polygon (rectangle selection)
  reference/src/ui/main_window.py:9789-9795   optional 90-degree rotation by text orientation
  reference/src/ui/main_window.py:9800        BGR -> RGB, wrapped as a PIL image for ``MangaOcr``

The arithmetic lives in OpenCV (``cv2.fillPoly`` / ``cv2.rotate`` / ``cv2.add``; not vendored, not
pinned by the reference - README "pip install opencv-python") and Pillow (``Image.crop``).  The mask
rasteriser below restates OpenCV 4.x ``fillPoly`` (modules/imgproc/src/drawing.cpp: CollectPolyEdges,
FillEdgeCollection, Line/LineIterator, clipLine) for ``lineType=LINE_8, shift=0``.  It is PINNED:
bit-identical to opencv-python 4.13.0 in the build container on the fixtures written by
``oracle/make_golden_staging.py`` (tests/golden/staging_kat.npz) and on 12,000 further random
polygons (in-bounds, out-of-bounds, degenerate, self-intersecting) during development.
"""
from __future__ import annotations

from typing import Optional, Sequence, Tuple

import numpy as np

XY_SHIFT = 16
XY_ONE = 1 << XY_SHIFT
HALF = XY_ONE >> 1

ROT_NONE, ROT_CW, ROT_CCW = 0, 1, 2


def _tdiv(a: int, b: int) -> int:
    """C++ integer division (truncation toward zero)."""
    q = abs(a) // abs(b)
    return -q if (a < 0) != (b < 0) else q


def clip_line(w: int, h: int, p1: Tuple[int, int], p2: Tuple[int, int]):
    """``cv::clipLine(Size(w, h), pt1, pt2)``; returns (inside, pt1, pt2) with the partial updates the
    C++ code leaves behind when the segment turns out to be outside (fillPoly uses them)."""
    x1, y1 = p1
    x2, y2 = p2
    right, bottom = w - 1, h - 1
    if w <= 0 or h <= 0:
        return False, (x1, y1), (x2, y2)
    c1 = (x1 < 0) + (x1 > right) * 2 + (y1 < 0) * 4 + (y1 > bottom) * 8
    c2 = (x2 < 0) + (x2 > right) * 2 + (y2 < 0) * 4 + (y2 > bottom) * 8
    if (c1 & c2) == 0 and (c1 | c2) != 0:
        if c1 & 12:
            a = 0 if c1 < 8 else bottom
            x1 += int(float(a - y1) * (x2 - x1) / (y2 - y1))
            y1 = a
            c1 = (x1 < 0) + (x1 > right) * 2
        if c2 & 12:
            a = 0 if c2 < 8 else bottom
            x2 += int(float(a - y2) * (x2 - x1) / (y2 - y1))
            y2 = a
            c2 = (x2 < 0) + (x2 > right) * 2
        if (c1 & c2) == 0 and (c1 | c2) != 0:
            if c1:
                a = 0 if c1 == 1 else right
                y1 += int(float(a - x1) * (y2 - y1) / (x2 - x1))
                x1 = a
                c1 = 0
            if c2:
                a = 0 if c2 == 1 else right
                y2 += int(float(a - x2) * (y2 - y1) / (x2 - x1))
                x2 = a
                c2 = 0
    return (c1 | c2) == 0, (x1, y1), (x2, y2)


def line8(mask: np.ndarray, p1: Tuple[int, int], p2: Tuple[int, int]) -> None:
    """``cv::line(mask, p1, p2, 255)`` for LINE_8, thickness 1: LineIterator (left to right) on the
    segment clipped to the image."""
    h, w = mask.shape
    x1, y1 = p1
    x2, y2 = p2
    if not (0 <= x1 < w and 0 <= x2 < w and 0 <= y1 < h and 0 <= y2 < h):
        ok, (x1, y1), (x2, y2) = clip_line(w, h, (x1, y1), (x2, y2))
        if not ok:
            return
    dx, dy = x2 - x1, y2 - y1
    sy = 1
    if dx < 0:                     # left_to_right: start from the left end
        dx, dy = -dx, -dy
        x1, y1 = x2, y2
    if dy < 0:
        dy, sy = -dy, -1
    vert = dy > dx
    if vert:
        dx, dy = dy, dx
    err = dx - (dy + dy)
    plus, minus = dx + dx, -(dy + dy)
    x, y = x1, y1
    for _ in range(dx + 1):
        mask[y, x] = 255
        neg = err < 0
        err += minus + (plus if neg else 0)
        if vert:
            y += sy
            x += 1 if neg else 0
        else:
            x += 1
            y += sy if neg else 0


def poly_edges(h: int, w: int, pts: Sequence[Sequence[int]]):
    """CollectPolyEdges: (line segments to draw, fill edges (y0, y1, x_fixed, dx_fixed)).  Edge x carries
    the +0.5 pixel offset of the non-antialiased path; an edge with an endpoint outside the image takes
    its slope and x from the CLIPPED segment (y only when that segment is not horizontal)."""
    n = len(pts)
    lines, edges = [], []
    if n == 0:
        return lines, edges
    pt0 = (int(pts[-1][0]), int(pts[-1][1]))
    for i in range(n):
        pt1 = (int(pts[i][0]), int(pts[i][1]))
        lines.append((pt0, pt1))
        c0x, c0y, c1x, c1y = pt0[0], pt0[1], pt1[0], pt1[1]
        if not (0 <= pt0[0] < w and 0 <= pt1[0] < w and 0 <= pt0[1] < h and 0 <= pt1[1] < h):
            _, t0, t1 = clip_line(w, h, pt0, pt1)
            c0x, c1x = t0[0], t1[0]
            if t0[1] != t1[1]:
                c0y, c1y = t0[1], t1[1]
        if pt0[1] != pt1[1]:
            f0, f1 = (c0x << XY_SHIFT) + HALF, (c1x << XY_SHIFT) + HALF
            dx = _tdiv(f1 - f0, c1y - c0y)
            if pt0[1] < pt1[1]:
                edges.append((pt0[1], pt1[1], f0 + (pt0[1] - c0y) * dx, dx))
            else:
                edges.append((pt1[1], pt0[1], f1 + (pt1[1] - c1y) * dx, dx))
        pt0 = pt1
    return lines, edges


def fill_poly_mask(h: int, w: int, pts: Sequence[Sequence[int]]) -> np.ndarray:
    """``mask = zeros((h, w), uint8); cv2.fillPoly(mask, [pts], 255)`` (main_window.py:6499-6502)."""
    mask = np.zeros((h, w), np.uint8)
    lines, edges = poly_edges(h, w, pts)
    for a, b in lines:
        line8(mask, a, b)
    if len(edges) < 2:
        return mask
    y_min = min(e[0] for e in edges)
    y_max = max(e[1] for e in edges)
    x_end = [e[2] + (e[1] - e[0]) * e[3] for e in edges]
    x_max = max(max(e[2] for e in edges), max(x_end))
    x_min = min(min(e[2] for e in edges), min(x_end))
    if y_max < 0 or y_min >= h or x_max < 0 or x_min >= (w << XY_SHIFT):
        return mask
    for y in range(max(y_min, 0), min(y_max, h)):
        xs = sorted(e[2] + (y - e[0]) * e[3] for e in edges if e[0] <= y < e[1])
        for k in range(0, len(xs) - 1, 2):
            xa = (xs[k] + HALF - 1) >> XY_SHIFT          # ceil of the crossing (x carries +0.5)
            xb = (xs[k + 1] - HALF) >> XY_SHIFT          # floor
            if xa < w and xb >= 0:
                mask[y, max(xa, 0):min(xb, w - 1) + 1] = 255
    return mask


def pil_crop(page: np.ndarray, box: Tuple[int, int, int, int]) -> np.ndarray:
    """``PIL.Image.crop(box)`` of an [H, W, C] uint8 page: size (right-left, bottom-top), zeros outside."""
    left, top, right, bottom = box
    h, w = max(bottom - top, 0), max(right - left, 0)
    out = np.zeros((h, w) + page.shape[2:], np.uint8)
    y0, y1 = max(top, 0), min(bottom, page.shape[0])
    x0, x1 = max(left, 0), min(right, page.shape[1])
    if y1 > y0 and x1 > x0:
        out[y0 - top:y1 - top, x0 - left:x1 - left] = page[y0:y1, x0:x1]
    return out


def rotation_for(orientation: Optional[str], h: int, w: int) -> int:
    """main_window.py:9789-9795: vertical text in a wide crop -> 90 degrees clockwise; horizontal text in
    a tall crop -> 90 degrees counter-clockwise."""
    if orientation == "Vertical" and w > h:
        return ROT_CW
    if orientation == "Horizontal" and h > w:
        return ROT_CCW
    return ROT_NONE


def rotate90(img: np.ndarray, rot: int) -> np.ndarray:
    """``cv2.rotate`` with ROTATE_90_CLOCKWISE / ROTATE_90_COUNTERCLOCKWISE."""
    if rot == ROT_CW:
        return np.ascontiguousarray(np.rot90(img, k=-1))
    if rot == ROT_CCW:
        return np.ascontiguousarray(np.rot90(img, k=1))
    return img


def stage_region(page_rgb: np.ndarray, box: Tuple[int, int, int, int], polygon: Optional[Sequence[Sequence[int]]] = None,
                 rot: int = ROT_NONE) -> np.ndarray:
    """The RGB uint8 image the reference hands to ``MangaOcr`` for one selection: crop, optional polygon
    composite on white (polygon in PAGE coordinates), optional rotation.  (The reference's RGB->BGR->RGB
    round trip is the identity on the pixels.)"""
    crop = pil_crop(page_rgb, box)
    if polygon is not None:
        rel = [(int(p[0]) - box[0], int(p[1]) - box[1]) for p in polygon]
        mask = fill_poly_mask(crop.shape[0], crop.shape[1], rel)
        # cv2.add(bitwise_and(img, img, mask), bitwise_and(white, white, ~mask)): img inside, 255 outside
        crop = np.where(mask[..., None] != 0, crop, np.uint8(255))
    return rotate90(crop, rot)
