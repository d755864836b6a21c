"""Crop staging: the oracle against the fixtures generated from the reference's own
library calls (PIL crop, cv2.fillPoly / bitwise_and / add / rotate - oracle/make_golden_staging.py)."""
import hashlib
import os

import numpy as np
import pytest

from oracle import crop_staging_np as S
from oracle import make_golden_staging as G


@pytest.fixture(scope="module")
def staging_kat():
    return np.load(os.path.join(os.path.dirname(__file__), "golden", "staging_kat.npz"))


def test_fill_poly_masks_match_opencv_fixtures(staging_kat):
    g = staging_kat
    n_masks = 0
    for i in range(int(g["n"])):
        if not int(g[f"has_poly{i}"]):
            continue
        box = [int(v) for v in g[f"box{i}"]]
        h, w = box[3] - box[1], box[2] - box[0]
        rel = [(int(p[0]) - box[0], int(p[1]) - box[1]) for p in g[f"poly{i}"]]
        assert np.array_equal(np.packbits(S.fill_poly_mask(h, w, rel) != 0), g[f"mask{i}"]), i
        n_masks += 1
    assert n_masks >= 20


def test_staged_images_match_reference_call_sequence(staging_kat):
    g = staging_kat
    page = G.page_rgb()
    for i in range(int(g["n"])):
        box = tuple(int(v) for v in g[f"box{i}"])
        poly = g[f"poly{i}"] if int(g[f"has_poly{i}"]) else None
        out = S.stage_region(page, box, poly, int(g[f"rot{i}"]))
        assert tuple(out.shape) == tuple(int(v) for v in g[f"shape{i}"]), i
        assert hashlib.sha256(np.ascontiguousarray(out).tobytes()).digest() == g[f"sha{i}"].tobytes(), i


def test_fill_poly_degenerate_inputs():
    assert not S.fill_poly_mask(5, 7, []).any()
    m = S.fill_poly_mask(5, 7, [(2, 3)])                       # one point: just that pixel
    assert m.sum() == 255 and m[3, 2] == 255
    m = S.fill_poly_mask(5, 7, [(1, 1), (5, 1)])               # two points: the segment
    assert np.array_equal(np.nonzero(m[1])[0], np.arange(1, 6)) and m.sum() == 5 * 255
    assert not S.fill_poly_mask(4, 4, [(-9, -9), (-5, -9), (-5, -5)]).any()     # entirely outside
    full = S.fill_poly_mask(4, 6, [(-3, -3), (20, -3), (20, 20), (-3, 20)])     # covers everything
    assert (full == 255).all()


def test_opencv_agrees_when_installed():
    """Live check against cv2 where it is installed (the build container); the fixtures above carry the pin."""
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(3)
    for t in range(300):
        h, w, n = int(rng.integers(1, 60)), int(rng.integers(1, 60)), int(rng.integers(1, 20))
        lo, hi = (-20, 20) if t % 2 else (0, 1)
        pts = np.stack([rng.integers(lo, w + hi, n), rng.integers(lo, h + hi, n)], 1).astype(np.int32)
        ref = np.zeros((h, w), np.uint8)
        cv2.fillPoly(ref, [pts], 255)
        assert np.array_equal(S.fill_poly_mask(h, w, pts), ref), (h, w, pts.tolist())


def test_region_from_qt_reproduces_reference_numbers():
    from manga_ocr_b200.engine import ROT_CCW, ROT_CW, ROT_NONE, Region
    r = Region.from_qt((10, 20, 100, 50), None, "Vertical")          # crop is 99 x 49: wider than tall -> clockwise
    assert r.box == (10, 20, 109, 69) and r.rotate == ROT_CW
    assert Region.from_qt((10, 20, 50, 100), None, "Horizontal").rotate == ROT_CCW
    assert Region.from_qt((10, 20, 50, 100), None, "Vertical").rotate == ROT_NONE
    assert Region.from_qt((10, 20, 50, 100), None, None).rotate == ROT_NONE
    assert S.rotation_for("Vertical", 49, 99) == ROT_CW and S.rotation_for("Horizontal", 99, 49) == ROT_CCW


def test_page_generator_is_deterministic():
    from manga_ocr_b200 import crops as C
    p1, s1 = C.page_with_selections(6)
    p2, s2 = C.page_with_selections(6)
    assert np.array_equal(p1, p2) and len(s1) == 6
    assert all((a[1] is None) == (b[1] is None) and a[0] == b[0] for a, b in zip(s1, s2))
