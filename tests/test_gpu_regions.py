"""Region staging on the B200, through the C ABI: page + (box, polygon, rotation)
-> the same uint8 224x224 plane, and the same token ids, as the reference's host-side crop / cv2.fillPoly
composite / cv2.rotate followed by the plain crop path.  Integer work: bit-exact."""
import os

import numpy as np
import pytest

from manga_ocr_b200 import crops as C
from oracle import crop_staging_np as S
from oracle import make_golden_staging as G
from oracle import preprocess_np as P

pytestmark = pytest.mark.gpu


def _regions_from_golden(g):
    from manga_ocr_b200.engine import Region
    out = []
    for i in range(int(g["n"])):
        poly = g[f"poly{i}"] if int(g[f"has_poly{i}"]) else None
        out.append(Region([int(v) for v in g[f"box{i}"]], poly, int(g[f"rot{i}"])))
    return out


@pytest.fixture(scope="module")
def staging_kat():
    return np.load(os.path.join(os.path.dirname(__file__), "golden", "staging_kat.npz"))


def _pixels(engine, page, regions, order=0):
    from manga_ocr_b200.engine import TAP_PIXELS
    engine.set_taps(TAP_PIXELS)
    engine.stage_regions(page, regions, order)
    engine.preprocess()
    return engine.pixels_u8()


def test_device_masks_match_opencv_fixtures(engine8, staging_kat):
    g = staging_kat
    page = G.page_rgb()
    regions = _regions_from_golden(g)
    for lo in range(0, len(regions), 8):
        chunk = regions[lo:lo + 8]
        engine8.stage_regions(page, chunk)
        for j, r in enumerate(chunk):
            if r.polygon is None:
                continue
            h, w = r.box[3] - r.box[1], r.box[2] - r.box[0]
            got = engine8.region_mask(j, (h, w))
            assert set(np.unique(got)) <= {0, 255}
            assert np.array_equal(np.packbits(got != 0), g[f"mask{lo + j}"]), lo + j


def test_device_masks_match_oracle_on_random_polygons(engine8):
    from manga_ocr_b200.engine import Region
    rng = np.random.default_rng(21)
    page = rng.integers(0, 256, (300, 280, 3), dtype=np.uint8)
    for rnd in range(12):
        regions = []
        for k in range(8):
            w, h = int(rng.integers(1, 200)), int(rng.integers(1, 200))
            x, y = int(rng.integers(-20, 250)), int(rng.integers(-20, 270))
            n = int(rng.integers(1, 64)) if k else int(rng.integers(200, 1025))
            spread = 30 if (rnd + k) % 2 else 0                  # every other polygon leaves its box
            pts = np.stack([x + rng.integers(-spread, w + spread + 1, n), y + rng.integers(-spread, h + spread + 1, n)], 1)
            regions.append(Region((x, y, x + w, y + h), pts.astype(np.int32), 0))
        engine8.stage_regions(page, regions)
        for j, r in enumerate(regions):
            h, w = r.box[3] - r.box[1], r.box[2] - r.box[0]
            want = S.fill_poly_mask(h, w, r.polygon - np.array(r.box[:2]))
            assert np.array_equal(engine8.region_mask(j, (h, w)), want), (rnd, j, r.box, len(r.polygon))


def test_staged_planes_bit_exact_golden_selections(engine8, staging_kat):
    """uint8 224x224 plane of every golden selection == Pillow-exact preprocess of the reference-staged image."""
    page = G.page_rgb()
    regions = _regions_from_golden(staging_kat)
    for lo in range(0, len(regions), 8):
        chunk = regions[lo:lo + 8]
        u8 = _pixels(engine8, page, chunk)
        for j, r in enumerate(chunk):
            staged = S.stage_region(page, r.box, r.polygon, r.rotate)
            assert np.array_equal(u8[j], P.preprocess(staged)[0]), (lo + j, r.box, r.rotate)


def test_staged_planes_page_config_and_layouts(engine8):
    from manga_ocr_b200.engine import Region
    page, sels = C.page_with_selections(16)
    regions = [Region.from_qt(rect, poly, orient) for rect, poly, orient in sels]
    assert any(r.rotate for r in regions) and any(r.polygon is None for r in regions)
    want = [P.preprocess(S.stage_region(page, r.box, r.polygon, r.rotate))[0] for r in regions]
    for lo in range(0, 16, 8):
        u8 = _pixels(engine8, page, regions[lo:lo + 8])
        for j in range(8):
            assert np.array_equal(u8[j], want[lo + j]), lo + j
    # BGR page (the app's cv2 side), a padded row stride, and a 1-channel page
    u8 = _pixels(engine8, np.ascontiguousarray(page[..., ::-1]), regions[:8], order=1)
    assert all(np.array_equal(u8[j], want[j]) for j in range(8))
    padded = np.zeros((page.shape[0], page.shape[1] + 37, 3), np.uint8)
    padded[:, :page.shape[1]] = page
    u8 = _pixels(engine8, padded[:, :page.shape[1]], regions[:8])
    assert all(np.array_equal(u8[j], want[j]) for j in range(8))
    gray = P.rgb_to_l(page)
    u8 = _pixels(engine8, gray, regions[:8])
    for j, r in enumerate(regions[:8]):
        staged = S.stage_region(gray[..., None], r.box, r.polygon, r.rotate)[..., 0]
        assert np.array_equal(u8[j], P.resize_l_224(staged)), j


def test_recognize_regions_equals_recognize_of_host_staged_crops(engine8):
    from manga_ocr_b200.engine import Region
    page, sels = C.page_with_selections(19, seed=77)
    regions = [Region.from_qt(rect, poly, orient) for rect, poly, orient in sels]
    ids_r, lens_r = engine8.recognize_regions(page, regions)                     # 19 > max_batch 8: chunked
    crops = [S.stage_region(page, r.box, r.polygon, r.rotate) for r in regions]
    ids_c, lens_c = engine8.recognize(crops)
    assert np.array_equal(ids_r, ids_c) and np.array_equal(lens_r, lens_c)
    assert ids_r.shape == (19, 24) and (ids_r[:, 0] == 2).all()


def test_region_errors_are_reported_and_handle_survives(engine8):
    from manga_ocr_b200.engine import MocrError, Region
    page = np.zeros((50, 60, 3), np.uint8)
    with pytest.raises(MocrError) as e:
        engine8.stage_regions(page, [Region((5, 5, 5, 20))])                     # empty box
    assert e.value.code == -1
    with pytest.raises(MocrError) as e:
        engine8.stage_regions(page, [Region((0, 0, 10, 10), None, 3)])           # bad rotation code
    assert e.value.code == -1
    with pytest.raises(MocrError) as e:
        engine8.stage_regions(page, [Region((0, 0, 10, 10), np.zeros((1025, 2), np.int32))])
    assert e.value.code == -4
    with pytest.raises(MocrError) as e:
        engine8.stage_regions(page, [Region((0, 0, 10, 10))] * 9)                # over max_batch
    assert e.value.code == -4
    engine8.stage_regions(page, [Region((0, 0, 10, 10))])
    with pytest.raises(MocrError):
        engine8.region_mask(0, (10, 10))                                         # no polygon on that region
    ids, _ = engine8.recognize_regions(page, [Region((-5, -5, 70, 60), [(0, 0), (59, 0), (30, 49)], 1)])
    assert ids.shape == (1, 24)


def test_manga_ocr_recognize_regions_strings(weights0):
    """The front-end call: page + Region.from_qt(...) selections -> the strings of the host-staged crops."""
    from manga_ocr_b200.engine import Region
    from manga_ocr_b200.ocr import MangaOcr
    page, sels = C.page_with_selections(11, seed=5)
    regions = [Region.from_qt(rect, poly, orient) for rect, poly, orient in sels]
    ocr = MangaOcr(weights=weights0, max_batch=8, max_length=16, warmup=False)
    try:
        got = ocr.recognize_regions(page, regions)
        want = ocr.recognize_batch([S.stage_region(page, r.box, r.polygon, r.rotate) for r in regions])
        assert got == want and len(got) == 11 and all(isinstance(t, str) for t in got)
        from PIL import Image
        assert ocr.recognize_regions(Image.fromarray(page), regions[:3]) == want[:3]
        assert ocr.recognize_regions(page, []) == []
    finally:
        ocr.close()
