"""The numpy restatement of the reference preprocessing (oracle/preprocess_np.py) against the
golden vectors produced by Pillow and transformers' ViTImageProcessorPil, and against Pillow
itself run live.  Bit-exact: this is integer work."""
import numpy as np
import pytest
from PIL import Image

from oracle import make_golden as G
from oracle import preprocess_np as P


def _rgb(a):
    if a.ndim == 2:
        return np.stack([a] * 3, -1)
    return a[..., :3]


def test_luma_matches_golden(golden_pre):
    for i, a in enumerate(G.pre_inputs()):
        l = P.rgb_to_l(_rgb(a))
        ref = golden_pre[f"l_{i}"]
        if l.size > 224 * 224:
            l = l[::7, ::7]
        assert np.array_equal(l, ref), i


def test_resize_matches_golden(golden_pre):
    for i, a in enumerate(G.pre_inputs()):
        u8, _ = P.preprocess(_rgb(a))
        assert np.array_equal(u8, golden_pre[f"u8_{i}"]), G.PRE_SHAPES[i]


def test_pixel_values_bit_exact(golden_pre):
    assert np.array_equal(P.normalize_lut().view(np.uint32), golden_pre["lut"].view(np.uint32))
    for i, a in enumerate(G.pre_inputs()[:3]):
        _, pv = P.preprocess(_rgb(a))
        assert np.array_equal(pv[0].view(np.uint32), golden_pre[f"pv_{i}"].view(np.uint32))
        assert np.array_equal(pv[0], pv[1]) and np.array_equal(pv[0], pv[2])


@pytest.mark.parametrize("shape", [(224, 224), (31, 500), (1600, 40), (3, 3), (225, 223), (448, 448), (112, 112), (1, 1), (1, 700)])
def test_resize_matches_live_pillow(shape):
    rng = np.random.default_rng(shape[0] * 7919 + shape[1])
    a = rng.integers(0, 256, size=shape + (3,), dtype=np.uint8)
    ref = np.asarray(Image.fromarray(a).convert("L").resize((224, 224), Image.BILINEAR))
    got, _ = P.preprocess(a)
    assert np.array_equal(got, ref)


def test_coefficients_sum_to_one():
    for n in (1, 2, 7, 224, 225, 1000, 4096):
        xmin, cnt, kk = P.resample_coeffs(n)
        assert (cnt >= 1).all() and (xmin >= 0).all() and (xmin + cnt <= n).all()
        s = kk.sum(axis=1)
        assert np.abs(s - (1 << P.PRECISION_BITS)).max() <= cnt.max()   # rounding of each tap


def test_torchvision_backend_delta_is_small_and_reported():
    """Informational ("preprocess-backend trap"): parity is defined against the PIL backend
    (`ViTImageProcessorPil`, what the reference's transformers 4.x stack ran); under transformers 5.x the name `ViTImageProcessor` is
    the torchvision backend, whose antialiased resize rounds differently.  The difference a user of a 5.x reference install would
    see against this engine: at most one uint8 step on a fraction of a percent of the pixels (DESIGN.md section 2 has the
    model-level effect: encoder rel-L2 2.4e-4, identical ids)."""
    import warnings
    pytest.importorskip("torchvision")
    from PIL import Image
    from transformers import ViTImageProcessor
    from transformers.models.vit.image_processing_pil_vit import ViTImageProcessorPil
    from manga_ocr_b200 import crops as C
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        pil = ViTImageProcessorPil(image_mean=[0.5] * 3, image_std=[0.5] * 3)
        tv = ViTImageProcessor(image_mean=[0.5] * 3, image_std=[0.5] * 3)
    if type(tv).__module__ == type(pil).__module__:
        pytest.skip("this transformers build has one backend only")
    total = differing = 0
    worst = 0.0
    for c in C.bubble_batch(6, seed=1002) + C.page_batch(6, seed=1003) + C.tall_batch(3, seed=1004):
        img = Image.fromarray(c).convert("L").convert("RGB")
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            a = pil(img, return_tensors="np").pixel_values[0]
            b = np.asarray(tv(img, return_tensors="pt").pixel_values[0])
        d = np.abs(a - b)
        total += d.size
        differing += int((d > 1e-6).sum())
        worst = max(worst, float(d.max()))
    print(f"torchvision vs PIL backend: {differing / total:.4%} of the pixel values differ, max {worst * 127.5:.2f} uint8 steps")
    assert worst <= 2.0 / 127.5 + 1e-6 and differing / total < 0.05
