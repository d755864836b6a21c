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
