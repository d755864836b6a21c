"""ORACLE (test infrastructure) - numpy restatement of the reference crop
preprocessing, integer-exact.

Follows, step by step:
  * upstream ``MangaOcr.__call__``: ``img.convert("L").convert("RGB")``
    (Pillow ``ImagingConvert`` rgb2l: ``L = (19595 R + 38470 G + 7471 B + 0x8000) >> 16``);
  * ``ViTImageProcessorPil`` -> ``PilBackend.resize`` -> ``PIL.Image.resize((224,224), BILINEAR)``
    (transformers/image_processing_backends.py:528-577, transformers/image_transforms.py:313-381,
    Pillow ``src/libImaging/Resample.c``: ``precompute_coeffs``, ``normalize_coeffs_8bpc``,
    ``ImagingResampleHorizontal_8bpc`` then ``ImagingResampleVertical_8bpc``);
  * ``rescale`` (x * 1/255 in float64, cast to float32; transformers/image_transforms.py:89-124)
    and ``normalize`` ((x - 0.5) / 0.5 in float32; transformers/image_transforms.py:384-442).

``tests/test_oracle_preprocess.py`` pins every function here against Pillow and
``ViTImageProcessorPil`` themselves, bit for bit.
"""
from __future__ import annotations

import numpy as np

OUT = 224
PRECISION_BITS = 32 - 8 - 2


def rgb_to_l(rgb: np.ndarray) -> np.ndarray:
    """uint8 [H,W,3] RGB -> uint8 [H,W] luma, exactly as Pillow's convert("L")."""
    r = rgb[..., 0].astype(np.uint32)
    g = rgb[..., 1].astype(np.uint32)
    b = rgb[..., 2].astype(np.uint32)
    return ((19595 * r + 38470 * g + 7471 * b + 0x8000) >> 16).astype(np.uint8)


def resample_coeffs(in_size: int, out_size: int = OUT):
    """Pillow ``precompute_coeffs`` + ``normalize_coeffs_8bpc`` for the bilinear
    (triangle, support 1) filter.  Returns (xmin[out], count[out], k[out, ksize] int32)."""
    scale = float(in_size) / float(out_size)
    filterscale = scale if scale >= 1.0 else 1.0
    support = 1.0 * filterscale
    ksize = int(np.ceil(support)) * 2 + 1
    ss = 1.0 / filterscale
    xmin = np.zeros(out_size, np.int32)
    cnt = np.zeros(out_size, np.int32)
    kk = np.zeros((out_size, ksize), np.int32)
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        lo = int(center - support + 0.5)       # C double -> int truncation
        if lo < 0:
            lo = 0
        hi = int(center + support + 0.5)
        if hi > in_size:
            hi = in_size
        n = hi - lo
        w = np.zeros(n, np.float64)
        ww = 0.0
        for x in range(n):
            v = (x + lo - center + 0.5) * ss
            if v < 0.0:
                v = -v
            wv = 1.0 - v if v < 1.0 else 0.0
            w[x] = wv
            ww += wv
        if ww != 0.0:
            w = w / ww
        for x in range(n):
            p = w[x] * float(1 << PRECISION_BITS)
            kk[xx, x] = int(p - 0.5) if w[x] < 0 else int(p + 0.5)
        xmin[xx] = lo
        cnt[xx] = n
    return xmin, cnt, kk


def _resample_axis_last(img: np.ndarray, out_size: int) -> np.ndarray:
    """One 8-bit resampling pass along the last axis (uint8 in, uint8 out)."""
    in_size = img.shape[-1]
    xmin, cnt, kk = resample_coeffs(in_size, out_size)
    out = np.empty(img.shape[:-1] + (out_size,), np.uint8)
    src = img.astype(np.int64)
    for xx in range(out_size):
        n = int(cnt[xx])
        acc = (src[..., xmin[xx]:xmin[xx] + n] * kk[xx, :n].astype(np.int64)).sum(axis=-1)
        acc = (acc + (1 << (PRECISION_BITS - 1))) >> PRECISION_BITS
        out[..., xx] = np.clip(acc, 0, 255).astype(np.uint8)
    return out


def resize_l_224(gray: np.ndarray) -> np.ndarray:
    """uint8 [H,W] -> uint8 [224,224]: horizontal pass first (rounded to uint8),
    then vertical; a pass is skipped when that extent is already 224."""
    h, w = gray.shape
    tmp = gray if w == OUT else _resample_axis_last(gray, OUT)
    if h == OUT:
        return np.ascontiguousarray(tmp)
    return np.ascontiguousarray(_resample_axis_last(np.ascontiguousarray(tmp.T), OUT).T)


def normalize_lut() -> np.ndarray:
    """The 256 float32 values rescale+normalize can produce, computed with the
    reference's own numpy expression order."""
    v = np.arange(256, dtype=np.uint8)
    x = (v.astype(np.float64) * (1 / 255)).astype(np.float32)
    mean = np.array(0.5, dtype=np.float32)
    std = np.array(0.5, dtype=np.float32)
    return ((x - mean) / std).astype(np.float32)


def preprocess(rgb: np.ndarray):
    """uint8 [H,W,3] RGB -> (uint8 [224,224] resized luma, float32 [3,224,224] pixel_values)."""
    u8 = resize_l_224(rgb_to_l(rgb))
    plane = normalize_lut()[u8]
    return u8, np.ascontiguousarray(np.broadcast_to(plane, (3, OUT, OUT)))
