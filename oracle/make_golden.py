"""ORACLE (test infrastructure) - generates the golden fixtures under tests/golden/ by running
the reference's own dependencies in the build container (transformers 5.5.0, torch 2.11.0 CPU
fp32, Pillow 12.2.0).  The reference repository holds no tests or vectors for this path
(SURVEY.md section 4), so these library outputs are what pins the oracle:

  preprocess_kat.npz  Pillow ``convert("L")`` / ``resize((224,224), BILINEAR)`` and
                      ``ViTImageProcessorPil`` pixel_values on seeded random crops
  model_kat.npz       ``VisionEncoderDecoderModel`` (random-init, numpy PCG64 seed 0) encoder
                      hidden rows, greedy ids, per-step logits (strided + top-2) on seeded crops
  text_kat.json       ids -> string through the oracle wrapper, and post_process known answers

Run:  python -m oracle.make_golden        (writes tests/golden/*, ~1 MB)
"""
from __future__ import annotations

import json
import os

import numpy as np
from PIL import Image

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "tests", "golden")

# (H, W, channels) of the seeded random-noise crops; inputs are regenerated from the seed.
PRE_SHAPES = [(224, 224, 3), (100, 60, 3), (480, 120, 3), (37, 301, 3), (64, 64, 3), (223, 225, 3),
              (17, 900, 3), (224, 100, 3), (2, 5, 3), (1000, 300, 3), (300, 224, 3), (50, 50, 1), (90, 40, 4)]
PRE_SEED = 4242
MODEL_T = 12
ENC_ROWS = [0, 1, 2, 97, 98, 150, 195, 196]
LOGIT_STRIDE = 16


def pre_inputs():
    rng = np.random.default_rng(PRE_SEED)
    out = []
    for h, w, c in PRE_SHAPES:
        shape = (h, w) if c == 1 else (h, w, c)
        a = rng.integers(0, 256, size=shape, dtype=np.uint8)
        # half of them smooth (low-frequency) so the antialias taps matter in both regimes
        if (h + w) % 2 == 0:
            a = (a.astype(np.float32) * 0.25 + 96).astype(np.uint8)
        out.append(a)
    return out


def model_inputs():
    from manga_ocr_b200 import crops as C
    rng = np.random.default_rng(77)
    return C.bubble_batch(2, seed=1002) + [C.make_crop(rng, 700, 60, tint=True)]


def main() -> None:
    from transformers.models.vit.image_processing_pil_vit import ViTImageProcessorPil
    from manga_ocr_b200 import weights as W
    from manga_ocr_b200.text import Vocab
    from oracle.reference_ocr import ReferenceMangaOcr, post_process

    os.makedirs(OUT, exist_ok=True)
    proc = ViTImageProcessorPil()

    # ---- preprocessing known answers straight from Pillow / transformers
    pre = {}
    for i, a in enumerate(pre_inputs()):
        img = Image.fromarray(a)
        if img.mode == "RGBA":
            ref_l = np.asarray(img.convert("L"))
        else:
            ref_l = np.asarray(img.convert("L"))
        rgb = img.convert("L").convert("RGB")
        pre[f"l_{i}"] = ref_l if ref_l.size <= 224 * 224 else ref_l[::7, ::7]
        pre[f"u8_{i}"] = np.asarray(rgb.resize((224, 224), Image.BILINEAR))[..., 0]
        pv = proc(rgb, return_tensors="np").pixel_values[0]
        assert np.array_equal(pv[0], pv[1]) and np.array_equal(pv[0], pv[2])
        if i < 3:
            pre[f"pv_{i}"] = pv[0].astype(np.float32)
    ramp = np.arange(256, dtype=np.uint8).repeat(196).reshape(224, 224)
    pv = proc(Image.fromarray(np.stack([ramp] * 3, -1)), return_tensors="np").pixel_values[0, 0]
    pre["lut"] = pv.reshape(256, 196)[:, 0].astype(np.float32)
    np.savez_compressed(os.path.join(OUT, "preprocess_kat.npz"), **pre)

    # ---- model known answers from transformers' VisionEncoderDecoderModel
    weights = W.random_init(0)
    vocab = Vocab.synthetic()
    ocr = ReferenceMangaOcr(weights, vocab.tokens, max_length=MODEL_T)
    crops = model_inputs()
    enc = ocr.encoder_hidden(crops)
    ids, logits = ocr.generate_batch(crops, max_length=MODEL_T)
    top2_idx = np.argsort(logits, axis=-1)[..., -2:][..., ::-1]
    top2_val = np.take_along_axis(logits, top2_idx, axis=-1)
    np.savez_compressed(
        os.path.join(OUT, "model_kat.npz"),
        enc_rows=enc[:, ENC_ROWS].astype(np.float32), enc_mean=enc.mean(axis=(1, 2)), enc_std=enc.std(axis=(1, 2)),
        enc_abs_sum=np.abs(enc).sum(axis=(1, 2)),
        ids=ids.astype(np.int32), logits_strided=logits[..., ::LOGIT_STRIDE].astype(np.float32),
        top2_idx=top2_idx.astype(np.int32), top2_val=top2_val.astype(np.float32))

    # ---- strings
    texts = [ocr(Image.fromarray(c)) for c in crops]
    kat = {
        "ids": ids.tolist(), "texts": texts,
        "post_process": [[s, post_process(s)] for s in [
            "こ ん に ち は", "え … ?", "・ ・ ・ ま さ か", "A B C 1 2 3 !", "ｶ ﾞ ｷ ﾞ ﾊ ﾟ ｱ", "そ う . . だ ね", "a . b", "  ", ""]],
    }
    with open(os.path.join(OUT, "text_kat.json"), "w", encoding="utf-8") as f:
        json.dump(kat, f, ensure_ascii=False, indent=1)
    for fn in sorted(os.listdir(OUT)):
        print(fn, os.path.getsize(os.path.join(OUT, fn)))


if __name__ == "__main__":
    main()
