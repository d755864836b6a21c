"""ORACLE (test infrastructure) - the reference ``MangaOcr(img) -> str`` path on
CPU, built from the reference's own dependencies.

The reference app imports ``MangaOcr`` from the pip package ``manga-ocr``
(reference/src/core/config.py:433), constructs it once
(reference/src/ui/main_window.py:3394) and calls it per crop
(reference/src/ui/main_window.py:9800-9801).  That package (unpinned, absent
offline) is a thin wrapper over ``transformers``; this module restates the
wrapper and calls the very transformers / Pillow classes
it calls:

  * ``img.convert("L").convert("RGB")``                        (Pillow)
  * ``ViTImageProcessorPil``  (transformers/models/vit/image_processing_pil_vit.py:20-27)
  * ``VisionEncoderDecoderModel.generate(x[None], max_length=300)`` greedy
    (transformers/generation/utils.py:2658-2810; ViT: models/vit/modeling_vit.py:428-458;
    BERT decoder: models/bert/modeling_bert.py:856-910)
  * ``tokenizer.decode(skip_special_tokens=True)``: transformers' own ``BertJapaneseTokenizer`` (``real_tokenizer``; it
    constructs offline with the basic word tokenizer, which decode never runs);
  * ``post_process`` + ``jaconv.h2z``: RESTATED from their published behaviour (both packages are absent offline) -
    this last step is "parity unpinned" against upstream.

fp32, eager, batch 1 per call exactly like the reference.  Nothing in
``manga_ocr_b200`` imports this file.
"""
from __future__ import annotations

import re
from typing import Dict, Iterable, List, Optional, Sequence

import numpy as np
import torch
from PIL import Image

MAX_LENGTH = 300
PAD_ID, UNK_ID, CLS_ID, SEP_ID, MASK_ID = 0, 1, 2, 3, 4


def build_model(weights: Dict[str, np.ndarray], tie_lm_head: Optional[bool] = None):
    """``VisionEncoderDecoderModel`` of the manga-ocr-base architecture with the
    given weights (reference state_dict names) loaded, eval mode, fp32."""
    from transformers import (BertConfig, ViTConfig, VisionEncoderDecoderConfig,
                              VisionEncoderDecoderModel)

    wemb = weights["decoder.bert.embeddings.word_embeddings.weight"]
    head = weights.get("decoder.cls.predictions.decoder.weight", wemb)
    if tie_lm_head is None:
        tie_lm_head = head is wemb or np.array_equal(head, wemb)
    enc = ViTConfig()
    dec = BertConfig(vocab_size=6144, num_hidden_layers=2, is_decoder=True,
                     add_cross_attention=True, pad_token_id=PAD_ID,
                     tie_word_embeddings=bool(tie_lm_head))
    cfg = VisionEncoderDecoderConfig.from_encoder_decoder_configs(enc, dec)
    cfg.decoder_start_token_id = CLS_ID
    cfg.eos_token_id = SEP_ID
    cfg.pad_token_id = PAD_ID
    cfg.tie_word_embeddings = False
    model = VisionEncoderDecoderModel(cfg).eval()
    sd = {k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in weights.items()}
    sd["decoder.cls.predictions.decoder.weight"] = torch.from_numpy(np.ascontiguousarray(head))
    sd["decoder.cls.predictions.decoder.bias"] = sd["decoder.cls.predictions.bias"]
    own = model.state_dict()
    for k in own:
        if k not in sd:
            if k.startswith("encoder.pooler."):      # dead on this path; keep its init
                sd[k] = own[k]
            else:
                raise KeyError(f"weights lack {k}")
    model.load_state_dict(sd, strict=True)
    model.generation_config.decoder_start_token_id = CLS_ID
    model.generation_config.eos_token_id = SEP_ID
    model.generation_config.pad_token_id = PAD_ID
    return model


def to_pixel_values(img: Image.Image) -> torch.Tensor:
    """``processor(img.convert("L").convert("RGB"), return_tensors="pt").pixel_values.squeeze()``."""
    from transformers.models.vit.image_processing_pil_vit import ViTImageProcessorPil
    global _PROC
    try:
        proc = _PROC
    except NameError:
        proc = _PROC = ViTImageProcessorPil()
    img = img.convert("L").convert("RGB")
    return proc(img, return_tensors="pt").pixel_values.squeeze()


# ---- tokenizer.decode + post_process, restated independently of the product ----

# jaconv/conv_table.py as published (remembered; the package is absent offline - "parity unpinned" for this row):
# HALF_ASCII '!"#$%&\'()*+,-./:;<=>?@[\\]^_`{|}~' -> FULL_ASCII with ” ’ ￥ ‘ in place of the block-shifted ＂ ＇ ＼ ｀
_JACONV_ASCII_SPECIAL = {chr(0x22): chr(0x201D), chr(0x27): chr(0x2019), chr(0x5C): chr(0xFFE5), chr(0x60): chr(0x2018)}


def _h2z(text: str) -> str:
    hw = "ｦｧｨｩｪｫｬｭｮｯｰｱｲｳｴｵｶｷｸｹｺｻｼｽｾｿﾀﾁﾂﾃﾄﾅﾆﾇﾈﾉﾊﾋﾌﾍﾎﾏﾐﾑﾒﾓﾔﾕﾖﾗﾘﾙﾚﾛﾜﾝ｡｢｣､･ﾞﾟ"
    fw = "ヲァィゥェォャュョッーアイウエオカキクケコサシスセソタチツテトナニヌネノハヒフヘホマミムメモヤユヨラリルレロワン。「」、・゛゜"
    voiced = dict(zip("ｶｷｸｹｺｻｼｽｾｿﾀﾁﾂﾃﾄﾊﾋﾌﾍﾎｳ", "ガギグゲゴザジズゼゾダヂヅデドバビブベボヴ"))
    semi = dict(zip("ﾊﾋﾌﾍﾎ", "パピプペポ"))
    single = dict(zip(hw, fw))
    out: List[str] = []
    i = 0
    while i < len(text):
        ch = text[i]
        nxt = text[i + 1] if i + 1 < len(text) else ""
        if nxt == "ﾞ" and ch in voiced:
            out.append(voiced[ch]); i += 2; continue
        if nxt == "ﾟ" and ch in semi:
            out.append(semi[ch]); i += 2; continue
        if ch in single:
            out.append(single[ch])
        elif ch in _JACONV_ASCII_SPECIAL:           # jaconv's FULL_ASCII row is not the plain block shift for these four
            out.append(_JACONV_ASCII_SPECIAL[ch])
        elif 0x21 <= ord(ch) <= 0x7E:
            out.append(chr(ord(ch) + 0xFEE0))
        elif ch == " ":
            out.append("　")
        else:
            out.append(ch)
        i += 1
    return "".join(out)


def post_process(text: str) -> str:
    text = "".join(text.split())
    text = text.replace("…", "...")
    text = re.sub("[・.]{2,}", lambda x: (x.end() - x.start()) * ".", text)
    return _h2z(text)


_TOKENIZERS: Dict[int, object] = {}


def real_tokenizer(tokens: Sequence[str]):
    """transformers' OWN ``BertJapaneseTokenizer`` over ``tokens`` (written out as a vocab.txt and read back by the class, as
    for a checkpoint).  ``AutoTokenizer`` cannot build the checkpoint's tokenizer offline - its word tokenizer is MeCab and
    needs fugashi / unidic - but ``decode`` never runs the word tokenizer: with
    ``word_tokenizer_type="basic"`` the class constructs here, and id -> text is the library's own code path
    (``convert_ids_to_tokens(skip_special_tokens=True)`` -> ``convert_tokens_to_string``, tokenization_bert_japanese.py:250-261)."""
    import os
    import tempfile
    from transformers.models.bert_japanese.tokenization_bert_japanese import BertJapaneseTokenizer
    key = hash(tuple(tokens))
    if key not in _TOKENIZERS:
        with tempfile.TemporaryDirectory() as d:
            path = os.path.join(d, "vocab.txt")
            with open(path, "w", encoding="utf-8", newline="\n") as f:
                f.write("\n".join(tokens) + "\n")
            _TOKENIZERS[key] = BertJapaneseTokenizer(path, word_tokenizer_type="basic", subword_tokenizer_type="character")
    return _TOKENIZERS[key]


def decode_ids(tokens: Sequence[str], ids: Iterable[int]) -> str:
    """``tokenizer.decode(ids, skip_special_tokens=True)`` - called on the real tokenizer class, not restated."""
    return real_tokenizer(tokens).decode([int(i) for i in ids], skip_special_tokens=True)


class ReferenceMangaOcr:
    """CPU oracle with the reference's call signature: ``ocr(img_or_path) -> str``."""

    def __init__(self, weights: Dict[str, np.ndarray], tokens: Sequence[str], max_length: int = MAX_LENGTH):
        self.model = build_model(weights)
        self.tokens = list(tokens)
        self.max_length = max_length

    def pixel_values(self, img) -> torch.Tensor:
        if isinstance(img, np.ndarray):
            img = Image.fromarray(img)
        return to_pixel_values(img)

    @torch.no_grad()
    def generate_ids(self, img) -> np.ndarray:
        x = self.pixel_values(img)
        ids = self.model.generate(x[None], max_length=self.max_length, num_beams=1, do_sample=False)[0]
        return ids.cpu().numpy()

    def __call__(self, img_or_path) -> str:
        if isinstance(img_or_path, (str,)) or hasattr(img_or_path, "__fspath__"):
            img = Image.open(img_or_path)
        elif isinstance(img_or_path, Image.Image):
            img = img_or_path
        else:
            raise ValueError(f"img_or_path must be a path or PIL.Image, instead got: {img_or_path}")
        ids = self.generate_ids(img)
        return post_process(decode_ids(self.tokens, ids))

    # ---- taps used by the parity tests ----
    @torch.no_grad()
    def encoder_hidden(self, imgs: Sequence) -> np.ndarray:
        x = torch.stack([self.pixel_values(i) for i in imgs])
        return self.model.encoder(pixel_values=x).last_hidden_state.numpy()

    @torch.no_grad()
    def generate_batch(self, imgs: Sequence, max_length: Optional[int] = None):
        """Batched greedy decode; returns (ids [B,T] padded with PAD, per-step logits [B,T-1,V])."""
        x = torch.stack([self.pixel_values(i) for i in imgs])
        out = self.model.generate(x, max_length=max_length or self.max_length, num_beams=1, do_sample=False,
                                  output_logits=True, return_dict_in_generate=True)
        logits = torch.stack(out.logits, dim=1).float().numpy()
        return out.sequences.numpy(), logits

    @torch.no_grad()
    def teacher_forced_logits(self, imgs: Sequence, ids: np.ndarray) -> np.ndarray:
        """One forward with ``decoder_input_ids = ids[:, :-1]`` -> logits [B, T-1, V]."""
        x = torch.stack([self.pixel_values(i) for i in imgs])
        dec_in = torch.from_numpy(np.asarray(ids[:, :-1], dtype=np.int64))
        return self.model(pixel_values=x, decoder_input_ids=dec_in).logits.float().numpy()
