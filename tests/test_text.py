"""Host-side tail of MangaOcr.__call__ (manga_ocr_b200/text.py): tokenizer.decode + post_process."""
import numpy as np

from manga_ocr_b200.text import Vocab, h2z, ids_to_text, ids_to_texts, post_process


def test_vocab_shape_and_specials():
    v = Vocab.synthetic()
    assert len(v.tokens) == 6144 and len(set(v.tokens)) == 6144
    assert v.tokens[:5] == ["[PAD]", "[UNK]", "[CLS]", "[SEP]", "[MASK]"]
    assert v.decode([2, 5, 6, 3, 0, 0]) == f"{v.tokens[5]} {v.tokens[6]}"
    assert v.decode([2, 3]) == "" and v.decode([]) == ""
    assert v.decode([2, 99999, 3]) == "[UNK]"


def test_post_process_known_answers(golden_text):
    for src, want in golden_text["post_process"]:
        assert post_process(src) == want, src
    assert post_process("え … ?") == "え．．．？"      # h2z(ascii=True) widens the dots too
    assert post_process("・ ・ ・") == "．．．"
    assert post_process("a . b") == "ａ．ｂ"          # a single dot is kept and widened
    assert post_process("1 2 3") == "１２３"


def test_h2z_ascii_punctuation_character_by_character():
    """jaconv.h2z(ascii=True, digit=True) on every ASCII punctuation mark, expected output spelled out.
    Source: jaconv/conv_table.py (HALF_ASCII -> FULL_ASCII) as published, restated from memory because the package
    is absent offline: the four quote-like marks map to their JIS look-alikes, everything else is the U+FEE0 block
    shift.  This is the one place to look when checking the restatement against the real package."""
    expected = {
        "!": "！", '"': "”", "#": "＃", "$": "＄", "%": "％", "&": "＆", "'": "’", "(": "（", ")": "）", "*": "＊", "+": "＋",
        ",": "，", "-": "－", ".": "．", "/": "／", ":": "：", ";": "；", "<": "＜", "=": "＝", ">": "＞", "?": "？", "@": "＠",
        "[": "［", "\\": "￥", "]": "］", "^": "＾", "_": "＿", "`": "‘", "{": "｛", "|": "｜", "}": "｝", "~": "～",
    }
    for src, want in expected.items():
        assert h2z(src) == want, (src, h2z(src), want)
        assert post_process(f"あ {src} い") == f"あ{want}い"
    assert h2z("Az09") == "Ａｚ０９"
    # the oracle's independent restatement makes the same statement
    from oracle.reference_ocr import post_process as oracle_post_process
    for src in expected:
        assert oracle_post_process(f"x {src}") == post_process(f"x {src}")


def test_h2z_katakana_marks():
    assert h2z("ｶﾞｷﾞﾊﾟｱ") == "ガギパア"
    assert h2z("ｳﾞ") == "ヴ"
    assert h2z("漢字かな") == "漢字かな"


def test_ids_to_text_matches_oracle_strings(golden_text):
    v = Vocab.synthetic()
    for ids, text in zip(golden_text["ids"], golden_text["texts"]):
        assert ids_to_text(v, ids) == text


def test_vocab_file_roundtrip(tmp_path):
    v = Vocab.synthetic()
    p = tmp_path / "vocab.txt"
    p.write_text("\n".join(v.tokens) + "\n", encoding="utf-8")
    assert Vocab.from_file(str(p)).tokens == v.tokens


def test_batch_conversion_equals_per_row_conversion():
    """ids_to_texts (vectorised) == ids_to_text row by row, on rows that hit every path: plain characters,
    specials in the middle, dots / ellipsis / half-width voiced marks (context-dependent), ASCII (widened)."""
    base = Vocab.synthetic()
    v = Vocab(base.tokens[:-2] + ["･", "ab"])      # + a half-width middle dot and a multi-character token
    rng = np.random.default_rng(7)
    ids = rng.integers(0, len(v.tokens), size=(40, 300)).astype(np.int32)
    ids[:, 0] = 2
    ids[5, 10:] = 0
    ids[6, 1:] = 3
    tok = {t: i for i, t in enumerate(v.tokens)}
    ids[7, 1:6] = [tok["."], tok["."], tok["…"], tok["・"], tok["A"]]
    ids[8, 1:4] = [tok["ｶ"], tok["ﾞ"], tok["1"]]
    dotty = np.array([tok["."], tok["…"], tok["・"], tok["A"], tok["あ"], tok["･"]], np.int32)
    ids[20:30, 1:] = rng.choice(dotty[:5], size=(10, 299))      # dot rules only: the vectorised path with the dot fix-up
    ids[30:34, 1:] = rng.choice(dotty, size=(4, 299))           # half-width middle dot: must not join a dot run
    plain = np.array([i for i, t in enumerate(v.tokens) if i > 4 and 0x4E00 <= ord(t[0])], np.int32)
    ids[9:20, 1:] = rng.choice(plain, size=(11, 299))          # rows that stay on the vectorised path
    assert ids_to_texts(v, ids) == [ids_to_text(v, r) for r in ids]
    assert ids_to_texts(v, ids[:0]) == []
