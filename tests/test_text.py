"""Host-side tail of MangaOcr.__call__ (manga_ocr_b200/text.py): tokenizer.decode + post_process."""
from manga_ocr_b200.text import Vocab, h2z, ids_to_text, post_process


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
