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


def test_decode_glues_wordpiece_continuations_like_transformers():
    """tokenizer.decode(skip_special_tokens=True) of BertJapaneseTokenizer is `" ".join(tokens).replace(" ##", "").strip()`
    (transformers/models/bert_japanese/tokenization_bert_japanese.py:256-261): a token that starts with "##" is glued to
    its predecessor, a leading one keeps its hashes.  Checked against transformers' own function, on a vocabulary that has
    such tokens (a real vocab.txt may), through every conversion path of the product."""
    from transformers.models.bert_japanese.tokenization_bert_japanese import BertJapaneseTokenizer

    class Char:
        subword_tokenizer_type = "character"

    from oracle.reference_ocr import post_process as oracle_post_process
    base = Vocab.synthetic()
    v = Vocab(base.tokens[:-6] + ["##あ", "##b", "##", "###c", "#x", "a b"])
    tok = {t: i for i, t in enumerate(v.tokens)}
    rows = [
        ["あ", "##あ", "い"], ["##あ", "い"], ["a", "###c"], ["a", "##"], ["#", "#", "x"], ["#", "#x"], ["a", "##b", ".", ".", "##あ"],
        ["a b", "##b"], ["[CLS]", "##b", "[SEP]"], ["ｶ", "##あ", "ﾞ"], ["##", "##"], [],
    ]
    T = 8
    ids = np.zeros((len(rows), T), np.int32)
    for r, toks in enumerate(rows):
        ids[r, :len(toks)] = [tok[t] for t in toks]
    want = []
    for r, toks in enumerate(rows):
        kept = [v.tokens[i] for i in ids[r] if i not in v.special_ids]
        decoded = BertJapaneseTokenizer.convert_tokens_to_string(Char(), kept)
        assert v.decode(ids[r]) == decoded == v.decode(ids[r].tolist())
        want.append(oracle_post_process(decoded))
    assert [ids_to_text(v, r) for r in ids] == want
    assert ids_to_texts(v, ids) == want
    assert want[0] == "ああい" and want[1] == "＃＃あい" and want[2] == "a＃c".replace("a", "ａ").replace("c", "ｃ")
    # rows without such tokens stay on the vectorised path and are unchanged by the rule
    rng = np.random.default_rng(3)
    plain = rng.integers(5, len(base.tokens) - 6, size=(6, 40)).astype(np.int32)
    assert ids_to_texts(v, plain) == ids_to_texts(base, plain) == [ids_to_text(base, r) for r in plain]


def test_vocab_file_and_decode_equal_the_real_tokenizer(tmp_path):
    """vocab.txt -> id table -> tokenizer.decode(skip_special_tokens=True) -> post_process, the product's against transformers' own
    BertJapaneseTokenizer built on the same file (word tokenizer "basic": decode never runs it, and MeCab is absent offline).  The
    file has what a real one may have: "##" continuation tokens, a multi-character token, a token on two lines (only the last
    line keeps it, the earlier id decodes to "[UNK]" and is NOT special), a space token, an empty line, out-of-range ids."""
    import warnings
    from transformers.models.bert_japanese.tokenization_bert_japanese import BertJapaneseTokenizer
    from oracle.reference_ocr import post_process as oracle_post_process
    base = Vocab.synthetic().tokens[:300]
    lines = base + ["##あ", "##b", "a b", "x", "x", " ", "", "'", "n't", ".", "[MASK]", "ｶ", "ﾞ"]
    path = tmp_path / "vocab.txt"
    path.write_text("\n".join(lines) + "\n", encoding="utf-8")
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        tk = BertJapaneseTokenizer(str(path), word_tokenizer_type="basic", subword_tokenizer_type="character")
    v = Vocab.from_file(str(path))
    assert len(v.tokens) == len(lines)
    assert v.special_ids == frozenset(tk.all_special_ids)
    assert [v.tokens[i] for i in range(len(lines))] == tk.convert_ids_to_tokens(list(range(len(lines))))
    rng = np.random.default_rng(11)
    ids = rng.integers(0, len(lines), size=(60, 40)).astype(np.int32)
    ids[:, 0] = 2
    ids[3, 5:] = 0
    ids[4, 1:9] = [300, 301, 302, 303, 304, 305, 306, 307]
    ids[5, 1:4] = [311, 312, 300]                      # ｶ ﾞ ##あ
    ids[6, 1:3] = [4, 310]                             # the first [MASK] line lost the token to the second one
    want = [oracle_post_process(tk.decode(row.tolist(), skip_special_tokens=True)) for row in ids]
    for row, w in zip(ids, want):
        assert v.decode(row) == tk.decode(row.tolist(), skip_special_tokens=True)
        assert ids_to_text(v, row) == w
    assert ids_to_texts(v, ids) == want
    wide = ids.copy()
    wide[7, 3] = len(lines) + 5                        # an id beyond the table: "[UNK]", kept (the check is by id)
    assert ids_to_text(v, wide[7]) == oracle_post_process(tk.decode(wide[7].tolist(), skip_special_tokens=True))
    assert ids_to_texts(v, wide)[7] == ids_to_text(v, wide[7])


def test_h2z_tables_against_unicode_compatibility_mappings():
    """An authority that is neither restatement: Unicode's own compatibility mappings (unicodedata, NFKC).  Half-width katakana
    letters and punctuation widen to exactly their NFKC form; a letter followed by a (semi-)voiced mark composes to the NFKC
    composition for every pair jaconv converts; widened ASCII narrows back to the original under NFKC - except the four look-alikes
    jaconv's table substitutes (” ’ ￥ ‘), which is what sets it apart from a plain block shift."""
    import unicodedata
    from manga_ocr_b200.text import H2Z_ASCII_SPECIAL, h2z
    from oracle.reference_ocr import _h2z as oracle_h2z

    def nfkc(t):
        return unicodedata.normalize("NFKC", t)
    for cp in list(range(0xFF66, 0xFF9E)) + [0xFF61, 0xFF62, 0xFF63, 0xFF64, 0xFF65]:       # ｦ..ﾝ and ｡｢｣､･
        ch = chr(cp)
        assert h2z(ch) == nfkc(ch) == oracle_h2z(ch), hex(cp)
    voiced = "ｶｷｸｹｺｻｼｽｾｿﾀﾁﾂﾃﾄﾊﾋﾌﾍﾎｳ"
    for base in voiced:
        assert h2z(base + "ﾞ") == nfkc(base + "ﾞ") == oracle_h2z(base + "ﾞ") and len(h2z(base + "ﾞ")) == 1, base
    for base in "ﾊﾋﾌﾍﾎ":
        assert h2z(base + "ﾟ") == nfkc(base + "ﾟ") == oracle_h2z(base + "ﾟ") and len(h2z(base + "ﾟ")) == 1, base
    assert h2z("ｱﾞ") == "ア゛" == oracle_h2z("ｱﾞ")             # no composition outside jaconv's list: the mark widens on its own
    for cp in range(0x21, 0x7F):
        ch = chr(cp)
        if ch in H2Z_ASCII_SPECIAL:
            assert nfkc(h2z(ch)) != ch                          # the look-alikes are not compatibility forms of the ASCII character
        else:
            assert nfkc(h2z(ch)) == ch and h2z(ch) == chr(cp + 0xFEE0), hex(cp)
