"""Token ids -> string: the host-side tail of ``MangaOcr.__call__``.

Follows the reference path:
  * ``tokenizer.decode(ids, skip_special_tokens=True)`` of a character-level
    BERT-Japanese tokenizer = drop special ids, ``" ".join(tokens)``
    (transformers/models/bert_japanese/tokenization_bert_japanese.py:256-261);
  * upstream ``manga_ocr.ocr.post_process``: remove all whitespace, ``…`` ->
    ``...``, runs of ``[・.]{2,}`` -> same-length run of ``.``, then
    ``jaconv.h2z(text, ascii=True, digit=True)``.

Neither the ``manga-ocr`` package nor ``jaconv`` nor the real ``vocab.txt`` is
available offline, so the vocabulary is the deterministic synthetic one defined
here and ``h2z`` is restated from jaconv's published behaviour (half-width
ASCII / digits / katakana -> full-width).  Both are assumptions recorded in
DESIGN.md; swapping in the real ``vocab.txt`` is ``Vocab.from_file``.
"""
from __future__ import annotations

import os
import re
from typing import Optional, Iterable, List, Sequence

import numpy as np

from .weights import VOCAB

SPECIAL_TOKENS = ["[PAD]", "[UNK]", "[CLS]", "[SEP]", "[MASK]"]
NUM_SPECIAL = len(SPECIAL_TOKENS)


def _synthetic_tokens() -> List[str]:
    toks = list(SPECIAL_TOKENS)
    seen = set()

    def add(chars: Iterable[str]) -> None:
        for ch in chars:
            if ch not in seen and len(toks) < VOCAB:
                seen.add(ch)
                toks.append(ch)

    add(chr(c) for c in range(0x3041, 0x3097))        # hiragana
    add(chr(c) for c in range(0x30A1, 0x30FB))        # katakana
    add("ー…・。、「」『』！？")
    add(chr(c) for c in range(0x21, 0x7F))            # half-width ASCII incl. digits
    add(chr(c) for c in range(0xFF66, 0xFFA0))        # half-width katakana + ﾞ ﾟ
    add(chr(c) for c in range(0x4E00, 0x4E00 + VOCAB))  # CJK ideographs fill the rest
    assert len(toks) == VOCAB
    return toks


class Vocab:
    """id -> token table of a character-level vocabulary (one token per id)."""

    def __init__(self, tokens: Sequence[str], special_ids: Optional[Iterable[int]] = None):
        self.tokens = list(tokens)
        # ids whose token is [..]-bracketed specials are dropped on decode
        self.special_ids = (frozenset(i for i, t in enumerate(self.tokens) if t in SPECIAL_TOKENS) if special_ids is None
                            else frozenset(int(i) for i in special_ids))
        self._special_mask = np.zeros(len(self.tokens), bool)
        self._special_mask[list(self.special_ids)] = True

    @classmethod
    def synthetic(cls) -> "Vocab":
        return cls(_synthetic_tokens())

    @classmethod
    def from_file(cls, path: str) -> "Vocab":
        """``vocab.txt`` as ``BertJapaneseTokenizer`` reads it (tokenization_bert_japanese.py ``load_vocab`` and the
        ``ids_to_tokens`` table built from it): one token per line, "\n" stripped; of a token that occurs on several lines only
        the LAST line keeps it, the earlier ids decode to "[UNK]" - and are not special: ``skip_special_tokens`` goes by id, and
        the special ids are the ones the five special token strings map to.  Checked against the tokenizer itself in
        tests/test_text.py."""
        with open(path, encoding="utf-8") as f:
            lines = [line.rstrip("\n") for line in f.readlines()]
        last = {t: i for i, t in enumerate(lines)}
        tokens = [t if last[t] == i else "[UNK]" for i, t in enumerate(lines)]
        special = {last.get(t, last.get("[UNK]")) for t in SPECIAL_TOKENS}
        return cls(tokens, special_ids=[i for i in special if i is not None])

    def _kept_tokens(self, ids: Iterable[int]) -> List[str]:
        n = len(self.tokens)
        if isinstance(ids, np.ndarray):
            row = ids.astype(np.int64, copy=False).ravel()
            mask = self._special_mask
            ok = (row >= 0) & (row < n)
            keep = np.ones(row.shape, bool)
            keep[ok] = ~mask[row[ok]]
            toks = self.tokens
            return [toks[i] if 0 <= i < n else "[UNK]" for i in row[keep].tolist()]
        sp = self.special_ids
        return [self.tokens[i] if 0 <= i < n else "[UNK]" for i in (int(x) for x in ids) if i not in sp]

    def _fast_tables(self):
        """(per-token post-processed strings, mask of tokens that need the context-aware path, UTF-32 code unit
        per single-character token (0: none), mask of tokens whose only context dependence is the dot rules)."""
        if getattr(self, "_fast", None) is None:
            # (a token that starts with "##" is glued to its predecessor by decode(): context-dependent as well)
            slow = np.array([t.startswith("##") or any(ch in _TRIGGERS or ch.isspace() for ch in t) for t in self.tokens], bool)
            tz = [t if s else h2z(t) for t, s in zip(self.tokens, slow)]
            dots = np.array([s and len(t) == 1 and t in _DOT_TRIGGERS for t, s in zip(self.tokens, slow)], bool)
            # code unit of every token whose (post-processed, or raw for the dot characters) form is ONE character
            cp = np.array([ord(z) if (len(z) == 1 and (not s or d)) else 0 for z, s, d in zip(tz, slow, dots)], np.uint32)
            self._fast = (tz, slow, cp, dots)
        return self._fast

    def decode(self, ids: Iterable[int]) -> str:
        """``tokenizer.decode(ids, skip_special_tokens=True)`` of the character-level ``BertJapaneseTokenizer``:
        the kept tokens joined by spaces, every " ##" removed (a word-piece continuation is glued to its
        predecessor; a leading one keeps its hashes) and the ends stripped
        (transformers/models/bert_japanese/tokenization_bert_japanese.py:256-261)."""
        return " ".join(self._kept_tokens(ids)).replace(" ##", "").strip()


# --- jaconv.h2z(ascii=True, digit=True, kana=True) restated -----------------

_HW_KANA = "ｦｧｨｩｪｫｬｭｮｯｰｱｲｳｴｵｶｷｸｹｺｻｼｽｾｿﾀﾁﾂﾃﾄﾅﾆﾇﾈﾉﾊﾋﾌﾍﾎﾏﾐﾑﾒﾓﾔﾕﾖﾗﾘﾙﾚﾛﾜﾝ"
_FW_KANA = "ヲァィゥェォャュョッーアイウエオカキクケコサシスセソタチツテトナニヌネノハヒフヘホマミムメモヤユヨラリルレロワン"
_HW_PUNCT = "｡｢｣､･ﾞﾟ"
_FW_PUNCT = "。「」、・゛゜"
_VOICED_SRC = "ｶｷｸｹｺｻｼｽｾｿﾀﾁﾂﾃﾄﾊﾋﾌﾍﾎｳ"
_VOICED_DST = "ガギグゲゴザジズゼゾダヂヅデドバビブベボヴ"
_SEMI_SRC = "ﾊﾋﾌﾍﾎ"
_SEMI_DST = "パピプペポ"

# ASCII punctuation: jaconv's conversion table (jaconv/conv_table.py, HALF_ASCII -> FULL_ASCII) is NOT the uniform
# U+FF01..U+FF5E block shift for four characters - it follows the JIS X 0208 look-alikes:
#     "  ->  ”  (U+201D)      '  ->  ’  (U+2019)      \  ->  ￥  (U+FFE5)      `  ->  ‘  (U+2018)
# Everything else in 0x21..0x7E (letters, digits, ~ -> ～ U+FF5E, - -> － U+FF0D, {|} -> ｛｜｝ ...) is the block shift.
# PROVENANCE: restated from the published table as remembered - the package is absent offline,
# so this row stays "unpinned against upstream"; tests/test_text.py lists the expected output character by character
# so that a reader with the package at hand can check it in one glance.  MOCR_H2Z_ASCII=unicode selects the plain
# block shift instead (what unicodedata-style widening would give).
H2Z_ASCII_SPECIAL = {'"': "”", "'": "’", "\\": "￥", "`": "‘"}


def _ascii_table(mode: str):
    t = {c: chr(c + 0xFEE0) for c in range(0x21, 0x7F)}   # ASCII + digits
    if mode != "unicode":
        t.update({ord(k): v for k, v in H2Z_ASCII_SPECIAL.items()})
    t[0x20] = "　"
    return t


_H2Z_SINGLE = {ord(a): b for a, b in zip(_HW_KANA + _HW_PUNCT, _FW_KANA + _FW_PUNCT)}
_H2Z_SINGLE.update(_ascii_table(os.environ.get("MOCR_H2Z_ASCII", "jaconv")))
_H2Z_PAIRS = {a + "ﾞ": b for a, b in zip(_VOICED_SRC, _VOICED_DST)}
_H2Z_PAIRS.update({a + "ﾟ": b for a, b in zip(_SEMI_SRC, _SEMI_DST)})
_PAIR_RE = re.compile("|".join(re.escape(k) for k in _H2Z_PAIRS))


def h2z(text: str) -> str:
    """Half-width -> full-width (ASCII, digits, katakana incl. (semi-)voiced marks)."""
    text = _PAIR_RE.sub(lambda m: _H2Z_PAIRS[m.group(0)], text)
    return text.translate(_H2Z_SINGLE)


_DOTS_RE = re.compile("[・.]{2,}")


def post_process(text: str) -> str:
    """Upstream ``manga_ocr.ocr.post_process`` restated."""
    text = "".join(text.split())
    text = text.replace("…", "...")
    text = _DOTS_RE.sub(lambda m: (m.end() - m.start()) * ".", text)
    return h2z(text)


_TRIGGERS = set("ﾞﾟ…・.･")     # characters whose post-processing depends on their neighbours (･ becomes ・ AFTER the dot rule)
_DOT_TRIGGERS = set("…・.")      # ... of which these only take part in the ellipsis / dot-run rules


def ids_to_text(vocab: Vocab, ids: Iterable[int]) -> str:
    """``post_process(tokenizer.decode(ids, skip_special_tokens=True))``.  The reference joins the
    tokens with spaces, drops every " ##" and then strips ALL whitespace; for tokens that do not start
    with "##" concatenating directly is the same string.
    Fast path: when no token of the row contains a context-dependent character (dots, ellipsis,
    half-width voiced marks) or whitespace, post_process acts on every character independently and
    the per-token results are precomputed."""
    if isinstance(ids, np.ndarray):
        fast = vocab._fast_tables()
        row = ids.astype(np.int64, copy=False).ravel()
        n = len(vocab.tokens)
        if row.size and row.min() >= 0 and row.max() < n:
            kept = row[~vocab._special_mask[row]]
            if not fast[1][kept].any():
                tz = fast[0]
                return "".join([tz[i] for i in kept.tolist()])
    return post_process(vocab.decode(ids))


def ids_to_texts(vocab: Vocab, ids: np.ndarray) -> List[str]:
    """``ids_to_text`` for a whole batch ``[n, T]``: one vectorised pass drops the special ids and maps
    every single-character token to its post-processed UTF-32 code unit; only rows that contain a
    context-dependent or multi-character token take the per-row path."""
    ids = np.asarray(ids)
    if ids.ndim != 2:
        raise ValueError("ids must be [n, T]")
    n_tok = len(vocab.tokens)
    if ids.size == 0 or int(ids.min()) < 0 or int(ids.max()) >= n_tok:
        return [ids_to_text(vocab, row) for row in ids]
    _, slow, cp, dots = vocab._fast_tables()
    keep = ~vocab._special_mask[ids]
    cps = cp[ids]
    per_row = (((slow[ids] & ~dots[ids]) | (cps == 0)) & keep).any(axis=1)
    has_dots = (dots[ids] & keep).any(axis=1)
    out = []
    for r in range(ids.shape[0]):
        if per_row[r]:
            out.append(ids_to_text(vocab, ids[r]))
            continue
        text = cps[r][keep[r]].tobytes().decode("utf-32-le")
        if has_dots[r]:
            # every other character is already in its final form and is not touched by these rules
            text = _DOTS_RE.sub(lambda m: (m.end() - m.start()) * ".", text.replace("…", "...")).replace(".", "．")
        out.append(text)
    return out
