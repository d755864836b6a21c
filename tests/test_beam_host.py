"""Beam-search bookkeeping (manga_ocr_b200/csrc/beam_search.h, through the C ABI) against transformers' own
``generate(num_beams=4, no_repeat_ngram_size=3, length_penalty=2.0)`` with the SAME fp32 CPU model supplying the
per-step log-probabilities: the restated selection logic must reproduce the reference's sequences exactly.
(reference: transformers generation/utils.py:3076-3370, logits_process.py:1012-1136.)"""
import numpy as np
import pytest

from manga_ocr_b200 import crops as C
from manga_ocr_b200 import weights as W
from manga_ocr_b200.engine import BeamScorer

T = 14
BEAMS = 4


@pytest.fixture(scope="module")
def beam_model():
    import torch
    from oracle.reference_ocr import ReferenceMangaOcr
    from manga_ocr_b200.text import Vocab
    torch.manual_seed(0)
    ocr = ReferenceMangaOcr(W.random_init(0, gain=3.0, eos_bias=4.2), Vocab.synthetic().tokens, max_length=T)
    imgs = C.bubble_batch(3, seed=31)
    x = torch.stack([ocr.pixel_values(i) for i in imgs])
    return ocr.model, x


def _drive_scorer(model, x, ngram, lp, early, BEAMS=BEAMS, T=T):
    """Our scorer in the loop, the oracle model as the source of log-probabilities (no cache: full decoder forward)."""
    import torch
    n = x.shape[0]
    sc = BeamScorer(n, BEAMS, T, ngram, lp, early)
    with torch.no_grad():
        enc = model.encoder(pixel_values=x).last_hidden_state
        enc_rows = enc.repeat_interleave(BEAMS, dim=0)
        seqs = torch.full((n * BEAMS, 1), 2, dtype=torch.long)
        unfinished = True
        while unfinished and seqs.shape[1] < T:
            logits = model.decoder(input_ids=seqs, encoder_hidden_states=enc_rows).logits[:, -1, :].float()
            logp = torch.nn.functional.log_softmax(logits, dim=-1)
            for r in range(n * BEAMS):
                banned = sc.banned(r)
                if banned:
                    logp[r, banned] = -float("inf")
            top = torch.topk(logp, k=2 * BEAMS, dim=-1)
            unfinished, nxt, par = sc.step(top.values.numpy(), top.indices.numpy().astype(np.int32))
            if not unfinished:
                break
            seqs = torch.cat([seqs[torch.from_numpy(par).long()], torch.from_numpy(nxt).long()[:, None]], dim=1)
    out = sc.result()
    sc.close()
    return out


@pytest.mark.parametrize("ngram,lp,early", [(3, 2.0, True), (3, 2.0, False), (0, 1.0, True), (2, 0.0, "never"), (3, 1.0, "never")])
def test_scorer_reproduces_transformers_beam_search(beam_model, ngram, lp, early):
    import torch
    model, x = beam_model
    with torch.no_grad():
        ref = model.generate(x, max_length=T, num_beams=BEAMS, do_sample=False, no_repeat_ngram_size=ngram, length_penalty=lp,
                             early_stopping=early, output_scores=True, return_dict_in_generate=True)
    ids, lens, scores = _drive_scorer(model, x, ngram, lp, early)
    ref_ids = ref.sequences.numpy()
    assert np.array_equal(ids[:, :ref_ids.shape[1]], ref_ids), (ids, ref_ids)
    assert (ids[:, ref_ids.shape[1]:] == 3).all()      # the reference's fill value is EOS when the PAD id is 0 (utils.py:3163)
    assert np.allclose(scores, ref.sequences_scores.numpy(), rtol=1e-5, atol=1e-5)
    assert lens.max() == ref_ids.shape[1]


@pytest.fixture(scope="module")
def repeating_model():
    """Weights whose plain greedy decoding repeats one token (no EOS bias): the n-gram ban changes every row."""
    import torch
    from oracle.reference_ocr import ReferenceMangaOcr
    from manga_ocr_b200.text import Vocab
    ocr = ReferenceMangaOcr(W.random_init(0, gain=3.0), Vocab.synthetic().tokens, max_length=T)
    x = torch.stack([ocr.pixel_values(i) for i in C.bubble_batch(3, seed=31)])
    return ocr.model, x


@pytest.mark.parametrize("ngram", [1, 2, 3])
def test_one_beam_search_is_greedy_decoding_with_the_ngram_ban(repeating_model, ngram):
    """generate(num_beams=1, no_repeat_ngram_size=n) is greedy decoding through NoRepeatNGramLogitsProcessor.  The engine's
    arg-max path has no ban list; MangaOcr routes such a checkpoint to the search with ONE beam and early_stopping=True
    (ocr.py::_beam_args), which must give the reference's greedy ids exactly: the single running beam follows the banned
    arg-max, and the search ends with the first finished hypothesis, where greedy decoding stops."""
    import torch
    model, x = repeating_model
    with torch.no_grad():
        ref = model.generate(x, max_length=T, num_beams=1, do_sample=False, no_repeat_ngram_size=ngram).numpy()
        plain = model.generate(x, max_length=T, num_beams=1, do_sample=False).numpy()
    assert ref.shape != plain.shape or not np.array_equal(ref, plain)          # the ban bites on this fixture
    ids, lens, _ = _drive_scorer(model, x, ngram, 1.0, True, BEAMS=1)
    for i in range(ref.shape[0]):
        row = ref[i]
        L = len(row)
        while L > 1 and row[L - 1] == 0:          # greedy pads a finished row with PAD (id 0)
            L -= 1
        assert int(lens[i]) == L and np.array_equal(ids[i, :L], row[:L]), (i, ids[i], row)


@pytest.mark.parametrize("ngram,lp,early", [(3, 2.0, True), (3, 2.0, False), (2, 1.0, True), (0, 1.0, False), (3, 0.5, "never"), (1, 2.0, True)])
def test_scorer_reproduces_transformers_beam_search_on_long_hypotheses(repeating_model, ngram, lp, early):
    """The same pin on weights without an EOS bias: every hypothesis runs to max_length, plain decoding would repeat one token,
    so the n-gram ban list, the beam reordering and the max-length finish are exercised at every step (the EOS-biased fixture
    above ends after 2-4 tokens)."""
    import torch
    model, x = repeating_model
    with torch.no_grad():
        ref = model.generate(x, max_length=T, num_beams=BEAMS, do_sample=False, no_repeat_ngram_size=ngram, length_penalty=lp,
                             early_stopping=early, output_scores=True, return_dict_in_generate=True)
    ids, lens, scores = _drive_scorer(model, x, ngram, lp, early)
    ref_ids = ref.sequences.numpy()
    assert ref_ids.shape[1] >= T - 1
    assert np.array_equal(ids[:, :ref_ids.shape[1]], ref_ids), (ids, ref_ids)
    assert np.allclose(scores, ref.sequences_scores.numpy(), rtol=1e-5, atol=1e-5)
    if ngram:
        for row in ref_ids:
            grams = [tuple(row[j:j + ngram]) for j in range(len(row) - ngram + 1)]
            assert len(grams) == len(set(grams))


@pytest.mark.parametrize("early", [True, False])
def test_scorer_reproduces_transformers_beam_search_at_40_tokens(repeating_model, early):
    """... and over 39 steps (the checkpoint's believed settings: 4 beams, no repeated 3-gram, length penalty 2)."""
    import torch
    model, x = repeating_model
    TL = 40
    with torch.no_grad():
        ref = model.generate(x[:2], max_length=TL, num_beams=BEAMS, do_sample=False, no_repeat_ngram_size=3, length_penalty=2.0,
                             early_stopping=early, output_scores=True, return_dict_in_generate=True)
    ids, lens, scores = _drive_scorer(model, x[:2], 3, 2.0, early, T=TL)
    ref_ids = ref.sequences.numpy()
    assert ref_ids.shape[1] == TL
    assert np.array_equal(ids[:, :TL], ref_ids), (ids, ref_ids)
    assert np.allclose(scores, ref.sequences_scores.numpy(), rtol=1e-5, atol=1e-5)


def test_ngram_ban_lists():
    sc = BeamScorer(1, 2, 12, 3, 1.0, True)
    assert sc.banned(0) == []                       # one token so far
    sc.close()
    # drive a fixed history through the scorer: beam 0 always takes its first candidate
    sc = BeamScorer(1, 1, 12, 3, 1.0, False)
    for tok in [7, 8, 9, 7, 8]:
        lp = np.array([[-0.1, -5.0]], np.float32)
        tk = np.array([[tok, 100]], np.int32)
        sc.step(lp, tk)
    assert sc.banned(0) == [9]                      # (7, 8) was followed by 9 before
    sc.step(np.array([[-0.1, -5.0]], np.float32), np.array([[11, 100]], np.int32))
    assert sc.banned(0) == []
    sc.close()


def test_scorer_rejects_bad_arguments():
    from manga_ocr_b200.engine import MocrError
    with pytest.raises(MocrError):
        BeamScorer(0, 4, 12)
    with pytest.raises(MocrError):
        BeamScorer(1, 4, 1)
    sc = BeamScorer(2, 4, 12)
    with pytest.raises(ValueError):
        sc.step(np.zeros((3, 8), np.float32), np.zeros((3, 8), np.int32))
    sc.close()
