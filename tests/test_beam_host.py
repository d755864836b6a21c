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


def _drive_scorer(model, x, ngram, lp, early):
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
