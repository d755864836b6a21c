"""Beam-search mode on the B200 against transformers' generate(num_beams=4,
no_repeat_ngram_size=3, length_penalty=2.0) on the fp32 CPU oracle.  The selection logic is pinned exactly by
tests/test_beam_host.py; here the bf16 device path supplies the log-probabilities, so a hypothesis may differ from the
oracle's only where the two are near-ties: |score difference| <= 5e-2 (logits tolerance of the greedy path: 6e-2)."""
import numpy as np
import pytest

from manga_ocr_b200 import crops as C
from manga_ocr_b200 import weights as W

pytestmark = pytest.mark.gpu
T = 20


@pytest.fixture(scope="module")
def beam_setup():
    import torch
    from manga_ocr_b200.engine import Engine
    from manga_ocr_b200.text import Vocab
    from oracle.reference_ocr import ReferenceMangaOcr
    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    w = W.random_init(0, gain=3.0, eos_bias=4.2)
    eng = Engine(w, device=0, max_batch=16, max_length=T)
    ocr = ReferenceMangaOcr(w, Vocab.synthetic().tokens, max_length=T)
    crops = C.bubble_batch(6, seed=41)
    yield eng, ocr, crops
    eng.close()


def _oracle_beam(ocr, crops, **kw):
    import torch
    x = torch.stack([ocr.pixel_values(c) for c in crops])
    with torch.no_grad():
        out = ocr.model.generate(x, max_length=T, do_sample=False, output_scores=True, return_dict_in_generate=True, **kw)
    return out.sequences.numpy(), out.sequences_scores.numpy()


def _oracle_score(ocr, crop, seq, lp):
    """Sum of the fp32 oracle's log-probabilities of the generated tokens / (their number ** length_penalty): what
    transformers reports as sequences_scores for a finished hypothesis."""
    import torch
    x = ocr.pixel_values(crop)[None]
    with torch.no_grad():
        enc = ocr.model.encoder(pixel_values=x).last_hidden_state
        t = torch.tensor(seq[None, :-1].astype(np.int64))
        logp = torch.nn.functional.log_softmax(ocr.model.decoder(input_ids=t, encoder_hidden_states=enc).logits[0].float(), dim=-1)
    gen = seq[1:]
    total = float(sum(logp[i, int(tok)] for i, tok in enumerate(gen)))
    return total / (len(gen) ** lp)


def _has_repeated_ngram(seq, n=3):
    grams = [tuple(seq[i:i + n]) for i in range(len(seq) - n + 1)]
    return len(grams) != len(set(grams))


@pytest.mark.parametrize("early", [True, False])
def test_beam_search_matches_oracle(beam_setup, early):
    """Most crops: the oracle's hypothesis exactly.  bf16 log-probabilities can send the search down another branch at
    a near-tie; such a hypothesis must still be legal (ends at EOS or max_length, no repeated 3-gram) and carry the
    score the fp32 oracle assigns to it (the device's scoring is faithful) - and is never much worse than the
    oracle's own result."""
    eng, ocr, crops = beam_setup
    ref_ids, ref_scores = _oracle_beam(ocr, crops, num_beams=4, no_repeat_ngram_size=3, length_penalty=2.0, early_stopping=early)
    ids, lens, scores = eng.recognize_beam(crops, max_length=T, num_beams=4, no_repeat_ngram_size=3, length_penalty=2.0,
                                           early_stopping=early)          # 6 crops x 4 beams > 16 rows: two chunks
    assert ids.shape == (6, T) and (ids[:, 0] == 2).all()
    same = 0
    for i in range(6):
        L = int(lens[i])
        seq = ids[i, :L]
        assert (ids[i, L:] == 3).all()
        assert seq[-1] == 3 or L == T, (i, seq)
        assert not _has_repeated_ngram(seq.tolist()), (i, seq)
        assert abs(_oracle_score(ocr, crops[i], seq, 2.0) - scores[i]) <= 2e-2, (i, seq, scores[i])
        ref = ref_ids[i]
        ref_len = len(ref)
        while ref_len > 1 and ref[ref_len - 1] == 3 and ref[ref_len - 2] == 3:
            ref_len -= 1                                                   # the reference fills with EOS past the end
        if L == ref_len and np.array_equal(seq, ref[:L]):
            same += 1
        else:
            assert scores[i] >= ref_scores[i] - 0.25, (i, seq, ref, scores[i], ref_scores[i])
    assert same >= 4, same


def test_single_beam_without_penalties_is_the_greedy_path(beam_setup):
    eng, ocr, crops = beam_setup
    g_ids, g_lens = eng.recognize(crops[:4], max_length=T)
    b_ids, b_lens, _ = eng.recognize_beam(crops[:4], max_length=T, num_beams=1, no_repeat_ngram_size=0, length_penalty=1.0,
                                          early_stopping=True)
    for i in range(4):
        L = int(g_lens[i])
        assert int(b_lens[i]) == L and np.array_equal(b_ids[i, :L], g_ids[i, :L]), i


def test_beam_capacity_and_argument_errors(beam_setup):
    from manga_ocr_b200.engine import MocrError
    eng, _, crops = beam_setup
    eng.stage(crops[:5]); eng.preprocess(); eng.encode()
    with pytest.raises(MocrError) as e:
        eng.decode_beam(4, T)                       # 5 x 4 = 20 rows > 16
    assert e.value.code == -4
    with pytest.raises(MocrError) as e:
        eng.decode_beam(0, T)
    assert e.value.code == -1
    ids, lens, scores = eng.decode_beam(3, T)       # 15 rows: fine, handle still usable
    assert ids.shape == (5, T) and np.isfinite(scores).all()


def test_manga_ocr_front_end_runs_the_checkpoint_generation_config(beam_setup, tmp_path):
    """MangaOcr(...) picks num_beams / no_repeat_ngram_size / length_penalty up from the checkpoint directory, as
    transformers' generate() does, and the single-image call returns the beam result."""
    import json
    from manga_ocr_b200.ocr import MangaOcr
    from manga_ocr_b200.text import Vocab, ids_to_texts
    eng, _, crops = beam_setup
    w = W.random_init(0, gain=3.0, eos_bias=4.2)
    np.savez(tmp_path / "weights.npz", **w)
    (tmp_path / "generation_config.json").write_text(json.dumps(
        {"num_beams": 4, "no_repeat_ngram_size": 3, "length_penalty": 2.0, "early_stopping": True, "max_length": 300}))
    ocr = MangaOcr(str(tmp_path), max_batch=16, max_length=T, warmup=False)
    try:
        assert ocr.generation["num_beams"] == 4 and ocr.generation["no_repeat_ngram_size"] == 3
        want = ids_to_texts(Vocab.synthetic(), eng.recognize_beam(crops[:3], max_length=T)[0])
        assert ocr.recognize_batch(crops[:3]) == want
        from PIL import Image
        assert ocr(Image.fromarray(crops[1])) == want[1]
    finally:
        ocr.close()
    greedy = MangaOcr(str(tmp_path), max_batch=16, max_length=T, warmup=False, num_beams=1, no_repeat_ngram_size=0)
    try:
        assert greedy.generation["num_beams"] == 1 and greedy._beam_args() is None      # the plain arg-max path
        assert greedy.recognize_batch(crops[:2]) == ids_to_texts(Vocab.synthetic(), eng.recognize(crops[:2], max_length=T)[0])
    finally:
        greedy.close()


def test_beam_and_region_calls_interleave_on_one_handle(beam_setup):
    """Scratch buffers of the two widened paths are independent: beam -> regions (growing mask arena) -> beam."""
    from manga_ocr_b200.engine import Region
    eng, _, crops = beam_setup
    first = eng.recognize_beam(crops[:3], max_length=T)
    page, sels = C.page_with_selections(12, seed=9, height=900, width=700)
    regions = [Region.from_qt(r, p, o) for r, p, o in sels]
    ids_a, _ = eng.recognize_regions(page, regions, max_length=T)
    big = np.full((1500, 1200, 3), 200, np.uint8)
    poly = np.array([[5, 5], [1190, 10], [1195, 1490], [10, 1495]], np.int32)
    eng.recognize_regions(big, [Region((0, 0, 1200, 1500), poly)], max_length=T)       # a 1.8 MB mask: the arena grows
    again = eng.recognize_beam(crops[:3], max_length=T)
    assert all(np.array_equal(a, b) for a, b in zip(first, again))
    ids_b, _ = eng.recognize_regions(page, regions, max_length=T)
    assert np.array_equal(ids_a, ids_b)


@pytest.mark.parametrize("early", [True, False, "never"])
@pytest.mark.parametrize("ngram,lp", [(3, 2.0), (0, 1.0), (2, 0.5)])
def test_device_resident_beam_search_equals_host_bookkeeping(beam_setup, early, ngram, lp):
    """The selection kernels (beam_device.cuh) restate beam_search.h, which tests/test_beam_host.py pins exactly against
    transformers: on the same device logits both must return the same hypotheses - including the n-gram ban computed on the
    device and the cache rows that follow the beams through the row table instead of the per-step gather."""
    eng, _, crops = beam_setup
    kw = dict(max_length=T, num_beams=4, no_repeat_ngram_size=ngram, length_penalty=lp, early_stopping=early)
    eng.set_option("beam_device", 0)
    try:
        ids_h, lens_h, scores_h = eng.recognize_beam(crops, **kw)
    finally:
        eng.set_option("beam_device", 1)
    ids_d, lens_d, scores_d = eng.recognize_beam(crops, **kw)
    assert np.array_equal(lens_d, lens_h), (lens_d, lens_h)
    assert np.array_equal(ids_d, ids_h)
    assert np.allclose(scores_d, scores_h, rtol=1e-6, atol=1e-7)
    eng.set_option("use_graph", 0)                 # the same steps launched one by one
    try:
        ids_n, lens_n, _ = eng.recognize_beam(crops, **kw)
    finally:
        eng.set_option("use_graph", 1)
    assert np.array_equal(ids_n, ids_d) and np.array_equal(lens_n, lens_d)


def test_device_beam_other_widths(beam_setup):
    eng, _, crops = beam_setup
    for beams in (1, 2, 3, 8):
        kw = dict(max_length=T, num_beams=beams, no_repeat_ngram_size=3, length_penalty=2.0, early_stopping=True)
        eng.set_option("beam_device", 0)
        try:
            want = eng.recognize_beam(crops[:2], **kw)
        finally:
            eng.set_option("beam_device", 1)
        got = eng.recognize_beam(crops[:2], **kw)
        assert np.array_equal(got[0], want[0]) and np.array_equal(got[1], want[1]), beams
