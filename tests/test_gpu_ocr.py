"""The drop-in ``MangaOcr`` on a B200: strings equal to the reference path's, the call contract
(path / PIL / ValueError), concurrent callers (the app's worker threads) and the full-size
batch-64 x max_length-300 configuration checked through size-independent properties."""
import threading
import time

import numpy as np
import pytest
from PIL import Image

from manga_ocr_b200 import crops as C

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", params=[True, False], ids=["admission", "batch_dispatcher"])
def ocr16(weights0, request):
    """__call__ traffic served by admission into a running decode (the default for greedy decoding) and by the batch dispatcher."""
    from manga_ocr_b200.ocr import MangaOcr
    o = MangaOcr(weights=weights0, devices=[0], max_batch=16, max_length=12, admission=request.param)
    assert o.admission == request.param
    yield o
    o.close()


def test_strings_match_reference(ocr16, oracle12, golden_text):
    from oracle import make_golden as G
    crops = G.model_inputs()
    for c, want in zip(crops, golden_text["texts"]):
        got = ocr16(Image.fromarray(c))
        ref = oracle12(Image.fromarray(c))
        assert ref == want
        if got != ref:      # only a near-tie of the reference's top-2 logits may differ
            _, logits = oracle12.generate_batch([c], max_length=12)
            top2 = np.sort(logits[0], axis=-1)[:, -2:]
            assert (top2[:, 1] - top2[:, 0]).min() <= 6e-2
    assert ocr16.recognize_batch(crops) == [ocr16(Image.fromarray(c)) for c in crops]


def test_call_contract(ocr16, tmp_path):
    crop = C.single_224()[0]
    p = tmp_path / "crop.png"
    Image.fromarray(crop).save(p)
    a = ocr16(Image.fromarray(crop))
    assert isinstance(a, str) and ocr16(str(p)) == a and ocr16(p) == a
    assert ocr16(Image.fromarray(crop[..., 0])) == a             # L-mode image of the same pixels
    with pytest.raises(ValueError, match="img_or_path must be a path or PIL.Image"):
        ocr16(crop)
    with pytest.raises(FileNotFoundError):
        ocr16(str(tmp_path / "missing.png"))
    assert ocr16(Image.fromarray(crop)) == a                      # still usable after the errors


def test_concurrent_callers_are_batched(ocr16):
    crops = C.bubble_batch(16, seed=77)
    want = ocr16.recognize_batch(crops)
    got = [None] * 48

    def worker(k):
        got[k] = ocr16(Image.fromarray(crops[k % 16]))

    ts = [threading.Thread(target=worker, args=(k,)) for k in range(48)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    assert got == want * 3


def test_full_size_batch64_len300_properties(weights0):
    """BASELINE configs[1] at full size: determinism, batch-invariance and the greedy invariants
    ([CLS] first, PAD only after [SEP], exactly max_length-1 steps when EOS never wins)."""
    from manga_ocr_b200.engine import Engine
    crops = C.bubble_batch(64)
    eng = Engine(weights0, device=0, max_batch=64, max_length=300)
    try:
        ids, lens = eng.recognize(crops)
        ids2, _ = eng.recognize(crops)
        assert np.array_equal(ids, ids2)
        assert (ids[:, 0] == 2).all() and ids.min() >= 0 and ids.max() < 6144
        for b in range(64):
            n = lens[b]
            assert (ids[b, n:] == 0).all() and (ids[b, 1:n] != 0).all()
            if n < 300:
                assert ids[b, n - 1] == 3
        sub, _ = eng.recognize(crops[40:47])
        assert np.array_equal(sub, ids[40:47])
        assert eng.last_steps <= 299
    finally:
        eng.close()


def test_checkpoint_directory_and_vocab_file(tmp_path, weights0):
    """`MangaOcr(path)` with a local checkpoint directory (weights.npz + vocab.txt) - the route a real
    kha-white/manga-ocr-base snapshot takes (model.safetensors is read by the same loader)."""
    from manga_ocr_b200.ocr import MangaOcr
    from manga_ocr_b200.text import Vocab
    d = tmp_path / "ckpt"
    d.mkdir()
    np.savez(d / "weights.npz", **{k: v for k, v in weights0.items() if k != "decoder.cls.predictions.decoder.weight"})   # tied head omitted
    toks = Vocab.synthetic().tokens
    toks[100], toks[101] = toks[101], toks[100]           # a vocabulary that differs from the built-in one
    (d / "vocab.txt").write_text("\n".join(toks) + "\n", encoding="utf-8")
    o = MangaOcr(str(d), max_batch=4, max_length=12)
    ref = MangaOcr(weights=weights0, max_batch=4, max_length=12, warmup=False)
    try:
        crop = C.single_224()[0]
        assert o.vocab.tokens[100] == toks[100]
        ids_a = o.recognize_ids([crop])
        ids_b = ref.recognize_ids([crop])
        assert np.array_equal(ids_a, ids_b)               # same weights through the file route
        assert isinstance(o(Image.fromarray(crop)), str)
    finally:
        o.close()
        ref.close()


def test_two_engines_share_a_gpu(weights0):
    """Two handles on one device (the in-process multi-GPU dispatcher uses one handle per GPU; on a
    single-GPU box both dispatchers land on cuda:0) give the same strings as one."""
    from manga_ocr_b200.ocr import MangaOcr
    crops = C.bubble_batch(12, seed=5)
    one = MangaOcr(weights=weights0, devices=[0], max_batch=8, max_length=12, warmup=False)
    two = MangaOcr(weights=weights0, devices=[0, 0], max_batch=8, max_length=12, warmup=False)
    try:
        want = one.recognize_batch(crops)
        assert two.recognize_batch(crops) == want          # splitter: two contiguous blocks, two worker threads
        got = [None] * 12
        ts = [threading.Thread(target=lambda k=k: got.__setitem__(k, two(Image.fromarray(crops[k])))) for k in range(12)]
        for t in ts:
            t.start()
        for t in ts:
            t.join()
        assert got == want
    finally:
        one.close()
        two.close()


def test_regions_follow_the_generation_config_like_the_other_entry_points(weights0):
    """With beam settings (what the shipped checkpoint's generation config selects), recognize_regions must give
    the strings recognize_batch gives for the same selections staged on the host (one generate() behaviour for
    every API of the drop-in)."""
    from manga_ocr_b200 import weights as W
    from manga_ocr_b200.engine import Region
    from manga_ocr_b200.ocr import MangaOcr
    from oracle import crop_staging_np as S
    w = W.random_init(0, gain=3.0, eos_bias=4.2)
    page, sels = C.page_with_selections(7, seed=21, height=700, width=600)
    regions = [Region.from_qt(rect, poly, orient) for rect, poly, orient in sels]
    with MangaOcr(weights=w, devices=[0], max_batch=16, max_length=20, warmup=False, num_beams=4, no_repeat_ngram_size=3,
                  length_penalty=2.0, early_stopping=True) as ocr:
        staged = [S.stage_region(page, r.box, r.polygon, r.rotate) for r in regions]
        want = ocr.recognize_batch(staged)                 # 7 crops x 4 beams > 16 rows: chunked inside the library
        assert ocr.recognize_regions(page, regions) == want
        greedy = ocr.engines[0].recognize_regions(page, regions, max_length=20)[0]
    from manga_ocr_b200.text import Vocab, ids_to_texts
    assert ids_to_texts(Vocab.synthetic(), greedy) != want    # the beam settings really changed the result


def test_beam_calls_from_many_threads_do_not_interleave(weights0):
    """Beam mode used to be four separately locked library calls (stage, preprocess, encode, decode_beam): two threads
    sharing an engine could decode each other's crops.  It is one call now."""
    from manga_ocr_b200 import weights as W
    from manga_ocr_b200.ocr import MangaOcr
    w = W.random_init(0, gain=3.0, eos_bias=4.2)
    crops = C.bubble_batch(12, seed=31)
    with MangaOcr(weights=w, devices=[0], max_batch=16, max_length=16, warmup=False, num_beams=4) as ocr:
        want = ocr.recognize_batch(crops)
        got = [None] * 12
        errs = []

        def worker(k):
            try:
                if k % 2:
                    got[k] = ocr.recognize_batch([crops[k]])[0]        # caller threads straight into the engine
                else:
                    got[k] = ocr(Image.fromarray(crops[k]))             # ... and through the dispatcher
            except Exception as e:     # noqa: BLE001
                errs.append(e)

        ts = [threading.Thread(target=worker, args=(k,)) for k in range(12)]
        for t in ts:
            t.start()
        for t in ts:
            t.join()
        assert not errs, errs
        assert got == want


def test_one_bad_request_does_not_fail_its_batch(ocr16):
    """The reference isolates failures per call (workers catch per item, reference/src/core/workers.py:241-244): a crop
    the engine refuses (here: wider than its 32768 px limit is caught before it joins a batch; a crop whose resampling
    tables exceed the shared-memory capacity fails inside the library) must not fail the callers batched with it."""
    crops = C.bubble_batch(6, seed=78)
    want = ocr16.recognize_batch(crops)
    huge = Image.fromarray(np.full((30000, 9, 3), 255, np.uint8))        # 30000 rows -> 224: the vertical strip does not fit
    results = {}

    def good(k):
        results[k] = ocr16(Image.fromarray(crops[k]))

    def bad():
        try:
            ocr16(huge)
            results["bad"] = "no error"
        except Exception as e:     # noqa: BLE001
            results["bad"] = e

    ts = [threading.Thread(target=good, args=(k,)) for k in range(6)] + [threading.Thread(target=bad)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    assert [results[k] for k in range(6)] == want
    assert isinstance(results["bad"], Exception) or isinstance(results["bad"], str)
    with pytest.raises(ValueError):
        ocr16(Image.fromarray(np.zeros((4, 40000, 3), np.uint8)))       # refused before it is queued
    assert ocr16(Image.fromarray(crops[0])) == want[0]                   # the instance stays usable


def test_instance_is_collected_without_close(weights0):
    """Dispatcher threads hold the instance weakly: dropping the last reference frees the engines."""
    import gc
    import weakref
    from manga_ocr_b200.ocr import MangaOcr
    ocr = MangaOcr(weights=weights0, devices=[0], max_batch=4, max_length=8, warmup=False)
    assert isinstance(ocr(Image.fromarray(C.single_224()[0])), str)
    threads = list(ocr._threads)
    ref = weakref.ref(ocr)
    del ocr
    for _ in range(200):          # (a dispatcher may still be closing the session that served the call)
        gc.collect()
        if ref() is None:
            break
        time.sleep(0.01)
    assert ref() is None
    for t in threads:
        t.join(timeout=5)
        assert not t.is_alive()
