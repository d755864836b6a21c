"""Parity of the CUDA path with the oracle and the golden vectors, through the C ABI, on a B200.

Tolerances (the survey measured transformers' own bf16-vs-fp32 gap):
  preprocess       bit-exact (integer work; fp32 pixel_values bit-identical)
  encoder hidden   rel-L2 <= 2e-2
  step logits      max-abs <= 6e-2, compared teacher-forced
  tokens           identical except where the reference's top-2 margin is <= 6e-2
"""
import numpy as np
import pytest

from manga_ocr_b200 import crops as C
from oracle import make_golden as G
from oracle import preprocess_np as P

pytestmark = pytest.mark.gpu

LOGIT_TOL = 6e-2
ENC_TOL = 2e-2


def _rgb(a):
    return np.stack([a] * 3, -1) if a.ndim == 2 else a[..., :3]


def _pre(engine, crops, order=0):
    from manga_ocr_b200.engine import TAP_ENCODER, TAP_LOGITS, TAP_PIXELS
    engine.set_taps(TAP_PIXELS | TAP_ENCODER | TAP_LOGITS)
    engine.stage(crops, order)
    engine.preprocess()
    return engine.pixels_u8(), engine.pixel_values()


def test_preprocess_golden_bit_exact(engine8, golden_pre):
    inputs = G.pre_inputs()
    for lo in range(0, len(inputs), 8):
        chunk = inputs[lo:lo + 8]
        u8, pv = _pre(engine8, chunk)
        for j in range(len(chunk)):
            i = lo + j
            assert np.array_equal(u8[j], golden_pre[f"u8_{i}"]), G.PRE_SHAPES[i]
            assert np.array_equal(pv[j].view(np.uint32), golden_pre["lut"][golden_pre[f"u8_{i}"]].view(np.uint32))
            if i < 3:
                assert np.array_equal(pv[j].view(np.uint32), golden_pre[f"pv_{i}"].view(np.uint32))


@pytest.mark.parametrize("maker", [lambda: C.single_224(), lambda: C.bubble_batch(8), lambda: C.page_batch(24, seed=1003),
                                   lambda: C.tall_batch(8)], ids=["cfg1", "cfg2", "cfg3", "cfg4"])
def test_preprocess_config_crops_bit_exact(engine8, maker):
    crops = maker()
    for lo in range(0, len(crops), 8):
        chunk = crops[lo:lo + 8]
        u8, pv = _pre(engine8, chunk)
        for j, c in enumerate(chunk):
            ru8, rpv = P.preprocess(c)
            assert np.array_equal(u8[j], ru8), c.shape
            assert np.array_equal(pv[j].view(np.uint32), rpv[0].view(np.uint32))


def test_preprocess_edge_shapes_and_layouts(engine8):
    rng = np.random.default_rng(5)
    shapes = [(1, 1), (1, 300), (300, 1), (2, 2), (223, 224), (224, 225), (2000, 30), (30, 2000)]
    crops = [rng.integers(0, 256, s + (3,), dtype=np.uint8) for s in shapes]
    u8, _ = _pre(engine8, crops)
    for j, c in enumerate(crops):
        assert np.array_equal(u8[j], P.preprocess(c)[0]), c.shape
    # BGR order, padded row stride, 1- and 4-channel inputs
    base = rng.integers(0, 256, (120, 90, 3), dtype=np.uint8)
    want = P.preprocess(base)[0]
    u8, _ = _pre(engine8, [np.ascontiguousarray(base[..., ::-1])], order=1)
    assert np.array_equal(u8[0], want)
    padded = np.zeros((120, 100, 3), np.uint8)
    padded[:, :90] = base
    u8, _ = _pre(engine8, [padded[:, :90]])
    assert np.array_equal(u8[0], want)
    rgba = np.concatenate([base, rng.integers(0, 256, (120, 90, 1), dtype=np.uint8)], -1)
    gray = P.rgb_to_l(base)
    u8, _ = _pre(engine8, [rgba, gray])
    assert np.array_equal(u8[0], want) and np.array_equal(u8[1], P.resize_l_224(gray))


def test_encoder_matches_golden_and_oracle(engine8, golden_model, oracle12):
    crops = G.model_inputs()
    _pre(engine8, crops)
    engine8.encode()
    enc = engine8.encoder_hidden()
    rows = enc[:, G.ENC_ROWS]
    rel = np.linalg.norm(rows - golden_model["enc_rows"]) / np.linalg.norm(golden_model["enc_rows"])
    assert rel <= ENC_TOL, rel
    ref = oracle12.encoder_hidden(crops)
    rel = np.linalg.norm(enc - ref) / np.linalg.norm(ref)
    assert rel <= ENC_TOL, rel
    assert np.abs(enc - ref).max() <= 0.15


@pytest.mark.parametrize("opts", [{"fuse_ln": 0}, {"big_rows": 1}, {"kv_prefetch": 1}, {"dec_tc": 0}],
                         ids=["split_k_ln", "large_batch_program", "kv_prefetch", "vocab_mma_sync"])
def test_decoder_program_variants_match_oracle(engine8, golden_model, oracle12, opts):
    """Every decoder program variant (not only the default) is held to the same teacher-forced tolerance."""
    crops = G.model_inputs()
    T = G.MODEL_T
    ids_ref = golden_model["ids"]
    defaults = {"fuse_ln": 1, "big_rows": 112, "kv_prefetch": 0, "dec_tc": 1}
    _pre(engine8, crops)
    engine8.encode()
    try:
        for k, v in opts.items():
            engine8.set_option(k, v)
        engine8.decode(T, forced_ids=ids_ref)
        logits = engine8.step_logits()
        ids, lens = engine8.fetch_ids()
    finally:
        for k in opts:
            engine8.set_option(k, defaults[k])
    _, logits_ref = oracle12.generate_batch(crops, max_length=T)
    assert np.abs(logits - logits_ref).max() <= LOGIT_TOL, (opts, float(np.abs(logits - logits_ref).max()))
    n_mis, n_hard = _margin_ok(ids, ids_ref, logits_ref)
    assert n_hard == 0, (opts, n_mis, n_hard)


def _margin_ok(ids, ids_ref, logits_ref, tol=LOGIT_TOL):
    top2 = np.sort(logits_ref, axis=-1)[..., -2:]
    margin = top2[..., 1] - top2[..., 0]
    mism = ids[:, 1:1 + logits_ref.shape[1]] != ids_ref[:, 1:]
    return int(mism.sum()), int((mism & (margin > tol)).sum())


def test_teacher_forced_logits_match_golden_and_oracle(engine8, golden_model, oracle12):
    crops = G.model_inputs()
    T = G.MODEL_T
    ids_ref = golden_model["ids"]
    _pre(engine8, crops)
    engine8.encode()
    engine8.decode(T, forced_ids=ids_ref)
    logits = engine8.step_logits()
    assert np.abs(logits[..., ::G.LOGIT_STRIDE] - golden_model["logits_strided"]).max() <= LOGIT_TOL
    top2 = np.take_along_axis(logits, golden_model["top2_idx"], axis=-1)
    assert np.abs(top2 - golden_model["top2_val"]).max() <= LOGIT_TOL
    _, logits_ref = oracle12.generate_batch(crops, max_length=T)
    assert np.abs(logits - logits_ref).max() <= LOGIT_TOL
    ids, lens = engine8.fetch_ids()
    n_mis, n_hard = _margin_ok(ids, ids_ref, logits_ref)
    assert n_hard == 0, (n_mis, n_hard)
    assert (ids[:, 0] == 2).all() and (lens == T).all()


def test_free_running_ids_and_eos_with_varied_weights():
    """Weights with gain 3 / EOS bias make the output image-dependent and the lengths ragged:
    EOS stop, PAD fill, finished-row masking and the all-finished early exit are exercised."""
    from manga_ocr_b200 import weights as W
    from manga_ocr_b200.engine import Engine, TAP_LOGITS
    from manga_ocr_b200.text import Vocab
    from oracle.reference_ocr import ReferenceMangaOcr
    T = 40
    w = W.random_init(0, gain=3.0, eos_bias=4.2)
    crops = C.bubble_batch(8)
    ref = ReferenceMangaOcr(w, Vocab.synthetic().tokens, max_length=T)
    ids_ref, logits_ref = ref.generate_batch(crops, max_length=T)
    ids_ref_full = np.zeros((8, T), np.int32)
    ids_ref_full[:, : ids_ref.shape[1]] = ids_ref
    eng = Engine(w, device=0, max_batch=8, max_length=T)
    try:
        eng.set_taps(TAP_LOGITS)
        eng.set_option("check_every", 4)
        eng.stage(crops)
        eng.preprocess()
        eng.encode()
        eng.decode(T, forced_ids=ids_ref_full)
        logits = eng.step_logits()[:, : logits_ref.shape[1]]
        live = ids_ref[:, 1:] != 0      # steps the reference actually decoded for that row (PAD afterwards)
        # rows the reference already finished are fed PAD and their logits are not comparable
        assert np.abs((logits - logits_ref)[live]).max() <= 3 * LOGIT_TOL      # gain 3 scales logits ~3x
        eng.decode(T)
        ids, lens = eng.fetch_ids()
        lens_ref = (ids_ref_full != 0).sum(axis=1)
        assert len(set(lens_ref.tolist())) > 2          # the fixture really is ragged
        top2 = np.sort(logits_ref, axis=-1)[..., -2:]
        margin = top2[..., 1] - top2[..., 0]
        for b in range(8):
            n = min(lens[b], lens_ref[b])
            diff = np.nonzero(ids[b, :n] != ids_ref_full[b, :n])[0]
            if len(diff) == 0:
                assert lens[b] == lens_ref[b]
                assert (ids[b, lens[b]:] == 0).all()                      # PAD fill after EOS (utils.py:2797)
                if lens[b] < T:
                    assert ids[b, lens[b] - 1] == 3                        # stopped by [SEP]
            else:                                                         # first divergence must be a near-tie
                assert margin[b, diff[0] - 1] <= 3 * LOGIT_TOL, (b, diff[0], margin[b, diff[0] - 1])
        assert eng.last_steps <= T - 1
    finally:
        eng.close()


def test_batch_invariance_and_determinism(engine8):
    """Crops are independent units: a crop's ids do not depend on what else is in the batch, on
    the chunking of mocr_recognize, or on CUDA-graph replay."""
    crops = C.page_batch(11, seed=9)
    T = 24
    ids_all, _ = engine8.recognize(crops, max_length=T)          # chunks of 8 + 3
    ids_again, _ = engine8.recognize(crops, max_length=T)
    assert np.array_equal(ids_all, ids_again)
    for i in (0, 5, 10):
        one, _ = engine8.recognize([crops[i]], max_length=T)
        assert np.array_equal(one[0], ids_all[i]), i
    engine8.set_option("use_graph", 0)
    ids_nograph, _ = engine8.recognize(crops, max_length=T)
    engine8.set_option("use_graph", 1)
    assert np.array_equal(ids_all, ids_nograph)


def test_errors_surface_and_handle_stays_usable(engine8):
    from manga_ocr_b200.engine import MocrError
    with pytest.raises(ValueError):
        engine8.recognize([np.zeros((4, 4, 2), np.uint8)])
    with pytest.raises(MocrError):
        engine8.stage([np.zeros((8, 8, 3), np.uint8)] * 9)       # beyond max_batch
    with pytest.raises(MocrError):
        engine8.recognize([np.zeros((8, 8, 3), np.uint8)], max_length=25)   # beyond max_length
    ids, lens = engine8.recognize([np.full((50, 40, 3), 255, np.uint8)], max_length=10)
    assert ids.shape == (1, 10) and ids[0, 0] == 2
    ids, lens = engine8.recognize([], max_length=10)
    assert ids.shape == (0, 10)


def test_large_ragged_batch_matches_small_batches(weights0):
    """More than one 64-row block and a ragged tail (130 crops) through every decoder stage.  Within one decoder
    program rows are independent: each crop's ids equal those of the same crop decoded in a batch of 8.  The
    large-batch program (tcgen05 GEMMs, selected by the row count) rounds differently, so it is held to
    agreement up to near-ties here and to the logits tolerance in test_decoder_program_variants_match_oracle."""
    from manga_ocr_b200.engine import Engine
    crops = C.page_batch(130, seed=11)
    T = 20
    big = Engine(weights0, device=0, max_batch=130, max_length=T)
    small = Engine(weights0, device=0, max_batch=8, max_length=T)
    try:
        big.set_option("big_rows", 4096)             # the small-batch program at 130 rows
        ids_big, lens_big = big.recognize(crops)
        for lo in (0, 56, 64, 122):
            ids_small, _ = small.recognize(crops[lo:lo + 8])
            assert np.array_equal(ids_big[lo:lo + 8], ids_small), lo
        # the other decoder programs agree on this batch too: split-K partials + LayerNorm stages instead of the fused
        # cluster kernel, and the large-batch program (every Linear on the tcgen05 kernel), which larger batches select
        for key, val in (("fuse_ln", 0), ("big_rows", 64)):
            big.set_option(key, val)
            ids_m, _ = big.recognize(crops)
            agree = (ids_m == ids_big).mean()
            assert agree > 0.9, (key, agree)        # near-ties may flip a token and everything after it
            if key == "fuse_ln":
                big.set_option(key, 1)
        ids_m2, _ = big.recognize(crops)             # the large-batch program is deterministic and batch-invariant too
        assert np.array_equal(ids_m, ids_m2)
        small.set_option("big_rows", 1)
        for lo in (0, 122):
            ids_small, _ = small.recognize(crops[lo:lo + 8])
            assert np.array_equal(ids_m[lo:lo + 8], ids_small), lo
    finally:
        big.close()
        small.close()


def test_slot_refill_gives_the_same_ids_in_fewer_steps():
    """In-flight slot refill (option "slots"): with fewer decoder rows than crops, a row that finishes takes the next
    waiting crop.  The reference instead pads finished rows and steps the whole batch until its longest sequence ends
    (generation/utils.py:2797-2805); rows are independent, so ids and lengths must be EXACTLY those of one row per crop,
    in a number of steps that follows the total token count."""
    from manga_ocr_b200 import weights as W
    from manga_ocr_b200.engine import Engine
    T = 40
    w = W.random_init(0, gain=3.0, eos_bias=3.7)        # on these crops: a mix of very short rows and rows that never emit EOS
    crops = C.bubble_batch(45, seed=91)
    eng = Engine(w, device=0, max_batch=48, max_length=T)
    try:
        eng.set_option("check_every", 4)
        ids_ref, lens_ref = eng.recognize(crops)
        steps_ref = eng.last_steps
        print("lens:", sorted(lens_ref.tolist()))
        assert int(lens_ref.max()) == T and int(lens_ref.min()) < T // 2 and len(set(lens_ref.tolist())) >= 3      # ragged lengths
        for slots in (8, 13, 44, 64):
            eng.set_option("slots", slots)
            ids, lens = eng.recognize(crops)
            assert np.array_equal(lens, lens_ref), slots
            assert np.array_equal(ids, ids_ref), slots
            if slots == 8:
                tokens = int(lens_ref.sum() - len(crops))
                assert eng.last_steps < (T - 1) * 6                  # the padded scheme in chunks of 8 rows: 6 x 39 steps
                print("slot refill: 45 crops,", tokens, "tokens, 8 rows:", eng.last_steps, "steps; one row per crop:", steps_ref, "steps")
        # the same with the encoder serialised in front of the decode (no second stream)
        eng.set_option("slots", 8)
        eng.set_option("pipeline", 0)
        ids, lens = eng.recognize(crops)
        assert np.array_equal(ids, ids_ref) and np.array_equal(lens, lens_ref)
        eng.set_option("pipeline", 1)
        # more crops than the handle takes at once: chunks of 48, each with its own queue
        many = crops + C.bubble_batch(30, seed=92)
        eng.set_option("slots", 0)
        ids_many_ref, lens_many_ref = eng.recognize(many)
        eng.set_option("slots", 8)
        ids_many, lens_many = eng.recognize(many)
        assert np.array_equal(ids_many, ids_many_ref) and np.array_equal(lens_many, lens_many_ref)
        # weights that never emit EOS: every crop runs to max_length, rows are handed over at the length limit
        eng.set_option("slots", 0)
    finally:
        eng.close()
    eng = Engine(W.random_init(0), device=0, max_batch=20, max_length=12)
    try:
        ids_ref, lens_ref = eng.recognize(crops[:20])
        eng.set_option("slots", 6)
        eng.set_option("pipeline", 0)
        ids, lens = eng.recognize(crops[:20])
        assert np.array_equal(ids, ids_ref) and np.array_equal(lens, lens_ref) and (lens == 12).all()
        assert eng.last_steps == 11 * 4                              # ceil(20 / 6) rounds of 11 steps
        eng.set_option("pipeline", 1)                                # a decode far shorter than the encoder of the waiting crops
        ids, lens = eng.recognize(crops[:20])
        assert np.array_equal(ids, ids_ref) and np.array_equal(lens, lens_ref)
    finally:
        eng.close()


def test_chunked_staging_gives_the_same_ids(weights0):
    """Large batches are staged in chunks on a copy stream, each chunk preprocessed and encoded as it arrives (engine.cu:
    stage_encode): ids and lengths equal the one-piece path, chunk boundaries inside and at the end of the batch included."""
    from manga_ocr_b200.engine import Engine
    crops = C.page_batch(300, seed=21)
    T = 8
    eng = Engine(weights0, device=0, max_batch=300, max_length=T)
    try:
        eng.set_option("stage_chunk", 0)
        ids0, lens0 = eng.recognize(crops)
        for chunk in (128, 100, 150):
            eng.set_option("stage_chunk", chunk)
            ids1, lens1 = eng.recognize(crops)
            assert np.array_equal(ids0, ids1) and np.array_equal(lens0, lens1), chunk
            ids2, _ = eng.recognize(crops[:257])          # a second batch on the same handle, ragged last chunk
            assert np.array_equal(ids2, ids0[:257]), chunk
    finally:
        eng.close()


def test_session_admission_gives_the_same_ids():
    """Admission into a running decode (include/mocr_b200.h mocr_session_*): crops added while other rows are stepping, more
    crops than slots over time (slots are reused), fewer decoder rows than slots - every crop's ids equal mocr_recognize's."""
    from manga_ocr_b200 import weights as W
    from manga_ocr_b200.engine import Engine, MocrError
    w = W.random_init(0, gain=3.0, eos_bias=3.7)          # ragged lengths: rows finish at different steps
    T = 24
    crops = C.page_batch(40, seed=5)
    eng = Engine(w, device=0, max_batch=16, max_length=T)
    try:
        ref = np.concatenate([eng.recognize(crops[i:i + 16])[0] for i in range(0, 40, 16)])
        eng.session_begin(rows=6, max_length=T)
        with pytest.raises(MocrError):
            eng.recognize(crops[:2])                        # the handle is the session's
        pending = list(range(40))
        inflight, got = {}, {}
        rounds = 0
        while pending or inflight:
            free = 16 - len(inflight)
            take = min(free, len(pending), 1 + rounds % 5)  # trickle: 1..5 crops per round
            if take:
                idx = [pending.pop(0) for _ in range(take)]
                for i, s in zip(idx, eng.session_add([crops[i] for i in idx])):
                    assert int(s) not in inflight
                    inflight[int(s)] = i
            lens = eng.session_run(3)
            done = [s for s in inflight if lens[s] > 0]
            ids = eng.session_fetch(done)
            for s, row in zip(done, ids):
                i = inflight.pop(s)
                got[i] = row
                assert lens[s] == (row != 0).sum()
            rounds += 1
            assert rounds < 2000
        eng.session_end()
        for i in range(40):
            assert np.array_equal(got[i], ref[i]), i
        # the dispatcher's pipelined form: a chunk is launched before the snapshot of the previous one is read, slots are reused
        # while older snapshots (that still show their previous occupant's length) are in flight
        eng.session_begin(rows=16, max_length=T)
        pending_snaps, pending, inflight, got = 0, list(range(40)) * 2, {}, {}
        order = []
        while pending or inflight:
            launched = bool(inflight)
            if launched:
                eng.session_run(2, wait=False)
                pending_snaps += 1
            take = min(16 - len(inflight), len(pending), 7)
            if take:
                idx = [pending.pop(0) for _ in range(take)]
                for i, s in zip(idx, eng.session_add([crops[i] for i in idx])):
                    inflight[int(s)] = (i, len(order))
                    order.append(i)
            if inflight and not launched:
                eng.session_run(2, wait=False)
                pending_snaps += 1
            if pending_snaps >= 2:
                lens = eng.session_run(0)
                pending_snaps -= 1
                done = [s for s in inflight if lens[s] > 0]
                for s, row in zip(done, eng.session_fetch(done)):
                    i, k = inflight.pop(s)
                    got[k] = (i, row, int(lens[s]))
        eng.session_end()
        assert len(got) == 80
        for k, (i, row, n) in got.items():
            assert np.array_equal(row, ref[i]) and n == (ref[i] != 0).sum(), (k, i)
        # one admission into NON-ADJACENT free slots: a single encoder pass whose cross-K/V rows are scattered through the slot
        # map; the occupants of the slots in between keep their K/V (their rows are still decoding when the pass runs)
        eng.session_begin(rows=16, max_length=T)
        first = [int(x) for x in eng.session_add(crops[:8])]
        assert first == list(range(8))
        held = {}
        for _ in range(T):
            lens = eng.session_run(1)
            for sl in (1, 3, 6):
                if sl not in held and lens[sl] > 0:
                    held[sl] = eng.session_fetch([sl])[0]          # released as soon as it ends, the others keep stepping
            if len(held) == 3:
                break
        assert len(held) == 3
        second = [int(x) for x in eng.session_add(crops[8:11])]
        assert second == [1, 3, 6]
        rows_a, rows_b = {}, {}
        for _ in range(2 * T):
            lens = eng.session_run(1)
            for sl in range(8):
                if lens[sl] > 0 and sl not in rows_b and (sl in (1, 3, 6) or sl not in rows_a):
                    (rows_b if sl in (1, 3, 6) else rows_a)[sl] = eng.session_fetch([sl], release=False)[0]
            if len(rows_a) == 5 and len(rows_b) == 3:
                break
        eng.session_end()
        for sl in range(8):
            assert np.array_equal(held[sl] if sl in held else rows_a[sl], ref[sl]), sl
        for k, sl in enumerate((1, 3, 6)):
            assert np.array_equal(rows_b[sl], ref[8 + k]), sl
        ids2, _ = eng.recognize(crops[:5])                  # the handle is usable again
        assert np.array_equal(ids2, ref[:5])
        with pytest.raises(MocrError):
            eng.session_run(1)
        eng.session_begin(rows=4, max_length=T)
        with pytest.raises(MocrError):
            eng.set_option("steps_per_graph", 5)            # the session's graph is in use
        eng.session_end()
    finally:
        eng.close()


def test_session_row_count_grows_with_the_load():
    """mocr_session_rows: a session of 40 rows steps its first 16 rows while lightly loaded and all of them later; the state is
    shared (crops admitted under one program finish under the other), ids per crop equal mocr_recognize's; shrinking is refused
    while slots are in use."""
    from manga_ocr_b200 import weights as W
    from manga_ocr_b200.engine import Engine, MocrError
    w = W.random_init(0, gain=3.0, eos_bias=3.7)
    T = 24
    crops = C.page_batch(40, seed=5)
    eng = Engine(w, device=0, max_batch=40, max_length=T)
    try:
        ref, _ = eng.recognize(crops)
        eng.session_begin(rows=40, max_length=T)
        assert eng.session_rows(16) == 16
        slots = {int(s): i for i, s in enumerate(eng.session_add(crops[:10]))}
        got = {}

        def collect(lens):
            done = [s for s in slots if lens[s] > 0]
            for s, row in zip(done, eng.session_fetch(done)):
                got[slots.pop(s)] = row
        collect(eng.session_run(3))                       # (3 steps requested = one graph of 13: the short texts end here)
        with pytest.raises(MocrError):
            eng.session_rows(41)
        assert eng.session_rows(40) == 40                 # more load: every row steps from the next chunk on
        if slots:
            with pytest.raises(MocrError):
                eng.session_rows(16)                      # crops are in flight
        for i, s in enumerate(eng.session_add(crops[10:])):
            slots[int(s)] = 10 + i
        for _ in range(8):
            collect(eng.session_run(3))
            if not slots:
                break
        assert not slots and len(got) == 40
        assert eng.session_rows(16) == 16                 # idle again: back to the small program, and it still works
        slots = {int(s): 3 + i for i, s in enumerate(eng.session_add(crops[3:6]))}
        for _ in range(4):
            collect(eng.session_run(3))
        eng.session_end()
        assert not slots
        for i in range(40):
            assert np.array_equal(got[i], ref[i]), i
    finally:
        eng.close()
