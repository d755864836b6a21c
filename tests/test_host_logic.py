"""Host-side logic of the drop-in class that needs no GPU: argument contract, image conversion,
checkpoint discovery, the job splitter and its final gather (gloo, world_size 2)."""
import os
import sys

import numpy as np
import pytest
from PIL import Image

from manga_ocr_b200 import ocr as O
from manga_ocr_b200.splitter import shard_bounds, shard_sizes

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shim_exports_mangaocr():
    import manga_ocr
    assert manga_ocr.MangaOcr is O.MangaOcr
    import inspect
    sig = inspect.signature(O.MangaOcr.__init__)
    names = list(sig.parameters)
    assert names[1:3] == ["pretrained_model_name_or_path", "force_cpu"]
    assert sig.parameters["pretrained_model_name_or_path"].default == "kha-white/manga-ocr-base"
    assert sig.parameters["force_cpu"].default is False


def test_force_cpu_is_refused():
    with pytest.raises(RuntimeError, match="no CPU path"):
        O.MangaOcr(force_cpu=True)


def test_missing_checkpoint_raises(tmp_path, monkeypatch):
    monkeypatch.setenv("HF_HOME", str(tmp_path))
    monkeypatch.delenv("MOCR_WEIGHTS", raising=False)
    with pytest.raises(FileNotFoundError):
        O.MangaOcr()      # the app treats any exception as "engine unavailable" (main_window.py:3396-3398)


def test_find_checkpoint(tmp_path, monkeypatch):
    d = tmp_path / "ckpt"
    d.mkdir()
    (d / "model.safetensors").write_bytes(b"x")
    (d / "vocab.txt").write_text("a\n")
    assert O._find_checkpoint(str(d)) == (str(d / "model.safetensors"), str(d / "vocab.txt"))
    monkeypatch.setenv("HF_HOME", str(tmp_path / "hf"))
    snap = tmp_path / "hf" / "hub" / "models--kha-white--manga-ocr-base" / "snapshots" / "abc"
    snap.mkdir(parents=True)
    (snap / "model.safetensors").write_bytes(b"x")
    assert O._find_checkpoint("kha-white/manga-ocr-base") == (str(snap / "model.safetensors"), None)
    # the snapshot refs/main points to wins over the lexicographically first one, and a snapshot that only
    # ships pytorch_model.bin (what from_pretrained would load) is accepted
    main = snap.parent / "zzz"
    main.mkdir()
    (main / "pytorch_model.bin").write_bytes(b"x")
    refs = snap.parent.parent / "refs"
    refs.mkdir()
    (refs / "main").write_text("zzz\n")
    assert O._find_checkpoint("kha-white/manga-ocr-base") == (str(main / "pytorch_model.bin"), None)


def test_torch_bin_checkpoint_is_read_without_torch(tmp_path):
    """pytorch_model.bin (torch.save zip) -> float32 arrays; written here WITH torch, read by the product WITHOUT it."""
    import numpy as np
    torch = pytest.importorskip("torch")
    from manga_ocr_b200 import weights as W
    sd = {"a.weight": torch.randn(5, 7), "half": torch.randn(3).half(), "bf": torch.randn(4, 4).bfloat16(),
          "position_ids": torch.arange(10)[None], "transposed": torch.randn(6, 4).t(), "param": torch.nn.Parameter(torch.randn(2, 3))}
    torch.save(sd, tmp_path / "pytorch_model.bin")
    out = W.load_torch_bin(str(tmp_path / "pytorch_model.bin"))
    assert "position_ids" not in out
    for k, v in sd.items():
        if v.is_floating_point():
            assert out[k].dtype == np.float32 and np.array_equal(out[k], v.detach().float().numpy()), k
    import pickle

    class Evil:
        def __reduce__(self):
            return (print, ("pwned",))
    import zipfile
    with zipfile.ZipFile(tmp_path / "evil.bin", "w") as z:
        z.writestr("archive/data.pkl", pickle.dumps({"x": Evil()}))
    with pytest.raises(pickle.UnpicklingError):
        W.load_torch_bin(str(tmp_path / "evil.bin"))


def test_image_to_array_modes():
    rng = np.random.default_rng(0)
    rgb = rng.integers(0, 256, (9, 7, 3), dtype=np.uint8)
    assert np.array_equal(O.image_to_array(Image.fromarray(rgb)), rgb)
    assert O.image_to_array(Image.fromarray(rgb[..., 0])).shape == (9, 7)
    assert O.image_to_array(Image.fromarray(rgb).convert("RGBA")).shape == (9, 7, 4)
    # what reaches the engine has the L plane upstream's img.convert("L") gives, for every Pillow mode: RGB / RGBA bytes go to
    # the GPU's luma conversion, the other modes are converted by Pillow itself (YCbCr -> L is the Y plane, NOT the luma of the
    # image's RGB expansion: 42 % of the pixels differ by 1)
    from oracle.preprocess_np import rgb_to_l
    base = Image.fromarray(rng.integers(0, 256, (33, 21, 3), dtype=np.uint8))
    imgs = [base.convert(m) for m in ("RGB", "RGBA", "L", "P", "PA", "1", "LA", "CMYK", "YCbCr", "HSV", "I", "F", "RGBX")]
    imgs.append(Image.fromarray(rng.integers(0, 65536, (33, 21), dtype=np.uint16)))              # I;16
    imgs.append(Image.fromarray(rng.integers(-500, 70000, (33, 21)).astype(np.int32)))           # I, out of the uint8 range
    imgs.append(Image.fromarray((rng.random((33, 21)) * 600 - 100).astype(np.float32)))          # F, likewise
    for im in imgs:
        a = O.image_to_array(im)
        assert a.dtype == np.uint8 and a.shape[:2] == (33, 21), im.mode
        plane = a if a.ndim == 2 else rgb_to_l(a[..., :3])
        assert np.array_equal(plane, np.asarray(im.convert("L"))), im.mode


def test_shard_bounds_cover_exactly():
    for n in (0, 1, 7, 64, 512, 4096, 4099):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_bounds(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = shard_sizes(n, world)
            assert sum(sizes) == n and max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_bounds(4, 2, 2)


def _gather_worker(rank, world, port, n_total, T, q):
    import torch.distributed as dist
    from manga_ocr_b200.splitter import gather_ids, shard_bounds
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    full = np.arange(n_total * T, dtype=np.int32).reshape(n_total, T)
    lo, hi = shard_bounds(n_total, world, rank)
    out = gather_ids(full[lo:hi], n_total)
    if rank == 0:
        q.put(bool(np.array_equal(out, full)))
    else:
        q.put(out is None)
    dist.destroy_process_group()


@pytest.mark.parametrize("n_total", [5, 8])
def test_gather_ids_gloo_world2(n_total):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() + n_total) % 2000
    ps = [ctx.Process(target=_gather_worker, args=(r, 2, port, n_total, 6, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = [q.get(timeout=120) for _ in ps]
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(res)


def test_hub_cache_resolution_equals_huggingface_hub(tmp_path, monkeypatch):
    """MangaOcr() with no argument finds kha-white/manga-ocr-base where from_pretrained would look for it offline: the snapshot
    `refs/main` names, in the cache directory huggingface_hub resolves from HF_HUB_CACHE / HUGGINGFACE_HUB_CACHE / HF_HOME /
    XDG_CACHE_HOME - compared with huggingface_hub.snapshot_download(local_files_only=True) on a synthetic cache."""
    import huggingface_hub as hh
    root = tmp_path / "cache" / "huggingface" / "hub"
    repo = root / "models--kha-white--manga-ocr-base"
    for rev in ("0aaa", "1bbb", "2ccc"):
        (repo / "snapshots" / rev).mkdir(parents=True)
        (repo / "snapshots" / rev / "model.safetensors").write_bytes(b"x")
        (repo / "snapshots" / rev / "config.json").write_text("{}")
    (repo / "refs").mkdir()
    (repo / "refs" / "main").write_text("1bbb")
    want = hh.snapshot_download("kha-white/manga-ocr-base", cache_dir=str(root), local_files_only=True)
    assert want.endswith("1bbb")
    for var in ("HF_HUB_CACHE", "HUGGINGFACE_HUB_CACHE", "HF_HOME", "XDG_CACHE_HOME", "MOCR_WEIGHTS"):
        monkeypatch.delenv(var, raising=False)
    for var, value in (("HF_HUB_CACHE", root), ("HUGGINGFACE_HUB_CACHE", root), ("HF_HOME", root.parent), ("XDG_CACHE_HOME", root.parent.parent)):
        monkeypatch.setenv(var, str(value))
        assert O._find_checkpoint("kha-white/manga-ocr-base") == (os.path.join(want, "model.safetensors"), None), var
        monkeypatch.delenv(var)
    monkeypatch.setenv("HF_HUB_CACHE", str(tmp_path / "empty"))
    monkeypatch.setenv("HF_HOME", str(root.parent))                    # HF_HUB_CACHE wins over HF_HOME, as in huggingface_hub
    assert O._find_checkpoint("kha-white/manga-ocr-base") is None


def test_generation_config_is_read_from_the_checkpoint_directory(tmp_path):
    """generate() uses the checkpoint's generation settings; no file -> greedy."""
    import json
    from manga_ocr_b200.ocr import GREEDY, _generation_config
    assert _generation_config(None) == GREEDY
    assert _generation_config(str(tmp_path / "model.safetensors")) == GREEDY
    (tmp_path / "config.json").write_text(json.dumps({"num_beams": 2, "length_penalty": 1.5, "model_type": "vision-encoder-decoder"}))
    assert _generation_config(str(tmp_path))["num_beams"] == 2
    # generation_config.json is used wholesale when it exists (not merged over config.json: length_penalty stays at its default)
    (tmp_path / "generation_config.json").write_text(json.dumps({"num_beams": 4, "no_repeat_ngram_size": 3, "early_stopping": True}))
    g = _generation_config(str(tmp_path / "model.safetensors"))
    assert g == {"num_beams": 4, "no_repeat_ngram_size": 3, "length_penalty": 1.0, "early_stopping": True}


def test_generation_config_resolution_equals_transformers(tmp_path):
    """The generation settings the product reads from a checkpoint directory are the ones transformers' own generate() would
    use for it: a tiny VisionEncoderDecoderModel is saved, its json files are edited into every combination (generation_config.json
    alone / flagged _from_model_config / next to legacy fields in config.json; legacy fields at the top level, in the decoder
    sub-config, in both; none) and `model._prepare_generation_config()` of the reloaded model is the reference
    (transformers/generation/utils.py, configuration_utils.py::from_model_config)."""
    import json
    import warnings
    from transformers import BertConfig, ViTConfig, VisionEncoderDecoderConfig, VisionEncoderDecoderModel
    from manga_ocr_b200.ocr import _generation_config
    enc = ViTConfig(hidden_size=32, num_hidden_layers=1, num_attention_heads=2, intermediate_size=64, image_size=32, patch_size=16)
    dec = BertConfig(hidden_size=32, num_hidden_layers=1, num_attention_heads=2, intermediate_size=64, vocab_size=50, is_decoder=True,
                     add_cross_attention=True)
    d = tmp_path / "ckpt"
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        m = VisionEncoderDecoderModel(VisionEncoderDecoderConfig.from_encoder_decoder_configs(enc, dec))
        m.generation_config.decoder_start_token_id, m.generation_config.eos_token_id, m.generation_config.pad_token_id = 2, 3, 0
        m.save_pretrained(str(d))
    cfg0 = json.loads((d / "config.json").read_text())
    base = {"decoder_start_token_id": 2, "eos_token_id": 3, "pad_token_id": 0}
    beams = {"num_beams": 4, "no_repeat_ngram_size": 3, "early_stopping": True}

    def with_decoder(extra_top, extra_dec):
        c = json.loads(json.dumps(cfg0))
        c.update(extra_top)
        c["decoder"].update(extra_dec)
        return c

    cases = {                                      # name: (generation_config.json or None, config.json)
        "generation_config alone": ({**base, **beams}, cfg0),
        "generation_config wins over legacy fields, wholesale": ({**base, **beams}, with_decoder({"num_beams": 2, "length_penalty": 1.5}, {})),
        "generation_config flagged _from_model_config": ({**base, **beams, "_from_model_config": True}, with_decoder({"num_beams": 2}, {})),
        "legacy fields at the top level": (None, with_decoder({"num_beams": 2, "length_penalty": 1.5}, {})),
        "legacy fields in the decoder sub-config": (None, with_decoder({}, {"num_beams": 3, "no_repeat_ngram_size": 2, "length_penalty": 2.0,
                                                                            "early_stopping": True})),
        "top level first, decoder per missing attribute": (None, with_decoder({"num_beams": 5}, {"num_beams": 3, "no_repeat_ngram_size": 2})),
        "no generation fields": (None, cfg0),
    }
    for name, (gen_json, cfg_json) in cases.items():
        (d / "config.json").write_text(json.dumps(cfg_json))
        if gen_json is None:
            (d / "generation_config.json").unlink(missing_ok=True)
        else:
            (d / "generation_config.json").write_text(json.dumps(gen_json))
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            gc, _ = VisionEncoderDecoderModel.from_pretrained(str(d))._prepare_generation_config(None)
        want = {"num_beams": gc.num_beams or 1, "no_repeat_ngram_size": gc.no_repeat_ngram_size or 0,
                "length_penalty": 1.0 if gc.length_penalty is None else gc.length_penalty, "early_stopping": gc.early_stopping or False}
        assert _generation_config(str(d)) == want, (name, _generation_config(str(d)), want)
        assert _generation_config(str(d / "model.safetensors")) == want, name


def test_unsupported_generation_settings_are_refused_not_ignored(tmp_path, monkeypatch):
    """A checkpoint whose generation settings ask for something this engine does not implement (repetition penalty, sampling,
    bad words ...) must not decode silently differently from generate(): the constructor raises (the app treats any exception as
    "engine unavailable", main_window.py:3396-3398); neutral values, as old config.json files spell them out, are fine."""
    import json
    d = tmp_path / "ckpt"
    d.mkdir()
    (d / "model.safetensors").write_bytes(b"x")
    neutral = {"do_sample": False, "num_beam_groups": 1, "diversity_penalty": 0.0, "repetition_penalty": 1.0, "min_length": 0,
               "bad_words_ids": None, "num_return_sequences": 1, "encoder_no_repeat_ngram_size": 0, "temperature": 1.0, "top_k": 50,
               "forced_eos_token_id": None, "suppress_tokens": None, "max_length": 20, "num_beams": 4, "no_repeat_ngram_size": 3}
    (d / "config.json").write_text(json.dumps({"model_type": "vision-encoder-decoder", "decoder": neutral}))
    assert O._unsupported_generation_settings(str(d)) == {}
    assert O._generation_config(str(d))["num_beams"] == 4
    (d / "generation_config.json").write_text(json.dumps({"num_beams": 4, "repetition_penalty": 1.2, "bad_words_ids": [[7]], "min_length": 0}))
    assert O._unsupported_generation_settings(str(d / "model.safetensors")) == {"repetition_penalty": 1.2, "bad_words_ids": [[7]]}
    (d / "generation_config.json").write_text(json.dumps({"decoder_start_token_id": 2, "eos_token_id": 3, "pad_token_id": 0}))
    assert O._unsupported_generation_settings(str(d)) == {}
    (d / "generation_config.json").write_text(json.dumps({"decoder_start_token_id": 101, "eos_token_id": [3, 102]}))
    assert O._unsupported_generation_settings(str(d)) == {"decoder_start_token_id": 101, "eos_token_id": [3, 102]}
    (d / "generation_config.json").write_text(json.dumps({"num_beams": 4, "repetition_penalty": 1.2, "bad_words_ids": [[7]], "min_length": 0}))
    _StubEngine.instances = []
    monkeypatch.setattr(O, "Engine", _StubEngine)
    monkeypatch.setattr(O.W, "complete", lambda w: w)
    monkeypatch.setattr(O.W, "load_weights", lambda path: {"x": np.zeros(1, np.float32)})
    with pytest.raises(NotImplementedError, match="repetition_penalty"):
        O.MangaOcr(str(d), warmup=False)
    monkeypatch.setenv("MOCR_IGNORE_GENERATION_EXTRAS", "1")
    ocr = O.MangaOcr(str(d), warmup=False)
    assert ocr.generation["num_beams"] == 4
    ocr.close()


# ---- the cross-thread micro-batcher, with stub engines (no GPU): batching, fair share, failure isolation, lifetime ----

class _StubEngine:
    """Stands in for manga_ocr_b200.engine.Engine: ids row = [CLS, marker of the crop, SEP]; a crop whose first pixel is 13
    makes the whole call fail, like one malformed crop fails a library call."""
    instances = []

    def __init__(self, weights, device=0, max_batch=64, max_length=300):
        self.device, self.max_batch, self.max_length = device, max_batch, max_length
        self.batches = []
        self.closed = False
        _StubEngine.instances.append(self)

    def set_option(self, key, value):
        pass

    def recognize(self, arrays, order=0, max_length=None):
        import time
        self.batches.append(len(arrays))
        time.sleep(0.02)                       # a GPU batch takes a while: callers pile up behind it
        if any(int(a.flat[0]) == 13 for a in arrays):
            raise RuntimeError("bad crop in batch")
        T = max_length or self.max_length
        ids = np.zeros((len(arrays), T), np.int32)
        ids[:, 0] = 2
        for i, a in enumerate(arrays):
            ids[i, 1] = 5 + int(a.flat[0])
            ids[i, 2] = 3
        return ids, np.full(len(arrays), 3, np.int32)

    # ---- sessions (admission into a running decode): a crop finishes `need` calls of session_run after it was added; need =
    #      1 + (first pixel % 3), so crops added together finish at different times
    def session_begin(self, rows, order=0, max_length=None):
        if getattr(self, "refuse_sessions", False):
            raise RuntimeError("sessions do not run with parity taps")
        assert not getattr(self, "sess", None), "session already active"
        self.sess = {"T": max_length or self.max_length, "slots": {}, "rows": rows}
        self.sessions = getattr(self, "sessions", 0) + 1
        self.adds = getattr(self, "adds", [])

    def session_add(self, arrays):
        import time
        time.sleep(0.002)
        if any(int(a.flat[0]) == 13 for a in arrays):
            raise RuntimeError("bad crop in batch")
        free = [s for s in range(self.max_batch) if s not in self.sess["slots"]]
        assert len(arrays) <= len(free)
        self.adds.append(len(arrays))
        out = []
        for a, s in zip(arrays, free):
            self.sess["slots"][s] = [int(a.flat[0]), 1 + int(a.flat[0]) % 3]
            out.append(s)
        return np.asarray(out, np.int32)

    def session_rows(self, rows):
        assert getattr(self, "sess", None) is not None
        now = 16 if rows <= 16 < self.sess["rows"] else self.sess["rows"]
        if now < self.sess.get("rows_now", self.sess["rows"]):
            assert not self.sess["slots"], "the row count shrinks only while no slot is in use"
        self.sess["rows_now"] = now
        self.row_counts = getattr(self, "row_counts", []) + [now]
        return now

    def session_run(self, steps, wait=True):
        import time
        if not wait:                               # launch only (+ a length snapshot): read later with session_run(0)
            self.pending = getattr(self, "pending", 0) + 1
            assert self.pending <= 2, "more than two snapshots pending"
            return None
        if steps > 0:
            self.pending = getattr(self, "pending", 0) + 1
        assert getattr(self, "pending", 0) > 0, "poll without a launch"
        self.pending -= 1
        time.sleep(0.005)
        lens = np.zeros((self.max_batch,), np.int32)
        for s, st in self.sess["slots"].items():
            st[1] -= 1
            if st[1] <= 0:
                lens[s] = 3
        return lens

    def session_fetch(self, slots, release=True):
        ids = np.zeros((len(slots), self.sess["T"]), np.int32)
        for i, s in enumerate(slots):
            ids[i, :3] = (2, 5 + self.sess["slots"][s][0], 3)
            if release:
                del self.sess["slots"][s]
        return ids

    def session_end(self):
        assert not self.sess["slots"], "session ended with crops in flight"
        self.sess = None
        self.pending = 0

    def close(self):
        self.closed = True


@pytest.fixture
def stub_ocr(monkeypatch):
    _StubEngine.instances = []
    monkeypatch.setattr(O, "Engine", _StubEngine)
    made = []

    def make(**kw):
        kw.setdefault("warmup", False)
        kw.setdefault("admission", False)          # the batch dispatcher unless a test asks for sessions
        o = O.MangaOcr(weights={"x": np.zeros(1, np.float32)}, **kw)
        made.append(o)
        return o

    monkeypatch.setattr(O.W, "complete", lambda w: w)
    yield make
    for o in made:
        o.close()


def _img(v):
    return Image.fromarray(np.full((4, 4, 3), v, np.uint8))


def _call_all(ocr, values):
    import threading
    out = {}

    def run(i, v):
        try:
            out[i] = ocr(_img(v))
        except Exception as e:      # noqa: BLE001
            out[i] = e

    ts = [threading.Thread(target=run, args=(i, v)) for i, v in enumerate(values)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    return [out[i] for i in range(len(values))]


def test_micro_batcher_gathers_concurrent_callers(stub_ocr):
    ocr = stub_ocr(devices=[0], max_batch=16, max_length=8, linger_ms=20)
    vocab = ocr.vocab
    got = _call_all(ocr, list(range(12)))
    assert got == [O.ids_to_texts(vocab, np.array([[2, 5 + v, 3, 0]]))[0] for v in range(12)]
    eng = _StubEngine.instances[0]
    assert sum(eng.batches) == 12 and max(eng.batches) >= 6, eng.batches      # not twelve batches of one


def test_micro_batcher_linger_extends_while_callers_keep_arriving(stub_ocr):
    """Callers that trickle in (a worker pool resubmitting: one every 8 ms for 64 ms, linger 60 ms) still share one batch: every
    arrival extends the wait by a third of the linger; a lone caller is not held longer than the linger."""
    import threading
    import time
    ocr = stub_ocr(devices=[0], max_batch=16, max_length=8, linger_ms=60)
    out = {}

    def run(i):
        time.sleep(0.008 * i)
        out[i] = ocr(_img(i))

    ts = [threading.Thread(target=run, args=(i,)) for i in range(9)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    eng = _StubEngine.instances[0]
    assert sum(eng.batches) == 9 and max(eng.batches) >= 8, eng.batches        # 9 arrivals over 64 ms: not cut off after 60 ms (8 of them are in by then)
    n_before = len(eng.batches)
    t0 = time.monotonic()
    ocr(_img(3))
    assert time.monotonic() - t0 < 0.2 and len(eng.batches) == n_before + 1    # alone: one linger (60 ms), not the 4x cap (240 ms)


def test_micro_batcher_shares_the_queue_between_gpus(stub_ocr):
    ocr = stub_ocr(devices=[0, 1], max_batch=64, max_length=8, linger_ms=20)
    got = _call_all(ocr, [v for v in range(41) if v != 13])
    assert all(isinstance(g, str) for g in got)
    a, b = (sum(e.batches) for e in _StubEngine.instances)
    assert a + b == 40 and min(a, b) >= 8, (a, b)                              # nobody grabbed the whole queue


def test_micro_batcher_isolates_a_failing_request(stub_ocr):
    ocr = stub_ocr(devices=[0], max_batch=16, max_length=8, linger_ms=20)
    got = _call_all(ocr, [1, 2, 13, 4, 5])
    assert isinstance(got[2], RuntimeError)
    assert [isinstance(g, str) for g in got] == [True, True, False, True, True]
    assert ocr(_img(7)) == got[0][:0] + O.ids_to_texts(ocr.vocab, np.array([[2, 12, 3]]))[0]      # still usable
    with pytest.raises(ValueError):
        ocr(Image.fromarray(np.zeros((3, 40000, 3), np.uint8)))                # refused before it is queued


def test_instance_lifetime_close_and_collection(stub_ocr):
    import gc
    import weakref
    ocr = stub_ocr(devices=[0], max_batch=4, max_length=8, linger_ms=0)
    assert isinstance(ocr(_img(3)), str)
    ocr.close()
    assert _StubEngine.instances[0].closed and all(not t.is_alive() for t in ocr._threads)
    with pytest.raises(RuntimeError):
        ocr(_img(3))
    ocr.close()                                                                # idempotent
    # an instance that is dropped without close() is collected: its dispatcher threads hold it weakly
    _StubEngine.instances = []
    o2 = O.MangaOcr(weights={"x": np.zeros(1, np.float32)}, devices=[0], max_batch=4, max_length=8, warmup=False)
    threads = list(o2._threads)
    ref = weakref.ref(o2)
    del o2
    gc.collect()
    assert ref() is None
    for t in threads:
        t.join(timeout=5)
        assert not t.is_alive()
    assert _StubEngine.instances[0].closed
    with stub_ocr(devices=[0], max_batch=4, max_length=8) as o3:               # context manager
        assert isinstance(o3(_img(1)), str)
    assert o3._closed


# ---- admission into a running decode (MangaOcr(admission=True), the default for greedy decoding), with stub engines ----

def test_admission_answers_each_caller_when_its_crop_finishes(stub_ocr):
    ocr = stub_ocr(devices=[0], max_batch=8, max_length=8, admission=True)
    vocab = ocr.vocab
    got = _call_all(ocr, list(range(12)) + [20, 21])            # more callers than slots: slots are reused
    assert got == [O.ids_to_texts(vocab, np.array([[2, 5 + v, 3, 0]]))[0] for v in list(range(12)) + [20, 21]]
    eng = _StubEngine.instances[0]
    assert sum(eng.adds) == 14 and eng.sess is None                            # everything admitted, the session ended when idle
    assert ocr(_img(4)) == O.ids_to_texts(vocab, np.array([[2, 9, 3]]))[0]      # a new session starts on demand
    assert eng.sessions >= 2


def test_admission_isolates_a_failing_request_and_yields_to_batch_calls(stub_ocr):
    import threading
    ocr = stub_ocr(devices=[0], max_batch=8, max_length=8, admission=True)
    got = _call_all(ocr, [1, 2, 13, 4, 5])
    assert isinstance(got[2], RuntimeError)
    assert [isinstance(g, str) for g in got] == [True, True, False, True, True]
    # a batch call on the same engine waits for the session to drain and runs between sessions
    out = {}
    ts = [threading.Thread(target=lambda i=i: out.__setitem__(i, ocr(_img(i)))) for i in range(6)]
    for t in ts:
        t.start()
    texts = ocr.recognize_batch([np.full((4, 4, 3), 7, np.uint8), np.full((4, 4, 3), 8, np.uint8)])
    for t in ts:
        t.join()
    assert texts == [O.ids_to_texts(ocr.vocab, np.array([[2, 12, 3]]))[0], O.ids_to_texts(ocr.vocab, np.array([[2, 13, 3]]))[0]]
    assert all(isinstance(out[i], str) for i in range(6))
    assert _StubEngine.instances[0].sess is None


def test_admission_shares_callers_between_gpus_and_closes_cleanly(stub_ocr):
    ocr = stub_ocr(devices=[0, 1], max_batch=16, max_length=8, admission=True)
    got = _call_all(ocr, [v for v in range(41) if v != 13])
    assert all(isinstance(g, str) for g in got)
    a, b = (sum(e.adds) for e in _StubEngine.instances)
    assert a + b == 40 and min(a, b) >= 5, (a, b)
    ocr.close()
    assert all(e.closed and e.sess is None for e in _StubEngine.instances)
    with pytest.raises(RuntimeError):
        ocr(_img(1))


def test_admission_session_starts_on_16_rows_and_grows_with_the_load(stub_ocr):
    """A session steps only its first 16 decoder rows while at most 16 crops are in flight and all 64 once more arrive; the row
    count never shrinks while slots are in use (the stub asserts it), and the next session starts small again."""
    ocr = stub_ocr(devices=[0], max_batch=64, max_length=8, admission=True)
    eng = _StubEngine.instances[0]
    got = _call_all(ocr, [1, 2, 3, 4])
    assert all(isinstance(g, str) for g in got)
    assert set(eng.row_counts) == {16}                            # a few callers: never grown
    eng.row_counts = []
    got = _call_all(ocr, [v for v in range(60) if v != 13])
    assert all(isinstance(g, str) for g in got)
    assert eng.row_counts[0] == 16 and 64 in eng.row_counts       # 59 callers at once: grown
    for a, b in zip(eng.row_counts, eng.row_counts[1:]):
        assert not (a == 64 and b == 64)                          # once per session
    small = stub_ocr(devices=[0], max_batch=8, max_length=8, admission=True)
    assert isinstance(small(_img(2)), str)
    assert not getattr(_StubEngine.instances[-1], "row_counts", [])      # a session of at most 16 rows has one program


def test_admission_serves_a_worker_pool_that_resubmits(stub_ocr):
    """The app's shape: a few worker threads, each calling again as soon as it has its answer.  The dispatcher launches the next
    chunk, gives the callers it has just answered a moment to come back, and admits them in the same round; every call gets its
    own text and the session ends when the pool is done."""
    import threading
    ocr = stub_ocr(devices=[0], max_batch=8, max_length=8, admission=True)
    out, lock = {}, threading.Lock()

    def worker(w):
        for j in range(6):
            v = (7 * w + j) % 12
            t = ocr(_img(v))
            with lock:
                out[(w, j)] = (v, t)

    ts = [threading.Thread(target=worker, args=(w,)) for w in range(4)]
    for t in ts:
        t.start()
    for t in ts:
        t.join(timeout=30)
    assert len(out) == 24
    for v, t in out.values():
        assert t == O.ids_to_texts(ocr.vocab, np.array([[2, 5 + v, 3, 0]]))[0]
    eng = _StubEngine.instances[0]
    assert sum(eng.adds) == 24 and max(eng.adds) <= 4
    deadline = __import__("time").monotonic() + 2
    while eng.sess is not None and __import__("time").monotonic() < deadline:
        __import__("time").sleep(0.01)
    assert eng.sess is None


def test_ngram_ban_with_one_beam_goes_to_the_one_beam_search(stub_ocr):
    """generate() applies no_repeat_ngram_size in greedy mode too; the arg-max path has no ban list, so such settings are served by
    the search entry points with ONE beam and early_stopping=True (= greedy decoding with the ban, tests/test_beam_host.py), on
    every entry point, and __call__ traffic uses the batch dispatcher (sessions are arg-max only)."""
    calls = []

    def recognize_beam(self, arrays, order, max_length, *beam_args):
        calls.append(beam_args)
        return self.recognize(arrays, order, max_length)

    _StubEngine.recognize_beam = recognize_beam
    try:
        ocr = stub_ocr(devices=[0], max_batch=8, max_length=8, admission=None, num_beams=1, no_repeat_ngram_size=3)
        assert ocr._beam_args() == (1, 3, 1.0, True) and not ocr.admission
        assert len(ocr.recognize_batch([np.full((4, 4, 3), 7, np.uint8)])) == 1
        assert isinstance(ocr(_img(5)), str)
        assert calls == [(1, 3, 1.0, True)] * 2
        plain = stub_ocr(devices=[0], max_batch=8, max_length=8, admission=None)
        assert plain._beam_args() is None and plain.admission
        four = stub_ocr(devices=[0], max_batch=8, max_length=8, num_beams=4, no_repeat_ngram_size=3, length_penalty=2.0, early_stopping=True)
        assert four._beam_args() == (4, 3, 2.0, True) and not four.admission
    finally:
        del _StubEngine.recognize_beam


def test_admission_falls_back_to_a_batch_when_no_session_can_start(stub_ocr):
    """An engine that refuses sessions (parity taps set, as __graft_entry__.smoke() does) must not leave callers waiting."""
    ocr = stub_ocr(devices=[0], max_batch=8, max_length=8, admission=True)
    _StubEngine.instances[0].refuse_sessions = True
    got = _call_all(ocr, [1, 2, 13, 4])
    assert [isinstance(g, str) for g in got] == [True, True, False, True]
    _StubEngine.instances[0].refuse_sessions = False
    assert isinstance(ocr(_img(3)), str)
