"""``MangaOcr`` - drop-in for the class the app imports from the ``manga-ocr`` package
(reference/src/core/config.py:433), constructs once with no arguments
(reference/src/ui/main_window.py:3394) and calls per crop from up to 50 worker threads with no
lock (reference/src/ui/main_window.py:9800-9801, 608-611, 4317-4327).

Same constructor signature, same ``__call__(img_or_path) -> str``, same ``ValueError`` for a
bad argument; exceptions propagate (the workers catch them, reference/src/core/workers.py:
241-244).  Underneath, concurrent single-crop callers are gathered into one GPU batch (the
reference never batches: upstream runs ``x[None]``), one dispatcher thread per GPU.
"""
from __future__ import annotations

import glob
import contextlib
import os
import threading
import time
import weakref
from collections import deque
from pathlib import Path
from typing import Dict, List, Optional, Sequence

import numpy as np

from . import weights as W
from .engine import BGR, MAX_LENGTH, RGB, Engine
from .text import Vocab, ids_to_text, ids_to_texts

DEFAULT_MODEL = "kha-white/manga-ocr-base"


WEIGHT_FILES = ("model.safetensors", "pytorch_model.bin", "weights.npz")


def _hub_snapshots(name: str) -> List[str]:
    """Snapshot directories of ``name`` in the local Hugging Face cache, the one ``refs/main`` points to first.
    Nothing is ever downloaded: the reference relies on ``from_pretrained`` fetching the checkpoint, this engine needs
    it on disk already."""
    # the cache directory as huggingface_hub resolves it (huggingface_hub/constants.py): HF_HUB_CACHE, its legacy alias
    # HUGGINGFACE_HUB_CACHE, else $HF_HOME/hub with HF_HOME defaulting to $XDG_CACHE_HOME/huggingface (~/.cache/huggingface)
    env = os.environ
    hf_home = env.get("HF_HOME") or os.path.join(env.get("XDG_CACHE_HOME") or "~/.cache", "huggingface")
    home = env.get("HF_HUB_CACHE") or env.get("HUGGINGFACE_HUB_CACHE") or os.path.join(hf_home, "hub")
    home = os.path.expandvars(os.path.expanduser(home))
    repo = os.path.join(home, "models--" + name.replace("/", "--"))
    snaps = sorted(glob.glob(os.path.join(repo, "snapshots", "*")))
    try:
        with open(os.path.join(repo, "refs", "main"), encoding="utf-8") as f:
            main = os.path.join(repo, "snapshots", f.read().strip())
        if main in snaps:
            snaps.remove(main)
            snaps.insert(0, main)
    except OSError:
        pass
    return snaps


def _find_checkpoint(name_or_path: str):
    """Resolve a local checkpoint directory/file; returns (weights_path, vocab_path | None) or None."""
    cands: List[str] = []
    if name_or_path and os.path.exists(name_or_path):
        cands.append(name_or_path)
    env = os.environ.get("MOCR_WEIGHTS", "")
    if env and os.path.exists(env):
        cands.append(env)
    cands.extend(_hub_snapshots(name_or_path))
    for c in cands:
        if os.path.isdir(c):
            for fn in WEIGHT_FILES:
                p = os.path.join(c, fn)
                if os.path.exists(p):
                    v = os.path.join(c, "vocab.txt")
                    return p, (v if os.path.exists(v) else None)
        elif c.endswith((".safetensors", ".npz", ".bin", ".pt", ".pth")):
            v = os.path.join(os.path.dirname(c), "vocab.txt")
            return c, (v if os.path.exists(v) else None)
    return None


GREEDY = {"num_beams": 1, "no_repeat_ngram_size": 0, "length_penalty": 1.0, "early_stopping": False}


def _generation_config(weights_path: Optional[str]) -> Dict:
    """``generate()`` reads the checkpoint's generation settings: ``generation_config.json`` WHOLESALE when the file
    exists (transformers builds the GenerationConfig from it alone), else the legacy generation fields of
    ``config.json`` - the top-level ones and, attribute by attribute where the top level has none, those of its
    ``decoder`` sub-config (``GenerationConfig.from_model_config``, transformers/generation/configuration_utils.py;
    checked against transformers 5.5.0 itself by tests/test_host_logic.py.  4.x compared with 1 / 0 / 1.0 / False instead
    of None when it decided whether the top level had "set" a field).  The shipped kha-white/manga-ocr-base is believed
    to carry num_beams=4, no_repeat_ngram_size=3, length_penalty=2.0, early_stopping=true;
    without a file the path is greedy (BASELINE)."""
    gen = dict(GREEDY)
    for k, v in _generation_fields(weights_path, list(gen)).items():
        if v is not None:
            gen[k] = v
    return gen


def _generation_fields(weights_path: Optional[str], keys: Sequence[str]) -> Dict:
    """{key: value or None} as generate() would resolve them for the checkpoint of ``weights_path`` (see _generation_config)."""
    import json
    out = {k: None for k in keys}
    if not weights_path:
        return out
    d = weights_path if os.path.isdir(weights_path) else os.path.dirname(weights_path)
    for fn in ("generation_config.json", "config.json"):
        fp = os.path.join(d, fn)
        if not os.path.exists(fp):
            continue
        try:
            with open(fp, encoding="utf-8") as f:
                cfg = json.load(f)
        except (OSError, ValueError):
            continue
        if not isinstance(cfg, dict):
            continue
        sub = {}
        if fn == "config.json":
            sub = next((cfg[n] for n in ("decoder", "generator", "text_config") if isinstance(cfg.get(n), dict) and cfg[n]), {})
        for k in keys:
            if cfg.get(k) is not None:
                out[k] = cfg[k]
            elif sub.get(k) is not None:
                out[k] = sub[k]
        break                      # the first file that exists decides; they are not merged
    return out


# Generation settings that change what generate() returns and that this engine does not implement, with their neutral values.
# A checkpoint that carries one of them would silently decode differently here: the constructor refuses it instead
# (MOCR_IGNORE_GENERATION_EXTRAS=1 overrides).  Sampling-only knobs (temperature, top_k, top_p ...) matter only with do_sample.
_UNSUPPORTED_GENERATION = {
    "do_sample": (False,), "num_beam_groups": (1,), "diversity_penalty": (0, 0.0), "repetition_penalty": (1, 1.0),
    "encoder_repetition_penalty": (1, 1.0), "encoder_no_repeat_ngram_size": (0,), "min_length": (0,), "min_new_tokens": (0,),
    "max_new_tokens": (), "bad_words_ids": ([],), "force_words_ids": ([],), "suppress_tokens": ([],), "begin_suppress_tokens": ([],),
    "forced_bos_token_id": (), "forced_eos_token_id": (), "num_return_sequences": (1,), "penalty_alpha": (0, 0.0),
    "exponential_decay_length_penalty": (), "sequence_bias": ({}, []), "renormalize_logits": (False,), "constraints": ([],),
    # the token ids the kernels are built around (start = [CLS] 2, stop = [SEP] 3, fill = [PAD] 0): a checkpoint that names others
    "decoder_start_token_id": (2,), "eos_token_id": (3, [3]), "pad_token_id": (0,),
}


def _unsupported_generation_settings(weights_path: Optional[str]) -> Dict:
    """{key: value} of the checkpoint's generation settings this engine would not honour (empty: none)."""
    found = _generation_fields(weights_path, list(_UNSUPPORTED_GENERATION))
    return {k: v for k, v in found.items() if v is not None and v not in _UNSUPPORTED_GENERATION[k]}


# Hyper-parameters of config.json that change the arithmetic without changing a tensor shape (shapes are checked by
# weights.complete): what the kernels implement (model constants).  A key that is absent has this value
# by the library's default.
_ARCHITECTURE = {
    "encoder": {"model_type": ("vit",), "hidden_act": ("gelu",), "layer_norm_eps": (1e-12,), "num_attention_heads": (12,),
                "patch_size": (16,), "image_size": (224, [224, 224]), "qkv_bias": (True,), "num_channels": (3,)},
    "decoder": {"model_type": ("bert",), "hidden_act": ("gelu",), "layer_norm_eps": (1e-12,), "num_attention_heads": (12,),
                "position_embedding_type": ("absolute",), "is_decoder": (True,), "add_cross_attention": (True,)},
}


_PREPROCESSOR = {
    "do_resize": (True,), "do_rescale": (True,), "do_normalize": (True,), "resample": (2,), "rescale_factor": (1 / 255,),
    "size": (224, [224, 224], {"height": 224, "width": 224}), "image_mean": ([0.5, 0.5, 0.5], 0.5), "image_std": ([0.5, 0.5, 0.5], 0.5),
    "do_center_crop": (False,), "do_convert_rgb": (False,),
}


def _unsupported_architecture(weights_path: Optional[str]) -> Dict:
    """{"encoder.hidden_act": "gelu_new", ...}: fields of the checkpoint's config.json the kernels do not implement (empty: none,
    or no config.json next to the weights)."""
    import json
    if not weights_path:
        return {}
    d = weights_path if os.path.isdir(weights_path) else os.path.dirname(weights_path)
    try:
        with open(os.path.join(d, "config.json"), encoding="utf-8") as f:
            cfg = json.load(f)
    except (OSError, ValueError):
        return {}
    bad = {}
    for part, want in _ARCHITECTURE.items():
        sub = cfg.get(part) if isinstance(cfg, dict) else None
        if not isinstance(sub, dict):
            continue
        for k, ok in want.items():
            if sub.get(k) is not None and sub[k] not in ok:
                bad[f"{part}.{k}"] = sub[k]
    # preprocessor_config.json (what ViTImageProcessor.from_pretrained would read): the preprocess kernel is Pillow's bilinear
    # resize to 224 x 224, x / 255, (x - 0.5) / 0.5 and nothing else
    try:
        with open(os.path.join(d, "preprocessor_config.json"), encoding="utf-8") as f:
            pre = json.load(f)
    except (OSError, ValueError):
        pre = None
    if isinstance(pre, dict):
        for k, ok in _PREPROCESSOR.items():
            if pre.get(k) is not None and pre[k] not in ok:
                bad[f"preprocessor.{k}"] = pre[k]
    return bad


def image_to_array(img) -> np.ndarray:
    """PIL image -> the uint8 array the engine reads.  The luma conversion itself
    (``img.convert("L")``, the upstream wrapper's first step) happens on the GPU for the modes
    whose ``convert("L")`` is the ITU-R 601 map of their RGB bytes (RGB, and RGBA: Pillow ignores
    alpha there); every other mode (P, 1, LA, CMYK, I, I;16, F, YCbCr - whose L is its Y plane,
    not the luma of its RGB expansion - ...) is converted by Pillow itself, exactly as upstream."""
    if img.mode not in ("RGB", "L", "RGBA"):
        img = img.convert("L")
    a = np.asarray(img)
    if a.dtype != np.uint8:
        a = a.astype(np.uint8)
    return a


def _validated(a: np.ndarray) -> np.ndarray:
    if a.ndim not in (2, 3) or (a.ndim == 3 and a.shape[2] not in (1, 3, 4)) or a.shape[0] < 1 or a.shape[1] < 1:
        raise ValueError(f"unsupported image array shape {a.shape}")
    if max(a.shape[0], a.shape[1]) > 32768:
        raise ValueError(f"image of {a.shape[1]} x {a.shape[0]} px is larger than the engine's 32768 px limit")
    return a[:, :, 0] if a.ndim == 3 and a.shape[2] == 1 else a


class _Request:
    __slots__ = ("crop", "event", "text", "error")

    def __init__(self, crop: np.ndarray):
        self.crop = crop
        self.event = threading.Event()
        self.text: Optional[str] = None
        self.error: Optional[BaseException] = None


class MangaOcr:
    def __init__(self, pretrained_model_name_or_path: str = DEFAULT_MODEL, force_cpu: bool = False, *,
                 weights: Optional[Dict[str, np.ndarray]] = None, vocab: Optional[Vocab] = None,
                 devices: Optional[Sequence[int]] = None, max_batch: int = 64, max_length: int = MAX_LENGTH,
                 warmup: bool = True, num_beams: Optional[int] = None, no_repeat_ngram_size: Optional[int] = None,
                 length_penalty: Optional[float] = None, early_stopping=None, linger_ms: Optional[float] = None,
                 slots: Optional[int] = None, admission: Optional[bool] = None):
        if force_cpu:
            raise RuntimeError("manga_ocr_b200 has no CPU path (force_cpu=True is not supported); it needs a B200 GPU")
        gen = dict(GREEDY)
        if weights is None:
            name = pretrained_model_name_or_path
            env = os.environ.get("MOCR_WEIGHTS", "")
            if name.startswith("random") or env.startswith("random"):
                spec = (name if name.startswith("random") else env).split(":")
                seed = int(spec[1]) if len(spec) > 1 else 0
                eos_bias = float(spec[2]) if len(spec) > 2 else 0.0
                weights = W.random_init(seed, eos_bias=eos_bias)
            else:
                found = _find_checkpoint(name)
                if found is None:
                    raise FileNotFoundError(
                        f"no local checkpoint for {name!r}: pass a directory holding model.safetensors or pytorch_model.bin "
                        "(+ vocab.txt), set MOCR_WEIGHTS to one, put the snapshot into the Hugging Face cache "
                        "(nothing is downloaded), or use 'random[:seed[:eos_bias]]' for random-init weights")
                weights = W.load_weights(found[0])
                gen = _generation_config(found[0])
                arch = _unsupported_architecture(found[0])
                if arch:
                    raise NotImplementedError(f"the checkpoint's config.json asks for {arch}: the kernels implement ViT-base/16-224 (erf-GELU, "
                                              "LayerNorm eps 1e-12, qkv bias) + a 12-head BERT decoder with absolute positions only")
                extras = _unsupported_generation_settings(found[0])
                if extras and os.environ.get("MOCR_IGNORE_GENERATION_EXTRAS", "0") != "1":
                    raise NotImplementedError(
                        f"the checkpoint's generation settings {extras} are not implemented by this engine (it does greedy and beam "
                        "search with no_repeat_ngram_size / length_penalty / early_stopping); its output would differ from "
                        "generate()'s.  MOCR_IGNORE_GENERATION_EXTRAS=1 decodes without them")
                if vocab is None and found[1]:
                    vocab = Vocab.from_file(found[1])
        else:
            weights = W.complete(weights)
        for k, v in (("num_beams", num_beams), ("no_repeat_ngram_size", no_repeat_ngram_size), ("length_penalty", length_penalty),
                     ("early_stopping", early_stopping)):
            if v is not None:
                gen[k] = v
        if int(gen["num_beams"]) < 1 or int(gen["num_beams"]) > max_batch:
            raise ValueError(f"num_beams={gen['num_beams']} outside [1, max_batch={max_batch}]")
        self.generation = gen          # num_beams == 1 -> the greedy path; > 1 -> beam search
        self.vocab = vocab or Vocab.synthetic()
        self.max_length = max_length
        self.max_batch = max_batch
        devs = list(devices) if devices is not None else [int(os.environ.get("MOCR_DEVICE", os.environ.get("LOCAL_RANK", "0")))]
        if len(devs) == 1:
            self.engines = [Engine(weights, device=devs[0], max_batch=max_batch, max_length=max_length)]
        else:       # one engine per GPU, built concurrently (the bf16 conversion + upload of 111 M weights runs in the library, GIL released)
            built: Dict[int, Engine] = {}
            errs: List[BaseException] = []

            def build(i: int, d: int) -> None:
                try:
                    built[i] = Engine(weights, device=d, max_batch=max_batch, max_length=max_length)
                except BaseException as e:      # noqa: BLE001 - re-raised below
                    errs.append(e)

            ts = [threading.Thread(target=build, args=(i, d)) for i, d in enumerate(devs)]
            for t in ts:
                t.start()
            for t in ts:
                t.join()
            if errs:
                for e in built.values():
                    e.close()
                raise errs[0]
            self.engines = [built[i] for i in range(len(devs))]
        if slots is not None:      # greedy decode with fewer decoder rows than crops: in-flight slot refill (ragged real-text lengths)
            for e in self.engines:
                e.set_option("slots", int(slots))
        self._queue: deque = deque()
        self._cv = threading.Condition()
        self._closed = False
        self._busy = 0                     # dispatchers currently running a batch
        # __call__ traffic: admission into a running decode (greedy only; engine sessions, include/mocr_b200.h mocr_session_*): a
        # caller's crop joins the rows that are already stepping and is answered when ITS crop finishes.  admission=False (or
        # MOCR_ADMISSION=0) keeps the batch dispatcher: callers that arrive together share a batch and wait for all of it.
        if admission is None:
            admission = os.environ.get("MOCR_ADMISSION", "1") != "0"
        self.admission = bool(admission) and int(gen["num_beams"]) <= 1 and int(gen["no_repeat_ngram_size"]) <= 0
        self.session_rows = max(1, min(max_batch, int(slots) if slots else 64))
        self._engine_locks = [threading.Lock() for _ in self.engines]      # a session owns its engine; batch callers wait for it to drain
        self._batch_waiting = [0] * len(self.engines)
        self._session_prof = {} if os.environ.get("MOCR_SESSION_PROF") else None     # dispatcher phase times (seconds, count)
        self.linger_s = float(os.environ.get("MOCR_LINGER_MS", "1.5")) * 1e-3 if linger_ms is None else linger_ms * 1e-3
        ref = weakref.ref(self)
        if self.admission:
            self._threads = [threading.Thread(target=MangaOcr._dispatch_session, args=(ref, k, e, self._cv), name=f"mocr-gpu{e.device}", daemon=True)
                             for k, e in enumerate(self.engines)]
        else:
            self._threads = [threading.Thread(target=MangaOcr._dispatch, args=(ref, e, self._cv), name=f"mocr-gpu{e.device}", daemon=True)
                             for e in self.engines]
        for t in self._threads:
            t.start()
        if warmup:   # upstream runs one example image through the model inside __init__
            self(_example_image())

    # ---- reference API ------------------------------------------------------------
    def __call__(self, img_or_path) -> str:
        from PIL import Image
        if isinstance(img_or_path, (str, Path)):
            img = Image.open(img_or_path)
        elif isinstance(img_or_path, Image.Image):
            img = img_or_path
        else:
            raise ValueError(f"img_or_path must be a path or PIL.Image, instead got: {img_or_path}")
        req = _Request(_validated(image_to_array(img)))       # a malformed crop fails its own call, before it can join a batch
        with self._cv:
            if self._closed:
                raise RuntimeError("MangaOcr instance is closed")
            self._queue.append(req)
            self._cv.notify()
        req.event.wait()
        if req.error is not None:
            raise req.error
        return req.text  # type: ignore[return-value]

    # ---- batch API (for callers that hold a whole page of crops) ----
    def recognize_batch(self, crops: Sequence, order: int = RGB) -> List[str]:
        """crops: PIL images or uint8 arrays ([H,W], [H,W,3], [H,W,4]) -> list of strings."""
        arrays = [c if isinstance(c, np.ndarray) else image_to_array(c) for c in crops]
        ids = self.recognize_ids(arrays, order)
        return ids_to_texts(self.vocab, ids)

    def recognize_regions(self, page, regions: Sequence, order: int = RGB) -> List[str]:
        """All selections of ONE page -> strings.  ``page``: PIL image or uint8 array; ``regions``: ``Region``
        objects (``manga_ocr_b200.engine.Region``; ``Region.from_qt(rect, polygon, orientation)`` reproduces the
        reference's numbers).  The page is uploaded once; crop, polygon composite on white and rotation
        (reference/src/ui/main_window.py:6497-6506, 9789-9800) happen on the GPU.  Decoding follows the same
        generation settings as ``__call__`` / ``recognize_batch`` (greedy, or beam search when the checkpoint says so)."""
        arr = page if isinstance(page, np.ndarray) else image_to_array(page)
        regions = list(regions)

        def run(engine: Engine, regs: Sequence) -> np.ndarray:
            b = self._beam_args()
            if b is None:
                return engine.recognize_regions(arr, regs, order, self.max_length)[0]
            return engine.recognize_regions_beam(arr, regs, order, self.max_length, *b)[0]

        def work(k: int, lo: int, hi: int) -> np.ndarray:
            with self._engine_excl(k):
                return run(self.engines[k], regions[lo:hi])

        return ids_to_texts(self.vocab, self._sharded(len(regions), work))

    def _engine_ids(self, engine: Engine, arrays: Sequence[np.ndarray], order: int) -> np.ndarray:
        b = self._beam_args()
        if b is None:
            return engine.recognize(arrays, order, self.max_length)[0]
        return engine.recognize_beam(arrays, order, self.max_length, *b)[0]

    def _beam_args(self):
        """None for the plain greedy path, else (num_beams, no_repeat_ngram_size, length_penalty, early_stopping) of the search
        entry points.  generate() applies ``no_repeat_ngram_size`` in greedy mode too (NoRepeatNGramLogitsProcessor,
        generation/utils.py): the arg-max kernels have no ban list, but a ONE-beam search that stops at its first finished
        hypothesis follows the banned arg-max token by token and ends where greedy decoding ends - the same ids
        (tests/test_beam_host.py pins that against generate(num_beams=1, no_repeat_ngram_size=n))."""
        g = self.generation
        if int(g["num_beams"]) > 1:
            return int(g["num_beams"]), int(g["no_repeat_ngram_size"]), float(g["length_penalty"]), g["early_stopping"]
        if int(g["no_repeat_ngram_size"]) > 0:
            return 1, int(g["no_repeat_ngram_size"]), 1.0, True
        return None

    def recognize_ids(self, arrays: Sequence[np.ndarray], order: int = RGB) -> np.ndarray:
        def work(k: int, lo: int, hi: int) -> np.ndarray:
            with self._engine_excl(k):
                return self._engine_ids(self.engines[k], arrays[lo:hi], order)

        return self._sharded(len(arrays), work)

    @contextlib.contextmanager
    def _engine_excl(self, k: int):
        """Engine k for a batch call: a running admission session stops admitting, drains and hands the engine over."""
        with self._cv:
            self._batch_waiting[k] += 1
        try:
            with self._engine_locks[k]:
                yield
        finally:
            with self._cv:
                self._batch_waiting[k] -= 1
                self._cv.notify_all()

    def _sharded(self, n: int, work) -> np.ndarray:
        """Host-side job splitter: contiguous blocks of the n units, one worker thread per GPU, no collective."""
        if len(self.engines) == 1 or n <= 1:
            return work(0, 0, n)
        from .splitter import shard_bounds
        out = np.zeros((n, self.max_length), np.int32)
        errs: List[BaseException] = []

        def run(k: int) -> None:
            lo, hi = shard_bounds(n, len(self.engines), k)
            try:
                if hi > lo:
                    out[lo:hi] = work(k, lo, hi)
            except BaseException as e:   # noqa: BLE001 - re-raised on the caller's thread
                errs.append(e)

        ts = [threading.Thread(target=run, args=(k,)) for k in range(len(self.engines))]
        for t in ts:
            t.start()
        for t in ts:
            t.join()
        if errs:
            raise errs[0]
        return out

    # ---- cross-thread micro-batcher ------------------------------------------------
    def _take_batch(self) -> Optional[list]:
        """Called by a dispatcher with the condition held and requests queued (or the instance closed): linger a moment
        so that callers arriving together share a batch, then take this GPU's share of what is queued (with several
        GPUs nobody grabs the whole queue).  None = closed and drained."""
        if self.linger_s > 0 and 0 < len(self._queue) < self.max_batch and not self._closed:
            # A lone caller waits linger_s.  While requests keep arriving (a pool of workers resubmitting after the previous
            # batch: 50 Python threads need a few ms to come back) every arrival extends the wait by a third of linger_s,
            # up to 4 x linger_s in total - otherwise the burst is served as two half batches, each a full decode long.
            now = time.monotonic()
            deadline, cap, seen = now + self.linger_s, now + 4.0 * self.linger_s, len(self._queue)
            while len(self._queue) < self.max_batch and not self._closed:
                left = min(deadline, cap) - time.monotonic()
                if left <= 0:
                    break
                self._cv.wait(left)
                if len(self._queue) > seen:
                    seen = len(self._queue)
                    deadline = max(deadline, time.monotonic() + self.linger_s / 3.0)
        if not self._queue:
            return None
        share = -(-len(self._queue) // max(1, len(self.engines) - self._busy))     # ceil(queued / idle GPUs)
        take = min(len(self._queue), self.max_batch, max(1, share))
        self._busy += 1
        return [self._queue.popleft() for _ in range(take)]

    def _run_requests(self, engine: Engine, batch: list) -> None:
        try:
            ids = self._engine_ids(engine, [r.crop for r in batch], RGB)
            for r, t in zip(batch, ids_to_texts(self.vocab, ids)):
                r.text = t
        except BaseException as e:   # noqa: BLE001
            if len(batch) == 1:
                batch[0].error = e
            else:
                # one bad crop must not fail its neighbours (the reference isolates failures per call,
                # reference/src/core/workers.py:241-244): run the requests of a failed batch one by one
                for r in batch:
                    self._run_requests(engine, [r])
                return
        for r in batch:
            r.event.set()

    @staticmethod
    def _dispatch(ref, engine: Engine, cv: threading.Condition) -> None:
        # The thread holds the MangaOcr only through a weak reference while it is idle, so an instance that is dropped
        # without close() is collected (and its GPU memory freed) instead of being kept alive by its own threads.
        while True:
            with cv:
                self = ref()
                if self is None:
                    return
                if not self._queue and not self._closed:
                    del self
                    cv.wait(0.25)
                    continue
                batch = self._take_batch()
            if batch is None:
                return
            try:
                self._run_requests(engine, batch)
            finally:
                with cv:
                    self._busy -= 1
            del self, batch

    @staticmethod
    def _dispatch_session(ref, k: int, engine: Engine, cv: threading.Condition) -> None:
        """__call__ traffic of one GPU with admission: while requests are in flight the engine runs a session - queued
        requests are staged, encoded and published between two chunks of decode steps, finished crops are answered at once."""
        steps = int(os.environ.get("MOCR_SESSION_STEPS", "13"))   # decode steps between two admissions / result polls (13 = one CUDA graph)
        few = int(os.environ.get("MOCR_SESSION_FEW", "8"))        # at most this many crops in flight: answer latency before queue depth
        settle = float(os.environ.get("MOCR_SESSION_SETTLE_MS", "0.3")) * 1e-3    # wait this long for the callers just answered to call again
        while True:
            with cv:
                self = ref()
                if self is None:
                    return
                if not self._queue or self._batch_waiting[k] > 0:
                    if self._closed and not self._queue:
                        return
                    del self
                    cv.wait(0.25)
                    continue
                lock = self._engine_locks[k]
            del self
            with lock:
                inflight: Dict[int, _Request] = {}
                try:
                    self = ref()
                    if self is None:
                        return
                    try:
                        engine.session_begin(self.session_rows, RGB, self.max_length)
                    except BaseException:       # noqa: BLE001 - no session to be had (e.g. parity taps are set on the engine):
                        # the callers that are waiting are served as one batch instead, with its per-request failure isolation
                        with cv:
                            batch = [self._queue.popleft() for _ in range(min(len(self._queue), self.max_batch))]
                        if batch:
                            self._run_requests(engine, batch)
                        continue
                    capacity = self.max_batch
                    pending = 0                     # length snapshots enqueued and not yet read (at most two)
                    # a lightly loaded session steps only its first 16 rows (92 instead of 119 us per step); it grows to all of them
                    # once more crops are in flight, and stays there until it ends
                    small = (self.session_rows > 16 and os.environ.get("MOCR_SESSION_SMALL", "1") != "0"
                             and engine.session_rows(16) < self.session_rows)
                    prof = self._session_prof       # None, or {phase: [seconds, count]} (MOCR_SESSION_PROF=1: tools/call_latency.py)
                    clock = time.perf_counter

                    def lap(name, t0, n=1):
                        if prof is not None:
                            e = prof.setdefault(name, [0.0, 0])
                            e[0] += clock() - t0
                            e[1] += n
                        return clock()
                    try:
                        answered = 0                # callers answered in the previous round (they are about to call again)
                        while True:
                            t0 = clock()
                            # launch a chunk, take what is queued and admit it, launch the first chunk of a session that was idle, then
                            # read the length snapshot of the chunk BEFORE the one just launched: one chunk is always queued while the
                            # host works
                            launched = bool(inflight)
                            if launched:
                                if small and len(inflight) + len(self._queue) > 16:
                                    engine.session_rows(self.session_rows)
                                    small = False
                                engine.session_run(steps, wait=False)
                                pending += 1
                                t0 = lap("launch", t0)
                                if answered and settle > 0:
                                    # The callers that were answered a moment ago need a few hundred microseconds to come back with their
                                    # next crop; taking the queue right away would miss them by that much and cost them a whole chunk.
                                    # The chunk just launched keeps the GPU busy meanwhile.
                                    with cv:
                                        end = clock() + settle
                                        while len(self._queue) < answered and not self._closed:
                                            left = end - clock()
                                            if left <= 0:
                                                break
                                            cv.wait(left)
                                    t0 = lap("settle", t0)
                            answered = 0
                            with cv:
                                reqs = []
                                free = capacity - len(inflight)
                                if self._queue and free > 0 and self._batch_waiting[k] == 0:
                                    share = -(-len(self._queue) // len(self.engines))          # leave work for the other GPUs
                                    reqs = [self._queue.popleft() for _ in range(min(free, share, 64))]
                                if not reqs and not inflight:
                                    break
                            if small and len(inflight) + len(reqs) > 16:
                                engine.session_rows(self.session_rows)
                                small = False
                            t0 = lap("take", t0)
                            if reqs:
                                try:
                                    for r, s in zip(reqs, engine.session_add([r.crop for r in reqs])):
                                        inflight[int(s)] = r
                                except BaseException:       # noqa: BLE001 - one bad crop must not fail its neighbours
                                    for r in reqs:
                                        try:
                                            inflight[int(engine.session_add([r.crop])[0])] = r
                                        except BaseException as e:   # noqa: BLE001
                                            r.error = e
                                            r.event.set()
                                t0 = lap("admit", t0, len(reqs))
                            if inflight and not launched:
                                engine.session_run(steps, wait=False)
                                pending += 1
                                t0 = lap("launch", t0)
                            # (with only a few crops in flight the snapshot of the chunk just launched is read instead: a short text is
                            #  answered one chunk earlier, and an idle GPU between two chunks costs nothing then)
                            if pending >= 2 or (pending == 1 and len(inflight) <= few):
                                lens = engine.session_run(0)
                                pending -= 1
                                t0 = lap("snapshot", t0)
                                done = [s for s in inflight if lens[s] > 0]
                                if done:
                                    texts = ids_to_texts(self.vocab, engine.session_fetch(done, release=True))
                                    for s, t in zip(done, texts):
                                        r = inflight.pop(s)
                                        r.text = t
                                        r.event.set()
                                    answered = len(done)
                                    lap("answer", t0, len(done))
                    finally:
                        engine.session_end()
                except BaseException as e:          # noqa: BLE001 - the engine failed: every request in flight gets the error
                    for r in inflight.values():
                        r.error = e
                        r.event.set()
                    inflight.clear()
                finally:
                    self = None

    def close(self) -> None:
        with self._cv:
            if self._closed:
                return
            self._closed = True
            self._cv.notify_all()
        for t in self._threads:
            if t is not threading.current_thread():
                t.join()                 # queued requests are served first; the engines are idle afterwards
        for e in self.engines:
            e.close()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _example_image():
    from PIL import Image
    from .crops import single_224
    return Image.fromarray(single_224()[0])
