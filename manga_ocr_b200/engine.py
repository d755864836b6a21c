"""Python face of the C ABI (include/mocr_b200.h): one ``Engine`` = one handle on one GPU.

Crops are ``uint8`` numpy arrays ``[H, W]``, ``[H, W, 3]`` or ``[H, W, 4]`` exactly as the app
builds them before calling the reference engine (reference/src/ui/main_window.py:9800).
All arithmetic happens in the CUDA library; this file only marshals pointers.
"""
from __future__ import annotations

import ctypes
from ctypes import POINTER, byref, c_double, c_float, c_int32, c_int64, c_uint8, c_void_p
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import _lib
from .weights import D, ENC_TOKENS, IMAGE, VOCAB

RGB, BGR = 0, 1
ROT_NONE, ROT_CW, ROT_CCW = 0, 1, 2
TAP_PIXELS, TAP_ENCODER, TAP_LOGITS = 1, 2, 4
MAX_LENGTH = 300   # hard-coded by upstream MangaOcr.__call__


class MocrError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"mocr_b200 error {code}: {message}")
        self.code = code


def _as_crop_array(crops: Sequence[np.ndarray]) -> Tuple[ctypes.Array, List[np.ndarray]]:
    """Build the mocr_crop_t[] view of the arrays (kept alive by the returned list)."""
    keep: List[np.ndarray] = []
    arr = (_lib.mocr_crop_t * len(crops))()
    for i, c in enumerate(crops):
        if not isinstance(c, np.ndarray) or c.dtype != np.uint8 or c.ndim not in (2, 3):
            raise ValueError(f"crop {i}: expected a uint8 array [H,W], [H,W,3] or [H,W,4]")
        ch = 1 if c.ndim == 2 else c.shape[2]
        if ch not in (1, 3, 4) or c.shape[0] < 1 or c.shape[1] < 1:
            raise ValueError(f"crop {i}: unsupported shape {c.shape}")
        if c.strides[-1] != 1 or (c.ndim == 3 and c.strides[1] != ch) or c.strides[0] < c.shape[1] * ch:
            c = np.ascontiguousarray(c)
        keep.append(c)
        arr[i].data = c.ctypes.data
        arr[i].height = c.shape[0]
        arr[i].width = c.shape[1]
        arr[i].stride = c.strides[0]
        arr[i].channels = ch
    return arr, keep


class Region:
    """One selection on a page, as the reference stages it in front of the engine
    (reference/src/ui/main_window.py:6497-6506, 6429-6430, 9789-9795): ``box`` = the PIL crop box
    (left, top, right, bottom), ``polygon`` = optional [n, 2] int points in PAGE coordinates (outside it
    the crop is white), ``rotate`` = ROT_NONE / ROT_CW / ROT_CCW."""

    __slots__ = ("box", "polygon", "rotate")

    def __init__(self, box: Sequence[int], polygon: Optional[Sequence[Sequence[int]]] = None, rotate: int = ROT_NONE):
        self.box = tuple(int(v) for v in box)
        if len(self.box) != 4:
            raise ValueError("box must be (left, top, right, bottom)")
        self.polygon = None if polygon is None else np.ascontiguousarray(np.asarray(polygon).reshape(-1, 2), dtype=np.int32)
        self.rotate = int(rotate)

    @classmethod
    def from_qt(cls, rect_xywh: Sequence[int], polygon: Optional[Sequence[Sequence[int]]] = None, orientation: Optional[str] = None) -> "Region":
        """The reference's own numbers: the crop box is (x, y, QRect.right(), QRect.bottom()) with right = x + w - 1
        (so the crop is one pixel short of the rectangle, main_window.py:6497), and a crop is rotated when the
        text orientation disagrees with its aspect (:9789-9795)."""
        x, y, w, h = (int(v) for v in rect_xywh)
        box = (x, y, x + w - 1, y + h - 1)
        ch, cw = box[3] - box[1], box[2] - box[0]
        rot = ROT_NONE
        if orientation == "Vertical" and cw > ch:
            rot = ROT_CW
        elif orientation == "Horizontal" and ch > cw:
            rot = ROT_CCW
        return cls(box, polygon, rot)


def _as_region_array(regions: Sequence[Region]) -> Tuple[ctypes.Array, list]:
    keep = []
    arr = (_lib.mocr_region_t * len(regions))()
    for i, r in enumerate(regions):
        if not isinstance(r, Region):
            r = Region(*r)
        arr[i].left, arr[i].top, arr[i].right, arr[i].bottom = r.box
        arr[i].rotate = r.rotate
        if r.polygon is not None and len(r.polygon) > 0:
            keep.append(r.polygon)
            arr[i].polygon = r.polygon.ctypes.data_as(POINTER(c_int32))
            arr[i].n_points = len(r.polygon)
        else:
            arr[i].polygon = None
            arr[i].n_points = 0
    return arr, keep


class Engine:
    """B200 recognition engine: preprocess -> ViT encoder -> greedy BERT decoder."""

    def __init__(self, weights: Dict[str, np.ndarray], device: int = 0, max_batch: int = 64, max_length: int = MAX_LENGTH):
        self._lib = _lib.load()
        self._h = c_void_p()
        self.device = device
        self.max_batch = max_batch
        self.max_length = max_length
        self.n = 0
        self._cur_len = 0
        rc = self._lib.mocr_create(device, max_batch, max_length, byref(self._h))
        if rc != 0:
            msg = self._lib.mocr_last_error(None)
            self._h = c_void_p()
            raise MocrError(rc, msg.decode() if msg else "mocr_create failed")
        try:
            for name, arr in weights.items():
                a = np.ascontiguousarray(arr, dtype=np.float32)
                shape = (c_int64 * max(a.ndim, 1))(*(a.shape if a.ndim else (1,)))
                self._ck(self._lib.mocr_set_weight(self._h, name.encode(), a.ctypes.data_as(POINTER(c_float)), shape, max(a.ndim, 1)))
            self._ck(self._lib.mocr_finalize_weights(self._h))
        except Exception:
            self.close()
            raise

    # ---- plumbing
    def _ck(self, rc: int) -> None:
        if rc != 0:
            msg = self._lib.mocr_last_error(self._h)
            raise MocrError(rc, msg.decode() if msg else "unknown error")

    def close(self) -> None:
        h, self._h = self._h, c_void_p()
        if h:
            self._lib.mocr_destroy(h)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def stream(self) -> int:
        return int(self._lib.mocr_stream(self._h) or 0)

    def sync(self) -> None:
        self._ck(self._lib.mocr_sync(self._h))

    @property
    def launch_count(self) -> int:
        return int(self._lib.mocr_launch_count(self._h))

    @property
    def last_steps(self) -> int:
        return int(self._lib.mocr_last_steps(self._h))

    def set_option(self, key: str, value: int) -> None:
        self._ck(self._lib.mocr_set_option(self._h, key.encode(), int(value)))

    def set_taps(self, taps: int) -> None:
        self._ck(self._lib.mocr_set_taps(self._h, int(taps)))

    # ---- the path
    def recognize(self, crops: Sequence[np.ndarray], order: int = RGB, max_length: Optional[int] = None):
        """crops -> (ids [n, max_length] int32 PAD-filled, lens [n]); any n (chunked by max_batch)."""
        T = max_length or self.max_length
        n = len(crops)
        ids = np.zeros((n, T), np.int32)
        lens = np.zeros((n,), np.int32)
        if n == 0:
            return ids, lens
        arr, keep = _as_crop_array(crops)
        self._ck(self._lib.mocr_recognize(self._h, arr, n, order, T, ids.ctypes.data_as(POINTER(c_int32)),
                                          lens.ctypes.data_as(POINTER(c_int32))))
        del keep
        return ids, lens

    # ---- admission into a running decode (include/mocr_b200.h: mocr_session_*)
    def session_begin(self, rows: int, order: int = RGB, max_length: Optional[int] = None) -> None:
        self._sess_T = max_length or self.max_length
        self._ck(self._lib.mocr_session_begin(self._h, order, self._sess_T, rows))

    def session_add(self, crops: Sequence[np.ndarray]) -> np.ndarray:
        """Stage, encode and publish crops to the running session; returns the slot of each (n <= free slots)."""
        arr, keep = _as_crop_array(crops)
        slots = np.zeros((len(crops),), np.int32)
        self._ck(self._lib.mocr_session_add(self._h, arr, len(crops), slots.ctypes.data_as(POINTER(c_int32))))
        del keep
        return slots

    def session_run(self, steps: int, wait: bool = True) -> Optional[np.ndarray]:
        """`steps` greedy steps of every active row.  wait=True: returns lens [max_batch], > 0 where a slot's crop has finished;
        wait=False launches only (admit crops meanwhile, then session_run(0) waits and reads)."""
        if not wait:
            self._ck(self._lib.mocr_session_run(self._h, steps, None))
            return None
        lens = np.zeros((self.max_batch,), np.int32)
        self._ck(self._lib.mocr_session_run(self._h, steps, lens.ctypes.data_as(POINTER(c_int32))))
        return lens

    def session_fetch(self, slots: Sequence[int], release: bool = True) -> np.ndarray:
        """ids [n, max_length] of finished slots; release=True frees the slots for new crops."""
        sl = np.ascontiguousarray(slots, np.int32)
        ids = np.zeros((len(sl), self._sess_T), np.int32)
        if len(sl):
            self._ck(self._lib.mocr_session_fetch(self._h, sl.ctypes.data_as(POINTER(c_int32)), len(sl), ids.ctypes.data_as(POINTER(c_int32)),
                                                  1 if release else 0))
        return ids

    def session_rows(self, rows: int) -> int:
        """The following chunks step only the first `rows` decoder rows (rounded up to 16 or the session's count; returned).
        Growing is always allowed, shrinking only while no slot is in use."""
        rc = self._lib.mocr_session_rows(self._h, int(rows))
        if rc < 0:
            self._ck(rc)
        return int(rc)

    def session_end(self) -> None:
        self._ck(self._lib.mocr_session_end(self._h))

    def stage(self, crops: Sequence[np.ndarray], order: int = RGB) -> None:
        arr, keep = _as_crop_array(crops)
        self._ck(self._lib.mocr_stage_crops(self._h, arr, len(crops), order))
        self.n = len(crops)
        del keep

    def stage_regions(self, page: np.ndarray, regions: Sequence[Region], order: int = RGB) -> None:
        """One page + its selections (crop box, polygon, rotation resolved on the device)."""
        parr, keep = _as_crop_array([page])
        rarr, rkeep = _as_region_array(regions)
        self._ck(self._lib.mocr_stage_regions(self._h, parr, rarr, len(regions), order))
        self.n = len(regions)
        del keep, rkeep

    def recognize_regions(self, page: np.ndarray, regions: Sequence[Region], order: int = RGB, max_length: Optional[int] = None):
        T = max_length or self.max_length
        n = len(regions)
        ids = np.zeros((n, T), np.int32)
        lens = np.zeros((n,), np.int32)
        if n == 0:
            return ids, lens
        parr, keep = _as_crop_array([page])
        rarr, rkeep = _as_region_array(regions)
        self._ck(self._lib.mocr_recognize_regions(self._h, parr, rarr, n, order, T, ids.ctypes.data_as(POINTER(c_int32)),
                                                  lens.ctypes.data_as(POINTER(c_int32))))
        del keep, rkeep
        return ids, lens

    def region_mask(self, index: int, shape: Tuple[int, int]) -> np.ndarray:
        """Polygon mask of staged region ``index`` as the device rasterised it (tests)."""
        out = np.zeros(shape, np.uint8)
        self._ck(self._lib.mocr_get_region_mask(self._h, index, out.ctypes.data_as(POINTER(c_uint8))))
        return out

    def preprocess(self) -> None:
        self._ck(self._lib.mocr_preprocess(self._h))

    def encode(self) -> None:
        self._ck(self._lib.mocr_encode(self._h))

    def decode(self, max_length: Optional[int] = None, forced_ids: Optional[np.ndarray] = None) -> None:
        T = max_length or self.max_length
        p = None
        if forced_ids is not None:
            forced_ids = np.ascontiguousarray(forced_ids, dtype=np.int32)
            if forced_ids.shape != (self.n, T):
                raise ValueError(f"forced_ids must be [{self.n}, {T}]")
            p = forced_ids.ctypes.data_as(POINTER(c_int32))
        self._ck(self._lib.mocr_decode_greedy(self._h, T, p))
        self._cur_len = T

    def run_resident(self, max_length: Optional[int] = None) -> None:
        T = max_length or self.max_length
        self._ck(self._lib.mocr_run_resident(self._h, T))
        self._cur_len = T

    def decode_beam(self, num_beams: int = 4, max_length: Optional[int] = None, no_repeat_ngram_size: int = 3,
                    length_penalty: float = 2.0, early_stopping=True):
        """Beam search over the encoded crops (n * num_beams <= max_batch); the defaults are the generation config the
        shipped checkpoint is believed to carry.  early_stopping: False / True / "never".
        Returns (ids [n, T] best hypothesis, EOS-filled past its end as the reference does; lens [n]; scores [n])."""
        T = max_length or self.max_length
        ids = np.zeros((self.n, T), np.int32)
        lens = np.zeros((self.n,), np.int32)
        scores = np.zeros((self.n,), np.float32)
        early = 2 if early_stopping == "never" else (1 if early_stopping else 0)
        self._ck(self._lib.mocr_decode_beam(self._h, num_beams, T, no_repeat_ngram_size, length_penalty, early,
                                            ids.ctypes.data_as(POINTER(c_int32)), lens.ctypes.data_as(POINTER(c_int32)),
                                            scores.ctypes.data_as(POINTER(c_float))))
        return ids, lens, scores

    def recognize_beam(self, crops: Sequence[np.ndarray], order: int = RGB, max_length: Optional[int] = None, num_beams: int = 4,
                       no_repeat_ngram_size: int = 3, length_penalty: float = 2.0, early_stopping=True):
        """crops -> (ids, lens, scores) with beam search, any n; ONE call into the library (the handle stays locked from
        staging to the result, so threads sharing this engine cannot interleave between the stages)."""
        T = max_length or self.max_length
        n = len(crops)
        ids = np.zeros((n, T), np.int32)
        lens = np.zeros((n,), np.int32)
        scores = np.zeros((n,), np.float32)
        if n == 0:
            return ids, lens, scores
        arr, keep = _as_crop_array(crops)
        early = 2 if early_stopping == "never" else (1 if early_stopping else 0)
        self._ck(self._lib.mocr_recognize_beam(self._h, arr, n, order, T, num_beams, no_repeat_ngram_size, length_penalty, early,
                                               ids.ctypes.data_as(POINTER(c_int32)), lens.ctypes.data_as(POINTER(c_int32)),
                                               scores.ctypes.data_as(POINTER(c_float))))
        del keep
        return ids, lens, scores

    def recognize_regions_beam(self, page: np.ndarray, regions: Sequence[Region], order: int = RGB, max_length: Optional[int] = None,
                               num_beams: int = 4, no_repeat_ngram_size: int = 3, length_penalty: float = 2.0, early_stopping=True):
        """The selections of one page -> (ids, lens, scores) with beam search (one library call, like recognize_beam)."""
        T = max_length or self.max_length
        n = len(regions)
        ids = np.zeros((n, T), np.int32)
        lens = np.zeros((n,), np.int32)
        scores = np.zeros((n,), np.float32)
        if n == 0:
            return ids, lens, scores
        parr, keep = _as_crop_array([page])
        rarr, rkeep = _as_region_array(regions)
        early = 2 if early_stopping == "never" else (1 if early_stopping else 0)
        self._ck(self._lib.mocr_recognize_regions_beam(self._h, parr, rarr, n, order, T, num_beams, no_repeat_ngram_size, length_penalty,
                                                       early, ids.ctypes.data_as(POINTER(c_int32)), lens.ctypes.data_as(POINTER(c_int32)),
                                                       scores.ctypes.data_as(POINTER(c_float))))
        del keep, rkeep
        return ids, lens, scores

    def fetch_ids(self):
        ids = np.zeros((self.n, self._cur_len), np.int32)
        lens = np.zeros((self.n,), np.int32)
        self._ck(self._lib.mocr_fetch_ids(self._h, ids.ctypes.data_as(POINTER(c_int32)), lens.ctypes.data_as(POINTER(c_int32))))
        return ids, lens

    # ---- parity taps
    def pixels_u8(self) -> np.ndarray:
        out = np.empty((self.n, IMAGE, IMAGE), np.uint8)
        self._ck(self._lib.mocr_get_pixels_u8(self._h, out.ctypes.data_as(POINTER(c_uint8))))
        return out

    def pixel_values(self) -> np.ndarray:
        out = np.empty((self.n, IMAGE, IMAGE), np.float32)
        self._ck(self._lib.mocr_get_pixel_values(self._h, out.ctypes.data_as(POINTER(c_float))))
        return out

    def encoder_hidden(self) -> np.ndarray:
        out = np.empty((self.n, ENC_TOKENS, D), np.float32)
        self._ck(self._lib.mocr_get_encoder_hidden(self._h, out.ctypes.data_as(POINTER(c_float))))
        return out

    def step_logits(self) -> np.ndarray:
        out = np.empty((self.n, self._cur_len - 1, VOCAB), np.float32)
        self._ck(self._lib.mocr_get_step_logits(self._h, out.ctypes.data_as(POINTER(c_float))))
        return out

    def time_kernel(self, name: str, iters: int = 20):
        """(ms per launch, algorithmic bytes, algorithmic flops) of one named kernel, CUDA-event timed."""
        ms, by, fl = c_float(), c_double(), c_double()
        self._ck(self._lib.mocr_time_kernel(self._h, name.encode(), iters, byref(ms), byref(by), byref(fl)))
        return ms.value, by.value, fl.value

    # ---- kernel-level unit hooks (tests)
    def test_gemm(self, epi: int, bn: int, A: np.ndarray, Wt: np.ndarray, bias: np.ndarray, resid: Optional[np.ndarray] = None):
        """out = epilogue(A @ Wt.T + bias) through the tcgen05 kernel; returns (out f32 [M,N], argmax [M])."""
        A = np.ascontiguousarray(A, np.float32)
        Wt = np.ascontiguousarray(Wt, np.float32)
        bias = np.ascontiguousarray(bias, np.float32)
        M, K = A.shape
        N = Wt.shape[0]
        out = np.zeros((M, N), np.float32)
        am = np.zeros((M,), np.int32)
        fp = lambda a: a.ctypes.data_as(POINTER(c_float))
        r = None if resid is None else np.ascontiguousarray(resid, np.float32)
        self._ck(self._lib.mocr_test_gemm(self._h, epi, bn, M, N, K, fp(A), fp(Wt), fp(bias), None if r is None else fp(r), fp(out),
                                          am.ctypes.data_as(POINTER(c_int32))))
        return out, am

    def test_encoder_attention(self, qkv: np.ndarray) -> np.ndarray:
        qkv = np.ascontiguousarray(qkv, np.float32)
        n = qkv.shape[0] // ENC_TOKENS
        out = np.zeros((n * ENC_TOKENS, D), np.float32)
        self._ck(self._lib.mocr_test_encoder_attention(self._h, n, qkv.ctypes.data_as(POINTER(c_float)), out.ctypes.data_as(POINTER(c_float))))
        return out

    def test_decode_attention(self, mode: int, q: np.ndarray, k: np.ndarray, v: np.ndarray, pos: Optional[np.ndarray] = None,
                              new_k: Optional[np.ndarray] = None, new_v: Optional[np.ndarray] = None):
        """One decode-step attention stage on caller data (include/mocr_b200.h: mocr_test_decode_attention).
        Returns ctx [n, 768] (and, for self-attention, the cache rows the kernel appended: (ctx, k_row, v_row))."""
        f = lambda a: None if a is None else np.ascontiguousarray(a, np.float32)
        q, k, v, new_k, new_v = f(q), f(k), f(v), f(new_k), f(new_v)
        n, n_ctx = q.shape[0], k.shape[1]
        fp = lambda a: None if a is None else a.ctypes.data_as(POINTER(c_float))
        ctx = np.zeros((n, D), np.float32)
        kr = np.zeros((n, D), np.float32)
        vr = np.zeros((n, D), np.float32)
        pp = None if pos is None else np.ascontiguousarray(pos, np.int32)
        self._ck(self._lib.mocr_test_decode_attention(self._h, mode, n, n_ctx, None if pp is None else pp.ctypes.data_as(POINTER(c_int32)),
                                                      fp(q), fp(k), fp(v), fp(new_k), fp(new_v), fp(ctx), fp(kr), fp(vr)))
        return (ctx, kr, vr) if mode in (1, 3) else ctx

    def test_stage_gemm(self, kind: int, A: np.ndarray, Wt: np.ndarray, bias: np.ndarray, resid: Optional[np.ndarray] = None,
                        gamma: Optional[np.ndarray] = None, beta: Optional[np.ndarray] = None, gelu: bool = False):
        """One small-M decoder GEMM stage on caller data (mocr_test_stage_gemm); returns (out [n, N], argmax [n])."""
        f = lambda a: None if a is None else np.ascontiguousarray(a, np.float32)
        A, Wt, bias, resid, gamma, beta = f(A), f(Wt), f(bias), f(resid), f(gamma), f(beta)
        n, K = A.shape
        N = Wt.shape[0]
        fp = lambda a: None if a is None else a.ctypes.data_as(POINTER(c_float))
        out = np.zeros((n, N), np.float32)
        am = np.zeros((n,), np.int32)
        self._ck(self._lib.mocr_test_stage_gemm(self._h, kind, n, N, K, fp(A), fp(Wt), fp(bias), fp(resid), fp(gamma), fp(beta), int(gelu), fp(out),
                                                am.ctypes.data_as(POINTER(c_int32))))
        return out, am

    def decode_profile(self, n: int = 4096) -> np.ndarray:
        out = np.zeros((n,), np.int64)
        self._ck(self._lib.mocr_get_decode_profile(self._h, out.ctypes.data_as(POINTER(c_int64)), n))
        return out


class BeamScorer:
    """The host bookkeeping of the beam search alone (mocr_beam_*; no device needed).  Rows = n * num_beams,
    K = 2 * num_beams candidates per row and step."""

    def __init__(self, n: int, num_beams: int = 4, max_length: int = MAX_LENGTH, no_repeat_ngram_size: int = 3,
                 length_penalty: float = 2.0, early_stopping=True):
        self._lib = _lib.load()
        self._b = c_void_p()
        self.n, self.beams, self.K, self.T = n, num_beams, 2 * num_beams, max_length
        early = 2 if early_stopping == "never" else (1 if early_stopping else 0)
        rc = self._lib.mocr_beam_create(n, num_beams, max_length, no_repeat_ngram_size, length_penalty, early, byref(self._b))
        if rc != 0:
            raise MocrError(rc, "mocr_beam_create: bad argument")

    def banned(self, row: int) -> List[int]:
        buf = np.zeros((self.T,), np.int32)
        cnt = self._lib.mocr_beam_banned(self._b, row, buf.ctypes.data_as(POINTER(c_int32)), self.T)
        if cnt < 0:
            raise MocrError(cnt, "mocr_beam_banned")
        return buf[:cnt].tolist()

    def step(self, cand_logprob: np.ndarray, cand_token: np.ndarray):
        """-> (unfinished, next_tokens [rows], parents [rows])"""
        lp = np.ascontiguousarray(cand_logprob, np.float32)
        tk = np.ascontiguousarray(cand_token, np.int32)
        rows = self.n * self.beams
        if lp.shape != (rows, self.K) or tk.shape != (rows, self.K):
            raise ValueError(f"candidates must be [{rows}, {self.K}]")
        nxt = np.zeros((rows,), np.int32)
        par = np.zeros((rows,), np.int32)
        rc = self._lib.mocr_beam_step(self._b, lp.ctypes.data_as(POINTER(c_float)), tk.ctypes.data_as(POINTER(c_int32)),
                                      nxt.ctypes.data_as(POINTER(c_int32)), par.ctypes.data_as(POINTER(c_int32)))
        if rc < 0:
            raise MocrError(rc, "mocr_beam_step")
        return rc == 1, nxt, par

    def result(self):
        ids = np.zeros((self.n, self.T), np.int32)
        lens = np.zeros((self.n,), np.int32)
        scores = np.zeros((self.n,), np.float32)
        self._lib.mocr_beam_result(self._b, ids.ctypes.data_as(POINTER(c_int32)), lens.ctypes.data_as(POINTER(c_int32)),
                                   scores.ctypes.data_as(POINTER(c_float)))
        return ids, lens, scores

    def close(self) -> None:
        if self._b:
            self._lib.mocr_beam_destroy(self._b)
            self._b = c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:      # noqa: BLE001 - interpreter shutdown
            pass
