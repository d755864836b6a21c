"""Weight manifest, deterministic random-init and file loading for the
manga-ocr-base architecture (ViT-base/16-224 encoder + 2-layer BERT decoder).

Tensor names and shapes are exactly those of the reference model's
``VisionEncoderDecoderModel.state_dict()`` (SURVEY.md Appendix A;
transformers/models/vit/modeling_vit.py:100-167,199-346 and
transformers/models/bert/modeling_bert.py:53-112,143-356,471-501), so a real
``kha-white/manga-ocr-base`` checkpoint (safetensors) loads through the same
path as the synthetic weights used by the tests and the benchmark.

The random generator is numpy PCG64 (bit-reproducible on every machine), not
torch, so golden vectors generated in one container stay valid on the GPU box.
"""
from __future__ import annotations

import json
import struct
from typing import Dict, Iterator, Tuple

import numpy as np

D = 768          # hidden size (encoder and decoder)
FFN = 3072
HEADS = 12
HEAD_DIM = 64
ENC_LAYERS = 12
DEC_LAYERS = 2
VOCAB = 6144
ENC_TOKENS = 197  # 196 patches + CLS
MAX_POS = 512
PATCH = 16
IMAGE = 224

PAD_ID, UNK_ID, CLS_ID, SEP_ID, MASK_ID = 0, 1, 2, 3, 4


def manifest() -> Iterator[Tuple[str, Tuple[int, ...], str]]:
    """Yield (name, shape, kind) for every tensor of the architecture.

    kind is one of "w" (matrix / embedding table), "b" (bias), "g" (LayerNorm
    gamma), "beta" (LayerNorm beta).
    """
    e = "encoder."
    yield e + "embeddings.cls_token", (1, 1, D), "w"
    yield e + "embeddings.position_embeddings", (1, ENC_TOKENS, D), "w"
    yield e + "embeddings.patch_embeddings.projection.weight", (D, 3, PATCH, PATCH), "w"
    yield e + "embeddings.patch_embeddings.projection.bias", (D,), "b"
    for i in range(ENC_LAYERS):
        p = f"{e}encoder.layer.{i}."
        for n in ("query", "key", "value"):
            yield p + f"attention.attention.{n}.weight", (D, D), "w"
            yield p + f"attention.attention.{n}.bias", (D,), "b"
        yield p + "attention.output.dense.weight", (D, D), "w"
        yield p + "attention.output.dense.bias", (D,), "b"
        yield p + "intermediate.dense.weight", (FFN, D), "w"
        yield p + "intermediate.dense.bias", (FFN,), "b"
        yield p + "output.dense.weight", (D, FFN), "w"
        yield p + "output.dense.bias", (D,), "b"
        yield p + "layernorm_before.weight", (D,), "g"
        yield p + "layernorm_before.bias", (D,), "beta"
        yield p + "layernorm_after.weight", (D,), "g"
        yield p + "layernorm_after.bias", (D,), "beta"
    yield e + "layernorm.weight", (D,), "g"
    yield e + "layernorm.bias", (D,), "beta"
    # encoder.pooler.* exists in the reference state_dict but its output is
    # never read on this path (modeling_vision_encoder_decoder.py:397) - it is
    # not part of this manifest.
    d = "decoder.bert."
    yield d + "embeddings.word_embeddings.weight", (VOCAB, D), "w"
    yield d + "embeddings.position_embeddings.weight", (MAX_POS, D), "w"
    yield d + "embeddings.token_type_embeddings.weight", (2, D), "w"
    yield d + "embeddings.LayerNorm.weight", (D,), "g"
    yield d + "embeddings.LayerNorm.bias", (D,), "beta"
    for i in range(DEC_LAYERS):
        p = f"{d}encoder.layer.{i}."
        for blk in ("attention", "crossattention"):
            for n in ("query", "key", "value"):
                yield p + f"{blk}.self.{n}.weight", (D, D), "w"
                yield p + f"{blk}.self.{n}.bias", (D,), "b"
            yield p + f"{blk}.output.dense.weight", (D, D), "w"
            yield p + f"{blk}.output.dense.bias", (D,), "b"
            yield p + f"{blk}.output.LayerNorm.weight", (D,), "g"
            yield p + f"{blk}.output.LayerNorm.bias", (D,), "beta"
        yield p + "intermediate.dense.weight", (FFN, D), "w"
        yield p + "intermediate.dense.bias", (FFN,), "b"
        yield p + "output.dense.weight", (D, FFN), "w"
        yield p + "output.dense.bias", (D,), "b"
        yield p + "output.LayerNorm.weight", (D,), "g"
        yield p + "output.LayerNorm.bias", (D,), "beta"
    c = "decoder.cls.predictions."
    yield c + "transform.dense.weight", (D, D), "w"
    yield c + "transform.dense.bias", (D,), "b"
    yield c + "transform.LayerNorm.weight", (D,), "g"
    yield c + "transform.LayerNorm.bias", (D,), "beta"
    yield c + "decoder.weight", (VOCAB, D), "w"   # tied to word_embeddings unless untied
    yield c + "bias", (VOCAB,), "b"


def random_init(seed: int = 0, *, plain_hf_init: bool = False, untie_lm_head: bool = False,
                eos_bias: float = 0.0, gain: float = 1.0) -> Dict[str, np.ndarray]:
    """Random weights of the manga-ocr-base architecture as {name: float32 array}.

    Matrices follow the reference's initialiser (trunc-normal, sigma 0.02, cut
    at +-2 sigma; modeling_vit.py:385-398).  With ``plain_hf_init`` biases are 0
    and LayerNorm is identity exactly as a fresh reference model; the default
    instead draws small non-zero biases and LayerNorm affine terms so that every
    bias / gamma / beta path of the kernels is exercised by the parity tests.

    gain scales sigma of every matrix: at the reference's sigma = 0.02 the decoder's output is
    almost independent of the image (the cross-attention contribution is tiny next to the
    residual stream), so tests that must tell crops apart use gain > 1.

    eos_bias is added to ``cls.predictions.bias[SEP_ID]`` so that greedy decode
    terminates at varied lengths (with a plain init EOS essentially never wins
    and every sequence runs the full max_length).
    """
    rng = np.random.default_rng(seed)
    out: Dict[str, np.ndarray] = {}
    for name, shape, kind in manifest():
        if kind == "w":
            a = rng.standard_normal(shape, dtype=np.float32) * np.float32(0.02 * gain)
            np.clip(a, -0.04 * gain, 0.04 * gain, out=a)
        elif kind == "b":
            a = (np.zeros(shape, np.float32) if plain_hf_init
                 else rng.standard_normal(shape, dtype=np.float32) * np.float32(0.02))
        elif kind == "g":
            a = (np.ones(shape, np.float32) if plain_hf_init
                 else (1.0 + 0.05 * rng.standard_normal(shape, dtype=np.float32)).astype(np.float32))
        else:
            a = (np.zeros(shape, np.float32) if plain_hf_init
                 else rng.standard_normal(shape, dtype=np.float32) * np.float32(0.02))
        out[name] = np.ascontiguousarray(a, dtype=np.float32)
    wemb = out["decoder.bert.embeddings.word_embeddings.weight"]
    wemb[PAD_ID] = 0.0   # nn.Embedding(padding_idx=0) zeroes this row (modeling_bert.py:58)
    if not untie_lm_head:
        out["decoder.cls.predictions.decoder.weight"] = wemb
    if eos_bias:
        out["decoder.cls.predictions.bias"][SEP_ID] += np.float32(eos_bias)
    return out


_ST_DTYPES = {"F32": np.float32, "F16": np.float16, "F64": np.float64}


def load_safetensors(path: str) -> Dict[str, np.ndarray]:
    """Minimal safetensors reader (8-byte LE header length, JSON header, raw
    little-endian data).  bf16 tensors are widened to float32."""
    out: Dict[str, np.ndarray] = {}
    with open(path, "rb") as f:
        (hlen,) = struct.unpack("<Q", f.read(8))
        header = json.loads(f.read(hlen))
        base = 8 + hlen
        for name, meta in header.items():
            if name == "__metadata__":
                continue
            lo, hi = meta["data_offsets"]
            f.seek(base + lo)
            raw = f.read(hi - lo)
            dt = meta["dtype"]
            if dt == "BF16":
                u16 = np.frombuffer(raw, dtype="<u2").astype(np.uint32) << 16
                arr = u16.view(np.float32)
            elif dt in _ST_DTYPES:
                arr = np.frombuffer(raw, dtype=np.dtype(_ST_DTYPES[dt]).newbyteorder("<")).astype(np.float32)
            else:
                continue  # integer buffers (position_ids ...) are not weights
            out[name] = np.ascontiguousarray(arr.reshape(meta["shape"]), dtype=np.float32)
    return out


_PT_STORAGE = {"FloatStorage": np.float32, "HalfStorage": np.float16, "DoubleStorage": np.float64, "BFloat16Storage": "bf16",
               "LongStorage": np.int64, "IntStorage": np.int32, "ShortStorage": np.int16, "CharStorage": np.int8, "ByteStorage": np.uint8,
               "BoolStorage": np.bool_}


def load_torch_bin(path: str) -> Dict[str, np.ndarray]:
    """``pytorch_model.bin`` (the zip container ``torch.save`` writes: ``<root>/data.pkl`` + one raw little-endian file
    per storage) read WITHOUT torch: a restricted unpickler that only knows how to rebuild tensors, so a hub
    snapshot that ships no ``model.safetensors`` still loads where the reference's ``from_pretrained`` works.
    Floating tensors come back as float32; integer buffers (``position_ids``) are dropped."""
    import pickle
    import zipfile
    from collections import OrderedDict

    with zipfile.ZipFile(path) as z:
        names = z.namelist()
        pkl = next((n for n in names if n.endswith("data.pkl")), None)
        if pkl is None:
            raise ValueError(f"{path}: not a torch zip checkpoint (no data.pkl); legacy (pre-1.6) files are not supported")
        root = pkl[: -len("data.pkl")]

        class _Storage:
            def __init__(self, kind):
                self.kind = kind

        def _rebuild_tensor_v2(storage, offset, size, stride, *unused):
            dt, raw = storage
            if dt == "bf16":
                arr = (np.frombuffer(raw, dtype="<u2").astype(np.uint32) << 16).view(np.float32)
            else:
                arr = np.frombuffer(raw, dtype=np.dtype(dt).newbyteorder("<"))
            size, stride = tuple(size), tuple(stride)
            if len(size) == 0:
                return arr[offset:offset + 1].reshape(())
            view = np.lib.stride_tricks.as_strided(arr[offset:], shape=size, strides=tuple(st * arr.itemsize for st in stride))
            return np.ascontiguousarray(view)

        class _Unpickler(pickle.Unpickler):
            def find_class(self, module, name):
                if module == "collections" and name == "OrderedDict":
                    return OrderedDict
                if module == "torch._utils" and name in ("_rebuild_tensor_v2", "_rebuild_tensor"):
                    return _rebuild_tensor_v2
                if module == "torch._utils" and name == "_rebuild_parameter":
                    return lambda data, requires_grad, hooks: data
                if module in ("torch", "torch.storage") and name in _PT_STORAGE:
                    return _Storage(_PT_STORAGE[name])
                raise pickle.UnpicklingError(f"{path}: refusing to unpickle {module}.{name}")

            def persistent_load(self, pid):
                # ('storage', storage_type, key, location, numel)
                if not isinstance(pid, tuple) or pid[0] != "storage" or not isinstance(pid[1], _Storage):
                    raise pickle.UnpicklingError(f"{path}: unexpected persistent id {pid!r}")
                return pid[1].kind, z.read(f"{root}data/{pid[2]}")

        with z.open(pkl) as f:
            sd = _Unpickler(f).load()
    if isinstance(sd, dict) and "state_dict" in sd and isinstance(sd["state_dict"], dict):
        sd = sd["state_dict"]
    out: Dict[str, np.ndarray] = {}
    for k, v in sd.items():
        if isinstance(v, np.ndarray) and v.dtype.kind == "f":
            out[k] = np.ascontiguousarray(v, dtype=np.float32)
    return out


def load_weights(path: str) -> Dict[str, np.ndarray]:
    """Load a checkpoint: ``*.safetensors``, ``*.bin`` / ``*.pt`` (torch zip) or ``*.npz``, keyed by reference names."""
    if path.endswith(".npz"):
        with np.load(path) as z:
            w = {k: np.ascontiguousarray(z[k], dtype=np.float32) for k in z.files}
    elif path.endswith((".bin", ".pt", ".pth")):
        w = load_torch_bin(path)
    else:
        w = load_safetensors(path)
    return complete(w)


def complete(w: Dict[str, np.ndarray]) -> Dict[str, np.ndarray]:
    """Resolve the tied tensors a checkpoint may omit and validate shapes."""
    w = dict(w)
    c = "decoder.cls.predictions."
    if c + "decoder.weight" not in w:
        w[c + "decoder.weight"] = w["decoder.bert.embeddings.word_embeddings.weight"]
    if c + "bias" not in w and c + "decoder.bias" in w:
        w[c + "bias"] = w[c + "decoder.bias"]
    for name, shape, _ in manifest():
        if name not in w:
            raise KeyError(f"checkpoint is missing tensor {name!r}")
        if tuple(w[name].shape) != shape:
            raise ValueError(f"tensor {name!r} has shape {tuple(w[name].shape)}, expected {shape}")
    return w
