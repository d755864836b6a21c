import json
import struct

import numpy as np
import pytest

from manga_ocr_b200 import weights as W


def test_manifest_parameter_count():
    n = sum(int(np.prod(s)) for _, s, _ in W.manifest())
    # reference state_dict: 111 005 952 parameters incl. the dead ViT pooler (590 592) with the LM
    # head tied to the word embeddings; the manifest lists the head separately.
    assert n == 111_005_952 - 590_592 + W.VOCAB * W.D


def test_random_init_is_deterministic_and_tied():
    a, b = W.random_init(3), W.random_init(3)
    assert all(np.array_equal(a[k], b[k]) for k in a)
    assert a["decoder.cls.predictions.decoder.weight"] is a["decoder.bert.embeddings.word_embeddings.weight"]
    assert not a["decoder.bert.embeddings.word_embeddings.weight"][0].any()
    assert np.abs(a["encoder.encoder.layer.0.intermediate.dense.weight"]).max() <= 0.04
    u = W.random_init(3, untie_lm_head=True)
    assert u["decoder.cls.predictions.decoder.weight"] is not u["decoder.bert.embeddings.word_embeddings.weight"]
    e = W.random_init(3, eos_bias=2.0)
    assert e["decoder.cls.predictions.bias"][W.SEP_ID] == pytest.approx(a["decoder.cls.predictions.bias"][W.SEP_ID] + 2.0)


def test_safetensors_roundtrip(tmp_path):
    w = {"a.weight": np.arange(12, dtype=np.float32).reshape(3, 4), "b": np.float32([1.5, -2.0])}
    header, blobs, off = {}, [], 0
    for k, v in w.items():
        raw = v.tobytes()
        header[k] = {"dtype": "F32", "shape": list(v.shape), "data_offsets": [off, off + len(raw)]}
        blobs.append(raw)
        off += len(raw)
    bf = (np.float32([1.0, -3.5]).view(np.uint32) >> 16).astype("<u2").tobytes()
    header["c"] = {"dtype": "BF16", "shape": [2], "data_offsets": [off, off + len(bf)]}
    blobs.append(bf)
    hj = json.dumps(header).encode()
    p = tmp_path / "m.safetensors"
    p.write_bytes(struct.pack("<Q", len(hj)) + hj + b"".join(blobs))
    got = W.load_safetensors(str(p))
    assert np.array_equal(got["a.weight"], w["a.weight"]) and np.array_equal(got["b"], w["b"])
    assert np.array_equal(got["c"], np.float32([1.0, -3.5]))


def test_complete_validates():
    w = W.random_init(0)
    bad = dict(w)
    del bad["encoder.layernorm.weight"]
    with pytest.raises(KeyError):
        W.complete(bad)
    bad = dict(w)
    bad["encoder.layernorm.weight"] = np.zeros(5, np.float32)
    with pytest.raises(ValueError):
        W.complete(bad)
    untied = {k: v for k, v in w.items() if k != "decoder.cls.predictions.decoder.weight"}
    assert W.complete(untied)["decoder.cls.predictions.decoder.weight"] is w["decoder.bert.embeddings.word_embeddings.weight"]
