"""The oracle wrapper (oracle/reference_ocr.py over transformers' VisionEncoderDecoderModel)
against the committed golden vectors; the vectors were produced by the same libraries in the
build container (oracle/make_golden.py), so this pins weights, configs and wiring."""
import numpy as np

from oracle import make_golden as G


def test_encoder_and_logits_match_golden(oracle12, golden_model):
    crops = G.model_inputs()
    enc = oracle12.encoder_hidden(crops)
    assert np.abs(enc[:, G.ENC_ROWS] - golden_model["enc_rows"]).max() < 2e-4
    assert np.allclose(enc.std(axis=(1, 2)), golden_model["enc_std"], atol=1e-4)
    ids, logits = oracle12.generate_batch(crops, max_length=G.MODEL_T)
    assert np.array_equal(ids, golden_model["ids"])
    assert np.abs(logits[..., ::G.LOGIT_STRIDE] - golden_model["logits_strided"]).max() < 2e-4
    # teacher-forced single forward reproduces the step-by-step logits
    tf = oracle12.teacher_forced_logits(crops, ids)
    assert np.abs(tf - logits).max() < 1e-4
    assert ids.shape == (3, G.MODEL_T) and (ids[:, 0] == 2).all()


def test_strings_match_golden(oracle12, golden_text):
    from PIL import Image
    from oracle.reference_ocr import decode_ids, post_process
    crops = G.model_inputs()
    assert oracle12(Image.fromarray(crops[0])) == golden_text["texts"][0]
    for ids, text in zip(golden_text["ids"], golden_text["texts"]):
        assert post_process(decode_ids(oracle12.tokens, ids)) == text


def test_oracle_rejects_bad_argument(oracle12):
    import pytest
    with pytest.raises(ValueError):
        oracle12(12345)


def test_checkpoints_written_by_transformers_load_with_the_product_loaders(oracle12, weights0, tmp_path):
    """Row N4 as far as it can be checked offline: a checkpoint directory written by transformers itself (save_pretrained of the
    full-size VisionEncoderDecoderModel: config.json, generation_config.json, model.safetensors) and a 4.x-style pytorch_model.bin
    (torch.save of the state dict with the `position_ids` buffer and an UNTIED LM-head copy) go through the product's torch-free
    readers: every parameter the engine uses arrives bit-exactly, buffers and the dead pooler are tolerated, an untied head is kept."""
    import warnings
    import torch
    from manga_ocr_b200 import weights as W
    from manga_ocr_b200.ocr import GREEDY, _find_checkpoint, _generation_config, _unsupported_generation_settings
    d = tmp_path / "ckpt"
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        oracle12.model.save_pretrained(str(d))
    assert _find_checkpoint(str(d)) == (str(d / "model.safetensors"), None)
    got = W.complete(W.load_weights(str(d / "model.safetensors")))
    for k, v in weights0.items():
        assert np.array_equal(got[k], v), k
    assert _generation_config(str(d)) == GREEDY and _unsupported_generation_settings(str(d)) == {}
    # the config.json transformers wrote for this architecture is accepted; one that changes the arithmetic without changing a
    # shape (another activation, another LayerNorm epsilon, relative positions) is named
    import json
    from manga_ocr_b200.ocr import _unsupported_architecture
    assert _unsupported_architecture(str(d / "model.safetensors")) == {}
    cfg = json.loads((d / "config.json").read_text())
    cfg["encoder"]["hidden_act"] = "gelu_new"
    cfg["decoder"]["layer_norm_eps"] = 1e-5
    cfg["decoder"]["position_embedding_type"] = "relative_key"
    # ... and the preprocessor_config.json transformers' own ViT image processor (PIL backend, the oracle's) writes is accepted,
    # one with another resampling filter or other statistics is named
    from transformers.models.vit.image_processing_pil_vit import ViTImageProcessorPil
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ViTImageProcessorPil(image_mean=[0.5, 0.5, 0.5], image_std=[0.5, 0.5, 0.5]).save_pretrained(str(d))
    assert (d / "preprocessor_config.json").exists()
    assert _unsupported_architecture(str(d)) == {}, json.loads((d / "preprocessor_config.json").read_text())
    (tmp_path / "other").mkdir()
    pre = json.loads((d / "preprocessor_config.json").read_text())
    pre.update({"resample": 3, "image_mean": [0.485, 0.456, 0.406]})
    (tmp_path / "other" / "preprocessor_config.json").write_text(json.dumps(pre))
    (tmp_path / "other" / "config.json").write_text(json.dumps(cfg))
    assert _unsupported_architecture(str(tmp_path / "other")) == {"encoder.hidden_act": "gelu_new", "decoder.layer_norm_eps": 1e-5,
                                                                   "decoder.position_embedding_type": "relative_key",
                                                                   "preprocessor.resample": 3,
                                                                   "preprocessor.image_mean": [0.485, 0.456, 0.406]}
    sd = {k: v.clone() for k, v in oracle12.model.state_dict().items()}
    sd["decoder.bert.embeddings.position_ids"] = torch.arange(512)[None]
    head = sd["decoder.bert.embeddings.word_embeddings.weight"].clone() + 0.5
    sd["decoder.cls.predictions.decoder.weight"] = head
    torch.save(sd, str(tmp_path / "pytorch_model.bin"))
    got = W.complete(W.load_weights(str(tmp_path / "pytorch_model.bin")))
    for k, v in weights0.items():
        if k != "decoder.cls.predictions.decoder.weight":
            assert np.array_equal(got[k], v), k
    assert np.array_equal(got["decoder.cls.predictions.decoder.weight"], head.numpy())
