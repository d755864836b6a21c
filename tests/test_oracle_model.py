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
