import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session", autouse=True)
def _library_present():
    """The C-ABI library is built in-tree by __graft_entry__.build(); build it here only when it is missing."""
    from manga_ocr_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        _lib.build()


@pytest.fixture(scope="session")
def golden_pre():
    return np.load(os.path.join(GOLDEN, "preprocess_kat.npz"))


@pytest.fixture(scope="session")
def golden_model():
    return np.load(os.path.join(GOLDEN, "model_kat.npz"))


@pytest.fixture(scope="session")
def golden_text():
    import json
    with open(os.path.join(GOLDEN, "text_kat.json"), encoding="utf-8") as f:
        return json.load(f)


@pytest.fixture(scope="session")
def weights0():
    from manga_ocr_b200 import weights as W
    return W.random_init(0)


@pytest.fixture(scope="session")
def oracle12(weights0):
    """The CPU oracle (transformers VisionEncoderDecoderModel) with the seed-0 weights."""
    from manga_ocr_b200.text import Vocab
    from oracle.reference_ocr import ReferenceMangaOcr
    return ReferenceMangaOcr(weights0, Vocab.synthetic().tokens, max_length=12)


def _have_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.fixture(scope="session")
def engine8(weights0):
    """One small engine shared by the GPU parity tests (8 crops x 24 tokens)."""
    if not _have_gpu():
        pytest.skip("no GPU")
    from manga_ocr_b200.engine import Engine
    e = Engine(weights0, device=0, max_batch=8, max_length=24)
    yield e
    e.close()
