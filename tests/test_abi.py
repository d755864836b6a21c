"""The C-ABI shared library: it builds in-tree, loads without a GPU, exports every symbol that
include/mocr_b200.h declares, and refuses loudly to compute when there is no B200 (there is no
CPU fallback).  No compute call is made here."""
import ctypes
import os
import re

import numpy as np
import pytest

from manga_ocr_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions():
    src = open(os.path.join(ROOT, "include", "mocr_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mocr_[a-z0-9_]+)\s*\(", src)))


def _gpu():
    import torch
    return torch.cuda.is_available()


def test_library_builds_and_exports_header_symbols():
    _lib.build()
    lib = _lib.load()
    names = _header_functions()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), f"{n} declared in mocr_b200.h but not exported"
    assert sorted(_lib.exported_symbols()) == names       # the ctypes table covers the whole header
    assert lib.mocr_abi_version() == 1


def test_header_is_plain_c(tmp_path):
    c = tmp_path / "t.c"
    c.write_text('#include "mocr_b200.h"\nint main(void){ mocr_crop_t c; (void)c; return MOCR_OK; }\n')
    import subprocess
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), "-c", str(c), "-o",
                    str(tmp_path / "t.o")], check=True)


def test_resample_table_matches_oracle():
    """Host-side coefficient set-up of the preprocess kernel == Pillow's, via the oracle restatement."""
    from oracle import preprocess_np as P
    lib = _lib.load()
    for n in (1, 2, 3, 37, 223, 225, 500, 1600, 4097):
        ks = ctypes.c_int32()
        need = lib.mocr_resample_table(n, ctypes.byref(ks), None, 0)
        assert need == 224 * (2 + ks.value)
        buf = np.zeros(need, np.int32)
        assert lib.mocr_resample_table(n, ctypes.byref(ks), buf.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)), need) == need
        xmin, cnt, kk = P.resample_coeffs(n)
        assert kk.shape[1] == ks.value
        assert np.array_equal(buf[:224], xmin) and np.array_equal(buf[224:448], cnt)
        assert np.array_equal(buf[448:].reshape(224, ks.value), kk)
    assert lib.mocr_resample_table(0, ctypes.byref(ks), None, 0) < 0


def test_no_gpu_fails_loudly():
    if _gpu():
        pytest.skip("GPU present")
    lib = _lib.load()
    h = ctypes.c_void_p()
    rc = lib.mocr_create(0, 4, 300, ctypes.byref(h))
    assert rc == -5 and not h
    assert b"no CPU path" in lib.mocr_last_error(None)
    from manga_ocr_b200.engine import Engine, MocrError
    with pytest.raises(MocrError):
        Engine({}, device=0, max_batch=1)
    from manga_ocr_b200.ocr import MangaOcr
    with pytest.raises(MocrError):
        MangaOcr("random", warmup=False)


def test_create_rejects_bad_arguments():
    lib = _lib.load()
    h = ctypes.c_void_p()
    assert lib.mocr_create(0, 0, 300, ctypes.byref(h)) == -1
    assert lib.mocr_create(0, 4, 1, ctypes.byref(h)) == -1
    assert lib.mocr_create(0, 4, 513, ctypes.byref(h)) == -1
    assert lib.mocr_destroy(None) == 0


def test_product_never_imports_oracle():
    """The oracle is test infrastructure: nothing under manga_ocr_b200/ or manga_ocr/ may use it."""
    for pkg in ("manga_ocr_b200", "manga_ocr"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, pkg)):
            for fn in files:
                if fn.endswith((".py", ".cu", ".cuh", ".h")):
                    txt = open(os.path.join(dirpath, fn), encoding="utf-8").read()
                    assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), fn
                    assert "import transformers" not in txt and "from transformers" not in txt, fn
