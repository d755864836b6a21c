"""Build and load ``libmocr_b200.so`` (the C ABI of include/mocr_b200.h) with ctypes.

There is no fallback: if the library is missing, or no sm_100 device is present,
the engine raises - it never computes on the CPU.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from ctypes import POINTER, c_char_p, c_double, c_float, c_int, c_int32, c_int64, c_uint8, c_void_p

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_PATH = os.environ.get("MOCR_LIB_PATH") or os.path.join(HERE, "libmocr_b200.so")   # (override: A/B runs of two builds)
HEADER = os.path.join(os.path.dirname(HERE), "include", "mocr_b200.h")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-shared", "-Xcompiler", "-fPIC",
]


class mocr_crop_t(ctypes.Structure):
    _fields_ = [("data", c_void_p), ("height", c_int32), ("width", c_int32), ("stride", c_int32), ("channels", c_int32)]


class mocr_region_t(ctypes.Structure):
    _fields_ = [("left", c_int32), ("top", c_int32), ("right", c_int32), ("bottom", c_int32), ("polygon", POINTER(c_int32)),
                ("n_points", c_int32), ("rotate", c_int32)]


def sources():
    return [os.path.join(CSRC, "engine.cu")]


def _stale() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [HEADER]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile the CUDA extension in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    if not force and not _stale():
        return LIB_PATH
    nvcc = os.environ.get("NVCC", "nvcc")
    cmd = [nvcc] + NVCC_FLAGS + os.environ.get("NVCC_EXTRA", "").split() + sources() + ["-o", LIB_PATH]
    if verbose:
        print(" ".join(cmd))
    subprocess.run(cmd, check=True)
    return LIB_PATH


_SIGNATURES = {
    "mocr_abi_version": (c_int, []),
    "mocr_create": (c_int, [c_int, c_int, c_int, POINTER(c_void_p)]),
    "mocr_destroy": (c_int, [c_void_p]),
    "mocr_set_weight": (c_int, [c_void_p, c_char_p, POINTER(c_float), POINTER(c_int64), c_int]),
    "mocr_finalize_weights": (c_int, [c_void_p]),
    "mocr_recognize": (c_int, [c_void_p, POINTER(mocr_crop_t), c_int, c_int, c_int, POINTER(c_int32), POINTER(c_int32)]),
    "mocr_recognize_regions": (c_int, [c_void_p, POINTER(mocr_crop_t), POINTER(mocr_region_t), c_int, c_int, c_int, POINTER(c_int32),
                                       POINTER(c_int32)]),
    "mocr_decode_beam": (c_int, [c_void_p, c_int, c_int, c_int, c_float, c_int, POINTER(c_int32), POINTER(c_int32), POINTER(c_float)]),
    "mocr_recognize_beam": (c_int, [c_void_p, POINTER(mocr_crop_t), c_int, c_int, c_int, c_int, c_int, c_float, c_int, POINTER(c_int32),
                                    POINTER(c_int32), POINTER(c_float)]),
    "mocr_recognize_regions_beam": (c_int, [c_void_p, POINTER(mocr_crop_t), POINTER(mocr_region_t), c_int, c_int, c_int, c_int, c_int, c_float,
                                            c_int, POINTER(c_int32), POINTER(c_int32), POINTER(c_float)]),
    "mocr_beam_create": (c_int, [c_int, c_int, c_int, c_int, c_float, c_int, POINTER(c_void_p)]),
    "mocr_beam_destroy": (c_int, [c_void_p]),
    "mocr_beam_banned": (c_int, [c_void_p, c_int, POINTER(c_int32), c_int]),
    "mocr_beam_step": (c_int, [c_void_p, POINTER(c_float), POINTER(c_int32), POINTER(c_int32), POINTER(c_int32)]),
    "mocr_beam_result": (c_int, [c_void_p, POINTER(c_int32), POINTER(c_int32), POINTER(c_float)]),
    "mocr_session_begin": (c_int, [c_void_p, c_int, c_int, c_int]),
    "mocr_session_add": (c_int, [c_void_p, POINTER(mocr_crop_t), c_int, POINTER(c_int32)]),
    "mocr_session_run": (c_int, [c_void_p, c_int, POINTER(c_int32)]),
    "mocr_session_fetch": (c_int, [c_void_p, POINTER(c_int32), c_int, POINTER(c_int32), c_int]),
    "mocr_session_end": (c_int, [c_void_p]),
    "mocr_session_rows": (c_int, [c_void_p, c_int]),
    "mocr_stage_crops": (c_int, [c_void_p, POINTER(mocr_crop_t), c_int, c_int]),
    "mocr_stage_regions": (c_int, [c_void_p, POINTER(mocr_crop_t), POINTER(mocr_region_t), c_int, c_int]),
    "mocr_get_region_mask": (c_int, [c_void_p, c_int, POINTER(c_uint8)]),
    "mocr_preprocess": (c_int, [c_void_p]),
    "mocr_encode": (c_int, [c_void_p]),
    "mocr_decode_greedy": (c_int, [c_void_p, c_int, POINTER(c_int32)]),
    "mocr_fetch_ids": (c_int, [c_void_p, POINTER(c_int32), POINTER(c_int32)]),
    "mocr_run_resident": (c_int, [c_void_p, c_int]),
    "mocr_set_taps": (c_int, [c_void_p, c_int]),
    "mocr_get_pixels_u8": (c_int, [c_void_p, POINTER(c_uint8)]),
    "mocr_get_pixel_values": (c_int, [c_void_p, POINTER(c_float)]),
    "mocr_get_encoder_hidden": (c_int, [c_void_p, POINTER(c_float)]),
    "mocr_get_step_logits": (c_int, [c_void_p, POINTER(c_float)]),
    "mocr_stream": (c_void_p, [c_void_p]),
    "mocr_sync": (c_int, [c_void_p]),
    "mocr_launch_count": (c_int64, [c_void_p]),
    "mocr_last_steps": (c_int, [c_void_p]),
    "mocr_set_option": (c_int, [c_void_p, c_char_p, c_int]),
    "mocr_time_kernel": (c_int, [c_void_p, c_char_p, c_int, POINTER(c_float), POINTER(c_double), POINTER(c_double)]),
    "mocr_test_gemm": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, POINTER(c_float), POINTER(c_float), POINTER(c_float),
                               POINTER(c_float), POINTER(c_float), POINTER(c_int32)]),
    "mocr_test_encoder_attention": (c_int, [c_void_p, c_int, POINTER(c_float), POINTER(c_float)]),
    "mocr_test_decode_attention": (c_int, [c_void_p, c_int, c_int, c_int, POINTER(c_int32), POINTER(c_float), POINTER(c_float), POINTER(c_float),
                                           POINTER(c_float), POINTER(c_float), POINTER(c_float), POINTER(c_float), POINTER(c_float)]),
    "mocr_test_stage_gemm": (c_int, [c_void_p, c_int, c_int, c_int, c_int, POINTER(c_float), POINTER(c_float), POINTER(c_float), POINTER(c_float),
                                     POINTER(c_float), POINTER(c_float), c_int, POINTER(c_float), POINTER(c_int32)]),
    "mocr_resample_table": (c_int, [c_int, POINTER(c_int32), POINTER(c_int32), c_int]),
    "mocr_get_decode_profile": (c_int, [c_void_p, POINTER(c_int64), c_int]),
    "mocr_last_error": (c_char_p, [c_void_p]),
}

_lib = None


def exported_symbols():
    """Names every build of the library must export (checked against the header by the tests)."""
    return sorted(_SIGNATURES)


def load() -> ctypes.CDLL:
    """dlopen the library (no CUDA call is made by loading) and type its entry points."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a). This engine has no CPU or PyTorch fallback.")
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(lib, name)       # AttributeError if the symbol is not exported
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib
