"""ctypes binding of libofq_b200.so (include/ofq_b200.h).

PyTorch is plumbing here: it owns device memory and streams; every kernel is reached through the
C-ABI with raw pointers.  There is no CPU fallback — if the library is missing this module raises.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

_HERE = Path(__file__).resolve().parent
LIB_PATH = _HERE / "libofq_b200.so"


class OfqError(RuntimeError):
    pass


class Operand(C.Structure):
    _fields_ = [("ptr", C.c_void_p), ("row_stride", C.c_longlong), ("k2_stride", C.c_longlong),
                ("bstride1", C.c_longlong), ("bstride2", C.c_longlong)]


class Vec(C.Structure):
    _fields_ = [("ptr", C.c_void_p), ("period", C.c_int), ("bstride1", C.c_longlong),
                ("bstride2", C.c_longlong)]


class GemmOut(C.Structure):
    _fields_ = [("ptr", C.c_void_p), ("ld", C.c_longlong), ("bstride1", C.c_longlong),
                ("bstride2", C.c_longlong), ("accumulate", C.c_int)]


_lib = None


def load(build_if_missing: bool = False):
    """Load the shared library (once). Raises OfqError if it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        if build_if_missing or os.environ.get("OFQ_B200_AUTOBUILD") == "1":
            from . import build as _build
            _build.build()
        else:
            raise OfqError(
                f"{LIB_PATH} is missing: run `python -m ofq_b200.build` (or __graft_entry__.build()). "
                "ofq_b200 has no CPU / PyTorch fallback.")
    lib = C.CDLL(str(LIB_PATH))
    lib.ofq_last_error.restype = C.c_char_p
    for name in EXPORTS:
        getattr(lib, name)  # raises AttributeError if an export is missing
    for name in EXPORTS:
        if name not in ("ofq_last_error",):
            getattr(lib, name).restype = C.c_int
    _lib = lib
    return lib


# every symbol include/ofq_b200.h declares (tests/test_abi.py cross-checks this list against the header)
EXPORTS = [
    "ofq_version", "ofq_last_error", "ofq_device_ok", "ofq_gemm",
    "ofq_statsq_codes", "ofq_lsq_effective_scale", "ofq_lsq_quant", "ofq_lsq_bwd", "ofq_lsq_bwd_finalize",
    "ofq_cvt_bf16", "ofq_cvt_bf16_t", "ofq_codes_bf16_t", "ofq_colsum",
    "ofq_softmax_quant", "ofq_softmax_quant_bwd",
    "ofq_wqk_compose", "ofq_wqk_compose_bwd",
    "ofq_cga_mask", "ofq_cga_adamw",
]


def check(rc: int) -> None:
    if rc != 0:
        msg = load().ofq_last_error().decode("utf-8", "replace")
        raise OfqError(f"ofq_b200 C-ABI call failed ({rc}): {msg}")


def stream_ptr() -> C.c_void_p:
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr(t) -> C.c_void_p:
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)
