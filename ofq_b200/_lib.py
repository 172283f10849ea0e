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
    _fields_ = [("ptr", C.c_void_p), ("row_stride", C.c_longlong), ("k2_stride", C.c_longlong), ("k2_mod", C.c_int),
                ("dual_delta", C.c_int), ("bstride1", C.c_longlong), ("bstride2", C.c_longlong), ("mn_major", C.c_int)]


class Vec(C.Structure):
    _fields_ = [("ptr", C.c_void_p), ("period", C.c_int), ("bstride1", C.c_longlong),
                ("bstride2", C.c_longlong)]


class GemmOut(C.Structure):
    _fields_ = [("ptr", C.c_void_p), ("ld", C.c_longlong), ("bstride1", C.c_longlong),
                ("bstride2", C.c_longlong), ("accumulate", C.c_int)]


class GemmLsq(C.Structure):
    _fields_ = [("codes", C.c_void_p), ("ld_codes", C.c_longlong), ("codes16", C.c_void_p), ("ld_codes16", C.c_longlong),
                ("fmt16", C.c_int), ("res16", C.c_void_p), ("ld_res16", C.c_longlong), ("b4", C.c_void_p), ("s_eff", C.c_void_p),
                ("inv_s", C.c_void_p), ("period", C.c_int), ("nseg", C.c_int), ("seg_len", C.c_int), ("qlo", C.c_float),
                ("qhi", C.c_float), ("dot_u", C.c_void_p), ("rowdot", C.c_void_p), ("workspace", C.c_void_p)]


_p, _i, _ll, _f, _d = C.c_void_p, C.c_int, C.c_longlong, C.c_float, C.c_double

# name -> (restype, argtypes); must list every symbol include/ofq_b200.h declares (tests/test_abi.py checks)
SIGNATURES = {
    "ofq_version": (_i, []),
    "ofq_last_error": (C.c_char_p, []),
    "ofq_device_ok": (_i, []),
    "ofq_gemm": (_i, [_i, C.POINTER(Operand), C.POINTER(Operand), C.POINTER(GemmOut), _i, _i, _i, _i, _i, _i, _i,
                      C.POINTER(Vec), C.POINTER(Vec), C.POINTER(Vec), C.POINTER(Vec), _p]),
    "ofq_gemm_ex": (_i, [_i, C.POINTER(Operand), C.POINTER(Operand), C.POINTER(GemmOut), _i, _i, _i, _i, _i, _i, _i,
                         C.POINTER(Vec), C.POINTER(Vec), C.POINTER(Vec), C.POINTER(Vec), _p, _p]),
    "ofq_gemm_lsq": (_i, [C.POINTER(Operand), C.POINTER(Operand), _i, _i, _i, C.POINTER(Vec), C.POINTER(Vec), C.POINTER(Vec),
                          C.POINTER(Vec), C.POINTER(GemmLsq), _p]),
    "ofq_statsq_codes": (_i, [_p, _i, _i, _ll, _i, _p, _ll, _p, _p, _p, _p, _p, _p, _p, _p]),
    "ofq_statsq_codes_ex": (_i, [_p, _i, _i, _ll, _i, _p, _ll, _p, _p, _p, _p, _p, _p, _p, _p, _i, _p]),
    "ofq_statsq_codes_multi": (_i, [_p, _i, _i, _p]),
    "ofq_lsq_effective_scale_multi": (_i, [_p, _i, _i, _p]),
    "ofq_wqk_compose_multi": (_i, [_p, _i, _i, _i, _i, _p]),
    "ofq_lsq_effective_scale": (_i, [_p, _i, _f, _p, _p, _p]),
    "ofq_lsq_quant": (_i, [_p, _ll, _i, _ll, _p, _p, _i, _i, _i, _i, _i, _p, _ll, _p]),
    "ofq_lsq_bwd_workspace": (_ll, [_ll, _i, _i]),
    "ofq_lsq_quant_ex": (_i, [_p, _ll, _i, _ll, _p, _p, _i, _i, _i, _i, _i, _i, _p, _ll, _p, _ll, _i, _p, _p, _p]),
    "ofq_lsq_bwd_ex": (_i, [_p, _ll, _p, _ll, _ll, _i, _p, _p, _i, _i, _i, _i, _i, _i, _p, _ll, _p, _ll, _i, _p, _p, _i, _p, _p, _p]),
    "ofq_scale_from_max": (_i, [_p, _i, _p, _i, _p, _i, _f, _i, _p, _p]),
    "ofq_lsq_bwd_act": (_i, [_p, _ll, _p, _ll, _ll, _i, _p, _p, _i, _i, _i, _i, _i, _i, _p, _ll, _p, _p]),
    "ofq_lsq_bwd": (_i, [_p, _ll, _p, _ll, _ll, _i, _p, _p, _i, _i, _i, _i, _i, _p, _ll, _p, _p]),
    "ofq_lsq_bwd_finalize": (_i, [_p, _ll, _i, _i, _i, _i, _f, _p, _p, _p, _i, _p]),
    "ofq_lsq_bwd_finalize_colsum": (_i, [_p, _ll, _i, _i, _i, _i, _f, _p, _p, _p, _i, _p, _p]),
    "ofq_lsq_bwd_finalize_scale": (_i, [_p, _ll, _i, _i, _i, _i, _f, _p, _p, _p, _i, _p, _i, _p, _i, _f, _i, _p, _p]),
    "ofq_lsq_bwd_scale": (_i, [_p, _ll, _i, _i, _p, _i, _p, _i, _f, _i, _p, _p]),
    "ofq_grad_prep": (_i, [_p, _i, _i, _i, _ll, _ll, _p, _p, _i, _i, _p, _ll, _p, _i, _p, _p, _i, _p, _i, _p, _i, _p]),
    "ofq_absmax_scale_workspace": (_ll, []),
    "ofq_absmax_scale": (_i, [_p, _i, _i, _i, _ll, _ll, _p, _p, _i, _p, _i, _p, _i, _f, _i, _p, _p, _p]),
    "ofq_codes_to_bf16": (_i, [_p, _i, _i, _i, _ll, _ll, _p, _ll, _ll, _i, _p]),
    "ofq_codes_to_16": (_i, [_p, _i, _i, _i, _ll, _ll, _p, _ll, _ll, _i, _i, _p]),
    "ofq_codes_transpose": (_i, [_p, _i, _i, _i, _ll, _ll, _p, _ll, _ll, _p]),
    "ofq_codes_rowdot": (_i, [_p, _ll, _i, _ll, _i, _p, _p, _p]),
    "ofq_softmax_quant": (_i, [_p, _i, _i, _ll, _i, _p, _p, _i, _p, _i, _p, _p, _ll, _p, _p]),
    "ofq_softmax_quant_ex": (_i, [_p, _i, _i, _ll, _i, _p, _p, _i, _p, _i, _p, _p, _ll, _p, _p, _i, _p]),
    "ofq_softmax_quant_bwd": (_i, [_p, _p, _i, _i, _ll, _i, _p, _i, _f, _f, _p, _i, _p, _i, _p, _p, _ll, _p, _p, _p, _i, _p, _i, _p]),
    "ofq_softmax_quant_bwd_ex": (_i, [_p, _p, _i, _i, _ll, _i, _p, _i, _f, _f, _p, _i, _p, _i, _p, _p, _ll, _p, _p, _p, _i, _p, _i, _p, _p]),
    "ofq_qkr_attn_fwd": (_i, [_p, _p, _p, _ll, _i, _i, _i, _i, _p, _p, _p, _f, _p, _i, _p, _p, _p, _ll, _p, _p, _ll, _p, _i, _p, _p, _p]),
    "ofq_qkr_attn_bwd": (_i, [_p, _p, _p, _p, _i, _i, _i, _i, _i, _p, _p, _p, _f, _p, _p, _i, _p, _p, _p, _p, _p, _i, _f, _p, _ll, _p, _p, _p,
                              _p, _p]),
    "ofq_wqk_compose": (_i, [_p, _p, _i, _i, _i, _p, _p]),
    "ofq_wqk_compose_bwd": (_i, [_p, _p, _p, _i, _i, _i, _p, _p, _p]),
    "ofq_cga_mask": (_i, [_p, _i, _i, _i, _d, _p, _p, _p, _p]),
    "ofq_cga_adamw": (_i, [_p, _p, _p, _p, _ll, _i, _i, _i, _d, _d, _d, _d, _d, _i, _d, _p, _p, _p, _p, _p]),
    "ofq_counter_increment": (_i, [_p, _p]),
    "ofq_layernorm_fwd": (_i, [_p, _ll, _i, _p, _p, _f, _p, _p, _p, _p]),
    "ofq_layernorm_fwd_add": (_i, [_p, _p, _ll, _i, _p, _p, _f, _p, _p, _p, _p, _p]),
    "ofq_layernorm_bwd_workspace": (_ll, [_ll, _i]),
    "ofq_layernorm_bwd": (_i, [_p, _p, _p, _p, _p, _ll, _i, _p, _p, _p, _p, _p]),
    "ofq_layernorm_bwd_nmax": (_ll, [_ll, _i]),
    "ofq_layernorm_bwd_max": (_i, [_p, _p, _p, _p, _p, _ll, _i, _p, _p, _p, _p, _p, _p, _p]),
    "ofq_layernorm_bwd_res": (_i, [_p, _p, _p, _p, _p, _ll, _i, _p, _p, _p, _p, _p, _p]),
    "ofq_adamw_multi": (_i, [_p, _i, _i, _i, _d, _d, _d, _d, _p, _p]),
    "ofq_gemm_dx_lsq_workspace": (_ll, [_i, _i]),
    "ofq_gemm_dx_lsq": (_i, [_i, C.POINTER(Operand), C.POINTER(Operand), _i, _i, _i, C.POINTER(Vec), C.POINTER(Vec), _p, _ll, _p, _i, _i, _f,
                             _p, _ll, _p, _p, _p, _p, _ll, _p, _p, _p, _p, _p]),
    "ofq_lsq_bwd_finalize_parts": (_i, [_p, _ll, _p, _ll, _ll, _i, _i, _f, _p, _p, _p, _p]),
    "ofq_packed_row_bytes": (_ll, [_i, _i]),
    "ofq_pack_codes": (_i, [_p, _ll, _i, _ll, _i, _p, _ll, _p]),
    "ofq_unpack_codes": (_i, [_p, _ll, _ll, _i, _i, _p, _ll, _p]),
    "ofq_kd_loss": (_i, [_p, _p, _p, _p, _i, _i, _f, _p, _p, _p, _p, _p]),
    "ofq_cga_adamw_multi": (_i, [_p, _i, _i, _i, _i, _d, _d, _d, _d, _i, _d, _p, _p]),
}
EXPORTS = list(SIGNATURES)

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
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if an export is missing
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int) -> None:
    if rc != 0:
        msg = load().ofq_last_error().decode("utf-8", "replace")
        raise OfqError(f"ofq_b200 C-ABI call failed ({rc}): {msg}")


def stream_ptr() -> int:
    import torch
    return torch.cuda.current_stream().cuda_stream


def ptr(t) -> int:
    return t.data_ptr() if t is not None else None
