"""ctypes binding of the C ABI declared in include/audiolab_b200.h.

There is NO fallback: if the shared library is missing or a call fails, a RuntimeError is
raised.  The product path never routes through PyTorch ops or the CPU oracle for the
spectral hot path.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libaudiolab_b200.so")

_lib = None
_lock = threading.Lock()

c_i64 = C.c_int64
c_f32p = C.c_void_p   # device pointers are passed as integers
c_i64p = C.c_void_p
c_i32p = C.c_void_p

# name -> (restype, argtypes); kept in the order of include/audiolab_b200.h
SIGNATURES = {
    "al_version": (C.c_int, []),
    "al_last_error": (C.c_char_p, []),
    "al_launch_count": (c_i64, []),
    "al_plan_create": (C.c_int, [C.c_int, C.c_int, C.c_void_p, C.c_int, C.POINTER(C.c_void_p)]),
    "al_plan_destroy": (C.c_int, [C.c_void_p]),
    "al_stft": (C.c_int, [C.c_void_p, c_f32p, c_i64, c_i64, C.c_int, c_i64p, c_i64, c_i64, C.c_int, C.c_int,
                          C.c_int, C.c_int, c_f32p, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "al_istft": (C.c_int, [C.c_void_p, c_f32p, c_f32p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                           C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, c_f32p, c_f32p, c_i64, c_i64, c_i64p,
                           c_i64, c_i64, c_i64, C.c_void_p]),
    "al_ola_gather": (C.c_int, [c_f32p, C.c_int, C.c_int, C.c_int, C.c_int, c_i64p, c_i32p, c_f32p, c_i32p, c_i64, c_i64,
                                c_i64, c_f32p, C.c_int, C.c_float, C.c_float, c_f32p, c_i64, C.c_void_p]),
    "al_resample_poly": (C.c_int, [c_f32p, c_i64, c_f32p, c_i64, C.c_int, c_i64, c_i64, C.c_int, C.c_int, c_f32p,
                                   C.c_int, C.c_void_p]),
    "al_sub": (C.c_int, [c_f32p, c_f32p, c_f32p, c_i64, C.c_void_p]),
    "al_rmsnorm_bf16": (C.c_int, [C.c_void_p, c_f32p, c_f32p, C.c_void_p, c_i64, C.c_int, C.c_float, C.c_float,
                                  C.c_void_p]),
    "al_rotary_bf16": (C.c_int, [C.c_void_p, C.c_void_p, c_f32p, c_i64, C.c_int, C.c_int, c_i64, C.c_int, C.c_void_p]),
    "al_gate_sigmoid_bf16": (C.c_int, [C.c_void_p, C.c_void_p, c_i64, C.c_int, C.c_int, C.c_void_p]),
    "al_gate_sigmoid_ld_bf16": (C.c_int, [C.c_void_p, C.c_void_p, c_i64, c_i64, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "al_gelu_bf16": (C.c_int, [C.c_void_p, c_i64, C.c_void_p]),
    "al_band_attention_bf16": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, c_i64, c_f32p, c_i64, C.c_int,
                                         C.c_int, C.c_int, C.c_float, C.c_int, C.c_void_p]),
    "al_time_attention_bf16": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, c_i64, c_i64, C.c_int,
                                         C.c_int, C.c_int, C.c_int, C.c_float, C.c_int, C.c_void_p]),
    "al_gemm_bf16": (C.c_int, [C.c_void_p, C.c_void_p]),
    "al_band_norm": (C.c_int, [c_f32p, c_i64, c_f32p, c_i32p, C.c_int, C.c_void_p, c_i64, c_i64, C.c_float, C.c_int, C.c_void_p]),
    "al_resid_prepare": (C.c_int, [c_f32p, c_f32p, c_f32p, c_f32p, C.c_void_p, c_f32p, c_i64, C.c_int, C.c_int, C.c_float,
                                   C.c_int, C.c_void_p]),
}
OPTIONAL = set()


def lib() -> C.CDLL:
    """Load libaudiolab_b200.so (built in-tree by audiolab_b200.build).  Raises if absent."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: run `python -m audiolab_b200.build` (or __graft_entry__.build()). "
                "audiolab_b200 has no CPU / PyTorch fallback for the spectral hot path.")
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            try:
                fn = getattr(handle, name)
            except AttributeError:
                if name in OPTIONAL:
                    continue
                raise RuntimeError(f"{LIB_PATH} does not export {name}; rebuild it")
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = lib().al_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"{what} failed (rc={rc}): {msg}")


def launch_count() -> int:
    return int(lib().al_launch_count())
