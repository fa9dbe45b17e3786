"""ctypes binding of the in-tree C-ABI library (include/mirror_b200.h).

The product path has NO fallback: if the library cannot be loaded (or built
from the in-tree sources with nvcc) importing any kernel raises immediately.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmirror_b200.so")
_lib = None


class GemmArgs(ctypes.Structure):
    _fields_ = [
        ("a", ctypes.c_void_p), ("b", ctypes.c_void_p),
        ("a_mn_major", ctypes.c_int32), ("b_mn_major", ctypes.c_int32),
        ("lda", ctypes.c_int64), ("ldb", ctypes.c_int64),
        ("a_bs1", ctypes.c_int64), ("a_bs2", ctypes.c_int64), ("b_bs1", ctypes.c_int64), ("b_bs2", ctypes.c_int64),
        ("M", ctypes.c_int32), ("N", ctypes.c_int32), ("K", ctypes.c_int32),
        ("batch1", ctypes.c_int32), ("batch2", ctypes.c_int32),
        ("alpha", ctypes.c_float),
        ("bias", ctypes.c_void_p),
        ("act", ctypes.c_int32),
        ("drop_p", ctypes.c_float),
        ("drop_seed", ctypes.c_uint64),
        ("res", ctypes.c_void_p),
        ("res_is_bf16", ctypes.c_int32),
        ("gamma", ctypes.c_float),
        ("ldr", ctypes.c_int64), ("r_bs1", ctypes.c_int64), ("r_bs2", ctypes.c_int64),
        ("beta", ctypes.c_float),
        ("out_f32", ctypes.c_void_p),
        ("ldc32", ctypes.c_int64), ("c32_bs1", ctypes.c_int64), ("c32_bs2", ctypes.c_int64),
        ("out_bf16", ctypes.c_void_p),
        ("ldc16", ctypes.c_int64), ("c16_bs1", ctypes.c_int64), ("c16_bs2", ctypes.c_int64),
        ("split_k", ctypes.c_int32),
    ]


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            from . import build as _build  # compile in-tree; raises if nvcc is missing
            _build.build()
        _lib = ctypes.CDLL(LIB_PATH)
        _lib.mirror_last_error.restype = ctypes.c_char_p
    return _lib


def check(rc: int, what: str):
    if rc != 0:
        raise RuntimeError(f"mirror_b200 {what} failed (code {rc}): {lib().mirror_last_error().decode()}")
