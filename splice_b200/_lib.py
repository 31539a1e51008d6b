"""ctypes binding of include/splice_b200.h.

The shared library is the product: if it is missing, or fails to load, importing this module raises — there
is no eager/PyTorch fallback anywhere in the package (north_star: "no CPU fallback").
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

_PKG = Path(__file__).resolve().parent
LIB_PATH = _PKG / "libsplice_b200.so"


class SpliceError(RuntimeError):
    """Raised when a C-ABI call returns a negative status (message from splice_last_error)."""


def _load() -> C.CDLL:
    if not LIB_PATH.exists():
        if os.environ.get("SPLICE_B200_AUTOBUILD", "1") == "1":
            from . import build as _build

            _build.build()
        if not LIB_PATH.exists():
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -m splice_b200.build` "
                "(or __graft_entry__.build()); splice_b200 has no fallback path"
            )
    return C.CDLL(str(LIB_PATH))


lib = _load()

c_void_p, c_int, c_float, c_longlong, c_size_t = C.c_void_p, C.c_int, C.c_float, C.c_longlong, C.c_size_t


class SpliceGemmArgs(C.Structure):
    _fields_ = [
        ("A", c_void_p), ("lda", c_int),
        ("B", c_void_p), ("ldb", c_int),
        ("M", c_int), ("N", c_int), ("K", c_int),
        ("c32", c_void_p), ("ldc32", c_int),
        ("c16", c_void_p), ("ldc16", c_int),
        ("bias", c_void_p),
        ("residual", c_void_p), ("ldr", c_int),
        ("act", c_int),
        ("aux16", c_void_p), ("ldaux", c_int),
        ("rows_per_seq", c_int),
        ("pos", c_void_p), ("ldpos", c_int),
        ("slice32", c_void_p), ("slice_c0", c_int), ("slice_c1", c_int), ("ldslice", c_int),
        ("impl", c_int),
        ("bn_hint", c_int),
    ]


def _sig(name, restype, argtypes):
    fn = getattr(lib, name)
    fn.restype = restype
    fn.argtypes = argtypes
    return fn


splice_version = _sig("splice_version", c_int, [])
splice_last_error = _sig("splice_last_error", C.c_char_p, [])
splice_launch_count = _sig("splice_launch_count", c_longlong, [])
splice_launch_count_reset = _sig("splice_launch_count_reset", None, [])
splice_gemm_bf16 = _sig("splice_gemm_bf16", c_int, [C.POINTER(SpliceGemmArgs), c_void_p])

# every symbol include/splice_b200.h declares (tests/test_abi.py checks the header against this list)
EXPORTS = [
    "splice_version", "splice_last_error", "splice_launch_count", "splice_launch_count_reset",
    "splice_gemm_bf16",
]


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = splice_last_error().decode("utf-8", "replace")
        raise SpliceError(f"{what or 'splice_b200 call'} failed (rc={rc}): {msg}")


def ptr(t) -> int | None:
    """Device pointer of a torch tensor (None -> NULL)."""
    return None if t is None else t.data_ptr()


def cur_stream() -> int:
    import torch

    return torch.cuda.current_stream().cuda_stream
