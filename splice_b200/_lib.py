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
    """Loads the library after checking that it was built from THIS tree: build.py writes a hash of the sources, headers
    and flags next to the .so (it travels with it to the GPU box); on a mismatch the library is rebuilt when nvcc is
    available (SPLICE_B200_AUTOBUILD=0 disables that) and refused otherwise - a stale library would turn an ABI or
    struct-layout change into memory corruption."""
    from . import build as _build

    stamp = _build.STAMP_PATH.read_text().strip() if _build.STAMP_PATH.exists() else None
    fresh = LIB_PATH.exists() and stamp == _build.tree_stamp()
    if not fresh:
        if os.environ.get("SPLICE_B200_AUTOBUILD", "1") == "1":
            try:
                _build.build()
                fresh = True
            except Exception as e:  # noqa: BLE001
                raise ImportError(f"{LIB_PATH} is missing or stale and could not be rebuilt: {e}") from e
        if not fresh:
            raise ImportError(
                f"{LIB_PATH} is missing or was built from different sources: build it with `python -m splice_b200.build` "
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


class SpliceVitDesc(C.Structure):
    _fields_ = [("patch", c_int), ("dim", c_int), ("heads", c_int), ("depth", c_int), ("n_pos", c_int), ("ln_eps", c_float)]


class SpliceImage(C.Structure):
    _fields_ = [("data", c_void_p), ("h", c_int), ("w", c_int)]


class SpliceVitForwardArgs(C.Structure):
    _fields_ = [
        ("images", C.POINTER(SpliceImage)), ("n_images", c_int),
        ("out_h", c_int), ("out_w", c_int),
        ("pos", c_void_p),
        ("n_grad", c_int), ("slot", c_int),
        ("keys32", c_void_p), ("cls32", c_void_p), ("qkv32_all", c_void_p), ("block32_all", c_void_p),
        ("gemm_impl", c_int),
        ("pre_normalized", c_int),
        ("use_graph", c_int),
        ("n_full", c_int),
    ]


class SpliceVitBackwardArgs(C.Structure):
    _fields_ = [
        ("slot", c_int),
        ("dkeys32", c_void_p), ("dcls32", c_void_p),
        ("grads", C.POINTER(SpliceImage)),
        ("gemm_impl", c_int),
        ("use_graph", c_int),
        ("dblock32_layers", C.POINTER(c_void_p)),
        ("dqkv32_layers", C.POINTER(c_void_p)),
    ]


class SpliceProfileEntry(C.Structure):
    _fields_ = [("count", c_longlong), ("ms", C.c_double), ("flops", C.c_double), ("bytes", C.c_double)]


GEN_PARAMS, GEN_BN = 112, 30


class SpliceGenPointers(C.Structure):
    _fields_ = [("param", c_void_p * GEN_PARAMS), ("grad", c_void_p * GEN_PARAMS),
                ("running_mean", c_void_p * GEN_BN), ("running_var", c_void_p * GEN_BN),
                ("num_batches_tracked", c_void_p * GEN_BN)]


GENX_MAX_SCALES = 8


class SpliceGenXConfig(C.Structure):
    _fields_ = [("n_scales", c_int), ("in_channels", c_int), ("out_channels", c_int),
                ("ch_down", c_int * GENX_MAX_SCALES), ("ch_up", c_int * GENX_MAX_SCALES), ("ch_skip", c_int * GENX_MAX_SCALES),
                ("k_down", c_int * GENX_MAX_SCALES), ("k_up", c_int * GENX_MAX_SCALES),
                ("k_skip", c_int), ("reflect", c_int), ("sigmoid", c_int)]


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

splice_layernorm_fwd = _sig("splice_layernorm_fwd", c_int,
                            [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_float, c_void_p])
splice_layernorm_bwd = _sig("splice_layernorm_bwd", c_int,
                            [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p])
splice_attention_fwd = _sig("splice_attention_fwd", c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p])
splice_attention_bwd = _sig("splice_attention_bwd", c_int,
                            [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p])
splice_resized_hw = _sig("splice_resized_hw", None, [c_int, c_int, c_int, c_int, C.POINTER(c_int), C.POINTER(c_int)])
splice_preprocess_fwd = _sig("splice_preprocess_fwd", c_int,
                             [c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_int, c_int, c_void_p])
splice_resize_normalize = _sig("splice_resize_normalize", c_int,
                               [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_int, c_void_p])
splice_preprocess_bwd = _sig("splice_preprocess_bwd", c_int,
                             [c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_int, c_void_p])
splice_vit_packed_floats = _sig("splice_vit_packed_floats", c_size_t, [C.POINTER(SpliceVitDesc)])
splice_vit_create = _sig("splice_vit_create", c_int,
                         [C.POINTER(c_void_p), C.POINTER(SpliceVitDesc), c_void_p, c_size_t, c_void_p])
splice_vit_destroy = _sig("splice_vit_destroy", c_int, [c_void_p])
splice_vit_forward = _sig("splice_vit_forward", c_int, [c_void_p, C.POINTER(SpliceVitForwardArgs), c_void_p])
splice_vit_backward = _sig("splice_vit_backward", c_int, [c_void_p, C.POINTER(SpliceVitBackwardArgs), c_void_p])
splice_vit_profile_enable = _sig("splice_vit_profile_enable", c_int, [c_void_p, c_int])
splice_vit_profile_read = _sig("splice_vit_profile_read", c_int, [c_void_p, C.POINTER(SpliceProfileEntry), c_int])
splice_loss_ssim = _sig("splice_loss_ssim", c_int,
                        [c_void_p, c_void_p, c_void_p, c_int, c_float, c_void_p, c_void_p, c_int, c_void_p])
splice_loss_mse = _sig("splice_loss_mse", c_int,
                       [c_void_p, c_void_p, c_void_p, c_int, c_int, c_float, c_void_p, c_void_p, c_void_p])
splice_keys_self_sim = _sig("splice_keys_self_sim", c_int, [c_void_p, c_void_p, c_int, c_void_p, c_int, c_void_p])
splice_weighted_total = _sig("splice_weighted_total", c_int, [c_void_p, C.POINTER(c_float), c_int, c_void_p, c_void_p])
splice_debug_spin = _sig("splice_debug_spin", c_int, [c_float, c_void_p])

splice_gen_create = _sig("splice_gen_create", c_int, [C.POINTER(c_void_p)])
splice_gen_destroy = _sig("splice_gen_destroy", c_int, [c_void_p])
splice_gen_forward = _sig("splice_gen_forward", c_int,
                          [c_void_p, C.POINTER(SpliceGenPointers), c_void_p, c_int, c_int, c_int, c_void_p, c_int, c_int, c_int,
                           c_void_p])
splice_gen_update_running = _sig("splice_gen_update_running", c_int, [c_void_p, C.POINTER(SpliceGenPointers), c_int, c_void_p])
splice_gen_set_graphs = _sig("splice_gen_set_graphs", c_int, [c_void_p, c_int])
splice_gen_backward = _sig("splice_gen_backward", c_int, [c_void_p, C.POINTER(SpliceGenPointers), c_void_p, c_int, c_int, c_void_p])
splice_gen_debug_conv = _sig("splice_gen_debug_conv", c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_int, c_int, c_void_p,
                                                               c_void_p, c_int, c_int, c_void_p])
GENX_SIGNATURES = {   # shared with the CPU emulation build of the same entry points (tests/emu)
    "splice_genx_create": (c_int, [C.POINTER(SpliceGenXConfig), C.POINTER(c_void_p)]),
    "splice_genx_destroy": (c_int, [c_void_p]),
    "splice_genx_counts": (c_int, [c_void_p, C.POINTER(c_int), C.POINTER(c_int)]),
    "splice_genx_bind": (c_int, [c_void_p, C.POINTER(c_void_p), C.POINTER(c_void_p), C.POINTER(c_void_p), C.POINTER(c_void_p),
                                 C.POINTER(c_void_p)]),
    "splice_genx_forward": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_int, c_int, c_void_p]),
    "splice_genx_backward": (c_int, [c_void_p, c_void_p, c_int, c_void_p]),
    "splice_genx_set_graphs": (c_int, [c_void_p, c_int]),
}
splice_genx_create = _sig("splice_genx_create", *GENX_SIGNATURES["splice_genx_create"])
splice_genx_destroy = _sig("splice_genx_destroy", *GENX_SIGNATURES["splice_genx_destroy"])
splice_genx_counts = _sig("splice_genx_counts", *GENX_SIGNATURES["splice_genx_counts"])
splice_genx_bind = _sig("splice_genx_bind", *GENX_SIGNATURES["splice_genx_bind"])
splice_genx_forward = _sig("splice_genx_forward", *GENX_SIGNATURES["splice_genx_forward"])
splice_genx_backward = _sig("splice_genx_backward", *GENX_SIGNATURES["splice_genx_backward"])
splice_genx_set_graphs = _sig("splice_genx_set_graphs", *GENX_SIGNATURES["splice_genx_set_graphs"])
splice_accumulate = _sig("splice_accumulate", c_int, [c_void_p, C.POINTER(c_void_p), c_int, c_size_t, c_void_p])
splice_adam_step = _sig("splice_adam_step", c_int,
                        [C.POINTER(c_void_p), C.POINTER(c_void_p), C.POINTER(c_void_p), C.POINTER(c_void_p),
                         C.POINTER(c_int), c_int, c_int, c_float, c_float, c_float, c_float, c_void_p])

# every symbol include/splice_b200.h declares (tests/test_abi.py checks the header against this list)
EXPORTS = [
    "splice_version", "splice_last_error", "splice_launch_count", "splice_launch_count_reset",
    "splice_gemm_bf16",
    "splice_layernorm_fwd", "splice_layernorm_bwd", "splice_attention_fwd", "splice_attention_bwd",
    "splice_resized_hw", "splice_preprocess_fwd", "splice_preprocess_bwd", "splice_resize_normalize",
    "splice_vit_packed_floats", "splice_vit_create", "splice_vit_destroy", "splice_vit_forward", "splice_vit_backward",
    "splice_loss_ssim", "splice_loss_mse", "splice_keys_self_sim", "splice_weighted_total", "splice_debug_spin",
    "splice_gen_create", "splice_gen_destroy", "splice_gen_forward", "splice_gen_backward", "splice_gen_set_graphs",
    "splice_accumulate", "splice_gen_update_running", "splice_gen_debug_conv",
    "splice_genx_create", "splice_genx_destroy", "splice_genx_counts", "splice_genx_bind", "splice_genx_forward",
    "splice_genx_backward", "splice_genx_set_graphs",
    "splice_adam_step", "splice_vit_profile_enable", "splice_vit_profile_read",
]


ABI_VERSION = 100
if splice_version() != ABI_VERSION:
    raise ImportError(f"{LIB_PATH} reports ABI version {splice_version()}, this package binds version {ABI_VERSION}")
_missing = [name for name in EXPORTS if not hasattr(lib, name)]
if _missing:
    raise ImportError(f"{LIB_PATH} does not export {_missing}")


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
