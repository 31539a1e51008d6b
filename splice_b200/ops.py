"""Thin Python wrappers over the per-kernel C-ABI entry points (used by the host-side mirror classes and by
the parity tests). Tensors are torch CUDA tensors; only their device pointers cross the boundary."""
from __future__ import annotations

import torch

from . import _lib
from ._lib import check, cur_stream, ptr

ACT_NONE, ACT_GELU, ACT_GELU_GRAD = 0, 1, 2
IMPL_TCGEN05, IMPL_SIMT = 0, 1


def _is_bf16_rowmajor(t: torch.Tensor) -> bool:
    return t.dtype == torch.bfloat16 and t.dim() == 2 and t.stride(1) == 1 and t.is_cuda


def gemm(A: torch.Tensor, B: torch.Tensor, *, out32: torch.Tensor | None = None, out16: torch.Tensor | None = None,
         bias: torch.Tensor | None = None, residual: torch.Tensor | None = None, act: int = ACT_NONE,
         aux16: torch.Tensor | None = None, rows_per_seq: int = 0, pos: torch.Tensor | None = None,
         slice32: torch.Tensor | None = None, slice_cols: tuple[int, int] = (0, 0), impl: int = IMPL_TCGEN05,
         bn_hint: int = 0) -> None:
    """C = epilogue(A @ B.T) with A [M,K], B [N,K] bf16 row-major. See SpliceGemmArgs in include/splice_b200.h."""
    assert _is_bf16_rowmajor(A) and _is_bf16_rowmajor(B), "A and B must be 2-D bf16 CUDA tensors with unit inner stride"
    M, K = A.shape
    N, K2 = B.shape
    assert K == K2
    a = _lib.SpliceGemmArgs()
    a.A, a.lda, a.B, a.ldb = ptr(A), A.stride(0), ptr(B), B.stride(0)
    a.M, a.N, a.K = M, N, K
    if out32 is not None:
        assert out32.dtype == torch.float32 and out32.stride(-1) == 1
        a.c32, a.ldc32 = ptr(out32), out32.stride(0)
    if out16 is not None:
        assert out16.dtype == torch.bfloat16 and out16.stride(-1) == 1
        a.c16, a.ldc16 = ptr(out16), out16.stride(0)
    if bias is not None:
        assert bias.dtype == torch.float32 and bias.numel() == N
        a.bias = ptr(bias)
    if residual is not None:
        assert residual.dtype == torch.float32
        a.residual, a.ldr = ptr(residual), residual.stride(0)
    a.act = act
    if aux16 is not None:
        assert aux16.dtype == torch.bfloat16
        a.aux16, a.ldaux = ptr(aux16), aux16.stride(0)
    a.rows_per_seq = rows_per_seq
    if pos is not None:
        assert pos.dtype == torch.float32
        a.pos, a.ldpos = ptr(pos), pos.stride(0)
    if slice32 is not None:
        assert slice32.dtype == torch.float32
        a.slice32, a.ldslice = ptr(slice32), slice32.stride(0)
        a.slice_c0, a.slice_c1 = slice_cols
    a.impl, a.bn_hint = impl, bn_hint
    check(_lib.splice_gemm_bf16(a, cur_stream()), "splice_gemm_bf16")
