"""Thin Python wrappers over the C ABI (one function per exported kernel family).

Tensors are containers only: every function validates dtype / contiguity, passes raw device pointers and the
current CUDA stream, and raises NativeError on a non-zero status.  No function here computes anything in PyTorch.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _native as N

bf16 = torch.bfloat16
f32 = torch.float32


def _ld(t: torch.Tensor) -> int:
    assert t.dim() == 2 and t.stride(1) == 1, "expect 2-D row-major (unit inner stride)"
    return t.stride(0)


def gemm(A: torch.Tensor, B: torch.Tensor, out: torch.Tensor, *, a_mn: bool = False, b_mn: bool = False,
         bias: torch.Tensor | None = None, gelu: bool = False, store_pre: torch.Tensor | None = None,
         dgelu_of: torch.Tensor | None = None, gamma: torch.Tensor | None = None,
         resid: torch.Tensor | None = None, accum: bool = False, alpha: float = 1.0, tile_n: int = 0) -> torch.Tensor:
    """out[M,N] = epilogue(alpha * A.B) on the tcgen05 tensor cores (d3_gemm_bf16).

    A is [M,K] (a_mn=False) or stored transposed [K,M] (a_mn=True); B is [N,K] (b_mn=False) or [K,N] (b_mn=True).
    """
    l = N.init()
    assert A.dtype == bf16 and B.dtype == bf16
    M, K = (A.shape[1], A.shape[0]) if a_mn else (A.shape[0], A.shape[1])
    Nn, Kb = (B.shape[1], B.shape[0]) if b_mn else (B.shape[0], B.shape[1])
    assert K == Kb, f"contraction mismatch {K} vs {Kb}"
    assert out.shape[0] == M and out.shape[1] == Nn and out.dtype in (bf16, f32)
    flags = 0
    ep = N.GemmEpilogue()
    ep.out = out.data_ptr(); ep.ld_out = _ld(out)
    if out.dtype == f32:
        flags |= N.EP_OUT_F32
    if bias is not None:
        assert bias.dtype == f32 and bias.numel() == Nn
        flags |= N.EP_BIAS; ep.bias = bias.data_ptr()
    if gelu:
        flags |= N.EP_GELU
    if store_pre is not None:
        assert store_pre.dtype == bf16 and store_pre.shape == out.shape
        flags |= N.EP_STORE_PRE; ep.aux_out = store_pre.data_ptr(); ep.ld_aux = _ld(store_pre)
    if dgelu_of is not None:
        assert dgelu_of.dtype == bf16 and dgelu_of.shape == out.shape and store_pre is None
        flags |= N.EP_MUL_DGELU; ep.aux_in = dgelu_of.data_ptr(); ep.ld_aux = _ld(dgelu_of)
    if gamma is not None:
        assert gamma.dtype == f32 and gamma.numel() == Nn
        flags |= N.EP_GAMMA; ep.gamma = gamma.data_ptr()
    if resid is not None:
        assert resid.dtype == f32 and resid.shape == out.shape
        flags |= N.EP_RESID; ep.resid = resid.data_ptr(); ep.ld_resid = _ld(resid)
    if accum:
        assert out.dtype == f32
        flags |= N.EP_ACCUM
    ep.flags = flags
    ep.alpha = float(alpha)
    N.check(l.d3_gemm_bf16(N.ptr(A), _ld(A), int(a_mn), N.ptr(B), _ld(B), int(b_mn), M, Nn, K, C.byref(ep),
                           int(tile_n), N.stream_ptr()), "d3_gemm_bf16")
    return out
