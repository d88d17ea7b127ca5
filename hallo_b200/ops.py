"""Python call wrappers over the C ABI: torch tensors in, raw pointers + sizes out.

Every function launches asynchronously on torch's current CUDA stream (so the calls can be
captured into a CUDA graph) and writes into caller-provided output tensors.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import lib as L


def _rowmajor_ld(t: torch.Tensor) -> int:
    assert t.dim() == 2 and t.stride(1) == 1, f"need row-major 2-D view, got {tuple(t.shape)} {t.stride()}"
    return t.stride(0)


def gemm(a: torch.Tensor, w: torch.Tensor, out: torch.Tensor, *, bias: Optional[torch.Tensor] = None,
         residual: Optional[torch.Tensor] = None, row_scale: Optional[torch.Tensor] = None,
         group_bias: Optional[torch.Tensor] = None, rows_per_group: int = 0, alpha: float = 1.0,
         geglu: bool = False, a2: Optional[torch.Tensor] = None) -> torch.Tensor:
    """out[M, N(/2)] = epilogue(cat([a, a2], 1) @ w.T); see include/hallo_b200.h."""
    p = L.GemmParams()
    p.dtype = L.dtype_code(a.dtype)
    M, K1 = a.shape
    N, K = w.shape
    p.M, p.N, p.K = M, N, K
    p.A, p.lda = L.ptr(a), _rowmajor_ld(a)
    if a2 is not None:
        assert a2.shape[0] == M and K1 + a2.shape[1] == K
        p.A2, p.lda2, p.K1 = L.ptr(a2), _rowmajor_ld(a2), K1
    else:
        assert K1 == K, (K1, K)
    p.W, p.ldw = L.ptr(w), _rowmajor_ld(w)
    p.C, p.ldc = L.ptr(out), _rowmajor_ld(out)
    assert out.shape[0] == M and out.shape[1] == (N // 2 if geglu else N), (out.shape, M, N)
    p.bias = L.ptr(bias)
    if group_bias is not None:
        p.group_bias, p.ld_group_bias, p.rows_per_group = L.ptr(group_bias), _rowmajor_ld(group_bias), rows_per_group
    p.row_scale = L.ptr(row_scale)
    if residual is not None:
        p.residual, p.ldr = L.ptr(residual), _rowmajor_ld(residual)
    p.alpha = alpha
    p.flags = L.HB_EPI_GEGLU if geglu else 0
    L.check(L.load().hallo_b200_gemm(C.byref(p), L.current_stream()), "gemm")
    return out


def conv3x3(x: torch.Tensor, w_packed: torch.Tensor, out: torch.Tensor, *, bias: Optional[torch.Tensor] = None,
            residual: Optional[torch.Tensor] = None, group_bias: Optional[torch.Tensor] = None,
            rows_per_group: int = 0) -> torch.Tensor:
    """x: NHWC [n, h, w, cin]; w_packed: [cout, 9*cin] ([cout][kh][kw][cin]); out: [n*h*w, cout]."""
    n, h, w_, cin = x.shape
    assert x.stride(3) == 1 and x.stride(1) == w_ * x.stride(2) and x.stride(0) == h * x.stride(1)
    p = L.GemmParams()
    p.dtype = L.dtype_code(x.dtype)
    p.M, p.N, p.K = n * h * w_, w_packed.shape[0], w_packed.shape[1]
    assert p.K == 9 * cin
    p.A, p.lda = L.ptr(x), x.stride(2)
    p.W, p.ldw = L.ptr(w_packed), _rowmajor_ld(w_packed)
    p.C, p.ldc = L.ptr(out), _rowmajor_ld(out)
    p.bias = L.ptr(bias)
    if group_bias is not None:
        p.group_bias, p.ld_group_bias, p.rows_per_group = L.ptr(group_bias), _rowmajor_ld(group_bias), rows_per_group
    if residual is not None:
        p.residual, p.ldr = L.ptr(residual), _rowmajor_ld(residual)
    p.alpha = 1.0
    p.conv3x3 = 1
    p.img_n, p.img_h, p.img_w = n, h, w_
    L.check(L.load().hallo_b200_gemm(C.byref(p), L.current_stream()), "conv3x3")
    return out


def pack_conv3x3_weight(w: torch.Tensor) -> torch.Tensor:
    """torch conv weight [cout, cin, 3, 3] -> [cout, 9*cin] with k = (kh*3 + kw)*cin + c."""
    cout, cin, kh, kw = w.shape
    assert kh == 3 and kw == 3
    return w.permute(0, 2, 3, 1).reshape(cout, 9 * cin).contiguous()


def pack_geglu_weight(w: torch.Tensor, b: Optional[torch.Tensor]):
    """FeedForward.net[0].proj ([8C, C]: value rows then gate rows) -> rows interleaved (v0,g0,v1,g1..)."""
    n2, k = w.shape
    half = n2 // 2
    wi = torch.stack([w[:half], w[half:]], dim=1).reshape(n2, k).contiguous()
    bi = None if b is None else torch.stack([b[:half], b[half:]], dim=1).reshape(n2).contiguous()
    return wi, bi
