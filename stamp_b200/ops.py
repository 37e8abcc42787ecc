"""Tensor-level wrappers over the C ABI: dtype / contiguity / device checks, output allocation
through torch's caching allocator, current-stream plumbing.  No arithmetic happens in Python.
"""

from __future__ import annotations

import ctypes as C
import math

import torch
from torch import Tensor

from . import _lib

ACT_NONE, ACT_GELU, ACT_RELU = 0, 1, 2
ST_16, ST_32, ST_RESID32, ST_SWIGLU16, ST_GATED16 = 0, 1, 2, 3, 4


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptr(t: Tensor | None) -> int | None:
    return None if t is None else t.data_ptr()


def _require_cuda(*ts: Tensor | None) -> None:
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("stamp_b200 kernels need CUDA tensors (no CPU fallback exists)")


def _f32(t: Tensor | None, name: str) -> Tensor | None:
    if t is None:
        return None
    if t.dtype != torch.float32 or not t.is_contiguous():
        raise TypeError(f"{name} must be a contiguous float32 tensor")
    return t


def gemm_tn(
    a: Tensor,
    w: Tensor,
    *,
    out: Tensor,
    bias: Tensor | None = None,
    gamma: Tensor | None = None,
    act: int = ACT_NONE,
    store: int = ST_16,
    table: Tensor | None = None,
    gin: int = 0,
    gout: int = 0,
    goff: int = 0,
) -> Tensor:
    """out = epilogue(a @ w.T).  a [M,K], w [N,K] fp16/bf16 (last dim contiguous)."""
    _require_cuda(a, w, out, bias, gamma, table)
    if a.dtype != w.dtype or a.dtype not in (torch.float16, torch.bfloat16, torch.float32):
        raise TypeError("a and w must share one dtype: float16, bfloat16 or float32 (consumed as TF32)")
    if a.dim() != 2 or w.dim() != 2 or a.shape[1] != w.shape[1]:
        raise ValueError(f"shape mismatch: a {tuple(a.shape)} w {tuple(w.shape)}")
    if a.stride(1) != 1 or w.stride(1) != 1 or out.stride(-1) != 1:
        raise ValueError("innermost dimensions must be contiguous")
    M, K = a.shape
    N = w.shape[0]
    want = torch.float32 if store in (ST_32, ST_RESID32) else (torch.float16 if a.dtype == torch.float32 else a.dtype)
    if out.dtype != want:
        raise TypeError(f"out must be {want} for store mode {store}")
    _f32(bias, "bias"), _f32(gamma, "gamma")
    ldt = 0
    if table is not None:
        _f32(table, "table")
        ldt = table.stride(0)
    code = _lib.load().stamp_gemm_tn(
        a.data_ptr(), a.stride(0), w.data_ptr(), w.stride(0), out.data_ptr(), out.stride(-2),
        M, N, K, _ptr(bias), _ptr(gamma), act, store,
        {torch.float16: 0, torch.bfloat16: 1, torch.float32: 2}[a.dtype],
        _ptr(table), ldt, gin, gout, goff, _stream(),
    )
    _lib.check(code, "stamp_gemm_tn")
    return out


def layernorm(x: Tensor, weight: Tensor, bias: Tensor, eps: float, out_dtype: torch.dtype,
              out: Tensor | None = None) -> Tensor:
    """Row-wise LayerNorm of an fp32 [rows, cols] matrix into fp16 / bf16 / fp32."""
    _require_cuda(x, weight, bias)
    if x.dtype != torch.float32 or x.dim() != 2 or x.stride(1) != 1:
        raise TypeError("x must be a float32 matrix with contiguous rows")
    _f32(weight, "weight"), _f32(bias, "bias")
    kind = {torch.float16: 0, torch.bfloat16: 1, torch.float32: 2}[out_dtype]
    if out is None:
        out = torch.empty(x.shape, dtype=out_dtype, device=x.device)
    code = _lib.load().stamp_layernorm(x.data_ptr(), x.stride(0), weight.data_ptr(), bias.data_ptr(),
                                       out.data_ptr(), None, out.stride(0), x.shape[0], x.shape[1],
                                       float(eps), kind, _stream())
    _lib.check(code, "stamp_layernorm")
    return out


def fill_rows(x: Tensor, groups: int, rows_per_group: int, row_off: int, src: Tensor,
              add: Tensor | None = None) -> None:
    """x[g*rows_per_group + row_off + r] = src[r] (+ add[r]) for every group g."""
    _require_cuda(x, src, add)
    _f32(src, "src"), _f32(add, "add")
    if x.dtype != torch.float32 or x.dim() != 2:
        raise TypeError("x must be a float32 matrix")
    code = _lib.load().stamp_fill_rows(x.data_ptr(), x.stride(0), groups, rows_per_group, row_off,
                                       src.data_ptr(), src.stride(0), _ptr(add),
                                       add.stride(0) if add is not None else 0, src.shape[0],
                                       src.shape[1], _stream())
    _lib.check(code, "stamp_fill_rows")


def tiles_to_patches(tiles: Tensor, patch: int, mean, std, dtype=torch.float16,
                     kpad: int | None = None, out: Tensor | None = None) -> Tensor:
    """uint8 [B,H,W,3] tiles -> normalised patch matrix [B*(H/P)^2, kpad]."""
    _require_cuda(tiles)
    if tiles.dtype != torch.uint8 or tiles.dim() != 4 or tiles.shape[-1] != 3 or not tiles.is_contiguous():
        raise TypeError("tiles must be a contiguous uint8 [B,H,W,3] tensor")
    B, H, W, _ = tiles.shape
    if H != W or H % patch:
        raise ValueError("tiles must be square with a side divisible by the patch size")
    k = 3 * patch * patch
    kpad = kpad or (k + 7) // 8 * 8
    if out is None:
        out = torch.empty((B * (H // patch) ** 2, kpad), dtype=dtype, device=tiles.device)
    m = (C.c_float * 3)(*[float(v) for v in mean])
    s = (C.c_float * 3)(*[float(v) for v in std])
    code = _lib.load().stamp_tiles_to_patches(tiles.data_ptr(), out.data_ptr(), B, H, patch, kpad,
                                              m, s, int(dtype == torch.bfloat16), _stream())
    _lib.check(code, "stamp_tiles_to_patches")
    return out


def attention(qkv: Tensor, n_heads: int, *, out: Tensor | None = None, coords: Tensor | None = None,
              slope: Tensor | None = None, mask: Tensor | None = None, mask_mode: int = 1,
              scale: float | None = None) -> Tensor:
    """Fused attention over a packed fp16 [B, S, 3, H, hd] (= [B, S, 3*H*hd]) projection.

    Plain attention returns fp16; the ALiBi variant returns fp32 values rounded to TF32 (the
    reference's distance term is unscaled and overflows fp16)."""
    _require_cuda(qkv, coords, slope, mask)
    if qkv.dtype != torch.float16 or qkv.dim() != 3 or not qkv.is_contiguous():
        raise TypeError("qkv must be a contiguous float16 [B, S, 3*D] tensor")
    B, S, D3 = qkv.shape
    D = D3 // 3
    hd = D // n_heads
    out_dtype = torch.float32 if coords is not None else torch.float16
    if out is None:
        out = torch.empty((B, S, D), dtype=out_dtype, device=qkv.device)
    if out.dtype != out_dtype or not out.is_contiguous():
        raise TypeError(f"out must be a contiguous {out_dtype} tensor")
    dscale = None
    lib = _lib.load()
    if coords is not None:
        _f32(coords, "coords"), _f32(slope, "slope")
        if coords.shape != (B, S, 2) or slope is None or slope.numel() != n_heads:
            raise ValueError("coords must be [B,S,2] and slope [H]")
        dscale = torch.empty((B, 2), dtype=torch.float32, device=qkv.device)
        _lib.check(lib.stamp_alibi_dist_scale(coords.data_ptr(), slope.data_ptr(), B, S, n_heads,
                                              dscale.data_ptr(), _stream()), "stamp_alibi_dist_scale")
    if mask is not None:
        if mask.dtype == torch.bool:
            mask = mask.to(torch.uint8)
        if mask.dtype != torch.uint8 or mask.shape != (B, S) or not mask.is_contiguous():
            raise TypeError("mask must be a contiguous bool/uint8 [B,S] tensor")
    esz = 1  # strides are in elements
    base = qkv.data_ptr()
    code = lib.stamp_attention_fwd(
        base, base + 2 * D, base + 4 * D, D3 * esz, S * D3 * esz, out.data_ptr(), D, S * D,
        int(out_dtype == torch.float32), B, S, n_heads, hd, float(scale if scale is not None else 1.0 / math.sqrt(hd)),
        _ptr(coords), _ptr(slope), _ptr(dscale), _ptr(mask), mask_mode, _stream(),
    )
    _lib.check(code, "stamp_attention_fwd")
    return out
