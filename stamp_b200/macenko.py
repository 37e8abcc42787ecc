"""Macenko stain normalisation of a tile batch on the GPU (one call into ``stamp_macenko_u8``).

The reference snapshot has no Macenko stage (README.md:35 only); BASELINE.json's north_star names
it as the first stage of the per-slide hot path.  Algorithm: SURVEY.md 8c.
"""

from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import torch
from torch import Tensor

from . import _lib


def _bind() -> C.CDLL:
    lib = _lib.load()
    if not getattr(lib, "_mac_bound", False):
        lib.stamp_macenko_workspace_bytes.restype = C.c_size_t
        lib.stamp_macenko_workspace_bytes.argtypes = [C.c_int, C.c_int]
        lib.stamp_macenko_u8.restype = C.c_int
        lib.stamp_macenko_u8.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                         C.c_float, C.c_float, C.c_float, C.c_void_p, C.c_void_p,
                                         C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
        lib._mac_bound = True
    return lib


@dataclass
class MacenkoFit:
    stain_matrix: Tensor   # [G, 3, 2] columns (haematoxylin, eosin)
    max_conc: Tensor       # [G, 2] 99th percentile stain concentrations
    valid: Tensor          # [G] bool: False = too little tissue, tiles passed through


def macenko_normalize(tiles: Tensor, *, tiles_per_fit: int | None = None, Io: float = 240.0,
                      alpha: float = 1.0, beta: float = 0.15, out: Tensor | None = None,
                      return_fit: bool = False):
    """uint8 ``[B, H, W, 3]`` CUDA tiles -> stain-normalised uint8 tiles (same shape)."""
    if not tiles.is_cuda:
        raise RuntimeError("macenko_normalize runs on a CUDA device only (no CPU fallback)")
    if tiles.dtype != torch.uint8 or tiles.dim() != 4 or tiles.shape[-1] != 3 or not tiles.is_contiguous():
        raise TypeError("tiles must be a contiguous uint8 [B,H,W,3] tensor")
    B, H, W, _ = tiles.shape
    if B == 0:
        return (tiles.clone(), None) if return_fit else tiles.clone()
    lib = _bind()
    tpf = int(tiles_per_fit) if tiles_per_fit else 0
    G = 1 if tpf <= 0 else (B + tpf - 1) // tpf
    if out is None:
        out = torch.empty_like(tiles)
    he = torch.empty((G, 3, 2), dtype=torch.float32, device=tiles.device)
    maxc = torch.empty((G, 2), dtype=torch.float32, device=tiles.device)
    valid = torch.empty((G,), dtype=torch.int32, device=tiles.device)
    ws = torch.empty(lib.stamp_macenko_workspace_bytes(B, tpf), dtype=torch.uint8, device=tiles.device)
    code = lib.stamp_macenko_u8(tiles.data_ptr(), out.data_ptr(), B, H, W, tpf, Io, alpha, beta,
                                he.data_ptr(), maxc.data_ptr(), valid.data_ptr(), ws.data_ptr(),
                                ws.numel(), torch.cuda.current_stream().cuda_stream)
    _lib.check(code, "stamp_macenko_u8")
    if return_fit:
        return out, MacenkoFit(he, maxc, valid.bool())
    return out
