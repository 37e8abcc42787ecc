"""Tile resampling on the GPU, bit-exact with Pillow (and therefore with torchvision's ``transforms.Resize`` on
PIL images, which is what the reference's extractor transforms run).

reference call site: src/stamp/preprocessing/extractor/gigapath.py:20-27 -- ``Resize(256, BICUBIC)`` +
``CenterCrop(224)`` ahead of ``ToTensor`` / ``Normalize``.  Pillow (third-party, 12.2.0 in this image; algorithm
unchanged since 3.x) resamples 8-bit images in ``src/libImaging/Resample.c``: ``precompute_coeffs`` evaluates the
filter in double precision and normalises every output pixel's taps, ``normalize_coeffs_8bpc`` turns them into
22-bit fixed point, ``ImagingResampleHorizontal_8bpc`` / ``...Vertical_8bpc`` accumulate in int32 from one half
and clip.  The tables are computed here with the same double operations in the same order; the passes run in
``stamp_resize_u8`` (csrc/resize.cu).
"""

from __future__ import annotations

import ctypes as C
import functools
import math

import numpy as np
import torch
from torch import Tensor

from . import _lib

PRECISION_BITS = 32 - 8 - 2
ROWS_PER_STRIP = 16


def _bicubic(x: float) -> float:
    a = -0.5
    if x < 0.0:
        x = -x
    if x < 1.0:
        return ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    if x < 2.0:
        return (((x - 5) * x + 8) * x - 4) * a
    return 0.0


def _bilinear(x: float) -> float:
    if x < 0.0:
        x = -x
    return 1.0 - x if x < 1.0 else 0.0


_FILTERS = {"bicubic": (_bicubic, 2.0), "bilinear": (_bilinear, 1.0)}


@functools.lru_cache(maxsize=64)
def pillow_resample_tables(in_size: int, out_size: int, filter: str = "bicubic") -> tuple[np.ndarray, np.ndarray]:
    """-> (coefficients int32 [out_size, ksize], bounds int32 [out_size, 2] = (first input index, tap count))."""
    fn, support = _FILTERS[filter]
    scale = filterscale = in_size / out_size
    if filterscale < 1.0:
        filterscale = 1.0
    support = support * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    coef = np.zeros((out_size, ksize), dtype=np.int32)
    bounds = np.zeros((out_size, 2), dtype=np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        xmin = max(int(center - support + 0.5), 0)
        xmax = min(int(center + support + 0.5), in_size) - xmin
        w = [fn((x + xmin - center + 0.5) * ss) for x in range(xmax)]
        ww = 0.0
        for v in w:
            ww += v
        for x, v in enumerate(w):
            if ww != 0.0:
                v = v / ww
            coef[xx, x] = int(-0.5 + v * (1 << PRECISION_BITS)) if v < 0 else int(0.5 + v * (1 << PRECISION_BITS))
        bounds[xx] = (xmin, xmax)
    return coef, bounds


def resized_shape(h: int, w: int, size: int | tuple[int, int]) -> tuple[int, int]:
    """torchvision ``Resize``: an int matches the smaller edge and keeps the aspect ratio (truncating)."""
    if isinstance(size, int):
        if w <= h:
            return int(size * h / w), size
        return size, int(size * w / h)
    return int(size[0]), int(size[1])


def _bind() -> C.CDLL:
    lib = _lib.load()
    if not getattr(lib, "_resize_bound", False):
        lib.stamp_resize_u8_smem_bytes.restype = C.c_size_t
        lib.stamp_resize_u8_smem_bytes.argtypes = [C.c_int] * 4
        lib.stamp_resize_u8.restype = C.c_int
        lib.stamp_resize_u8.argtypes = ([C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p] + [C.c_int] * 4 +
                                        [C.c_void_p, C.c_void_p, C.c_int] * 2 + [C.c_int, C.c_int, C.c_void_p])
        lib._resize_bound = True
    return lib


@functools.lru_cache(maxsize=32)
def _device_tables(in_size: int, out_size: int, filter: str, device: torch.device) -> tuple[Tensor, Tensor, int]:
    coef, bounds = pillow_resample_tables(in_size, out_size, filter)
    return (torch.from_numpy(coef).to(device).contiguous(), torch.from_numpy(bounds).to(device).contiguous(),
            coef.shape[1])


def resize_center_crop(tiles: Tensor, size: int | tuple[int, int], crop: int | tuple[int, int] | None = None,
                       filter: str = "bicubic") -> Tensor:
    """uint8 ``[B, H, W, 3]`` CUDA tiles -> ``Resize(size, filter)`` + ``CenterCrop(crop)`` as uint8 ``[B, h, w, 3]``;
    identical to running the two torchvision transforms on the PIL image of every tile."""
    if not tiles.is_cuda:
        raise RuntimeError("resize_center_crop runs on a CUDA device only (no CPU fallback)")
    if tiles.dtype != torch.uint8 or tiles.dim() != 4 or tiles.shape[-1] != 3 or not tiles.is_contiguous():
        raise TypeError("tiles must be a contiguous uint8 [B,H,W,3] tensor")
    B, H, W, _ = tiles.shape
    rh, rw = resized_shape(H, W, size)
    if crop is None:
        ch, cw = rh, rw
    else:
        ch, cw = (crop, crop) if isinstance(crop, int) else (int(crop[0]), int(crop[1]))
    if ch > rh or cw > rw:
        raise ValueError(f"crop {ch}x{cw} larger than the resized tile {rh}x{rw} (padding crops are not supported)")
    cy, cx = int(round((rh - ch) / 2.0)), int(round((rw - cw) / 2.0))
    kx, bx, ksx = _device_tables(W, rw, filter, tiles.device)
    ky, by, ksy = _device_tables(H, rh, filter, tiles.device)
    _, by_host = pillow_resample_tables(H, rh, filter)
    max_in = 0
    for r0 in range(0, ch, ROWS_PER_STRIP):
        last = cy + min(r0 + ROWS_PER_STRIP, ch) - 1
        max_in = max(max_in, int(by_host[last, 0] + by_host[last, 1] - by_host[cy + r0, 0]))
    out = torch.empty((B, ch, cw, 3), dtype=torch.uint8, device=tiles.device)
    lib = _bind()
    stream = torch.cuda.current_stream().cuda_stream
    for s in range(0, B, 65535):
        b = min(65535, B - s)
        code = lib.stamp_resize_u8(tiles[s:s + b].data_ptr(), b, H, W, out[s:s + b].data_ptr(), ch, cw, cy, cx,
                                   kx.data_ptr(), bx.data_ptr(), ksx, ky.data_ptr(), by.data_ptr(), ksy,
                                   ROWS_PER_STRIP, max_in, stream)
        _lib.check(code, "stamp_resize_u8")
    return out
