"""Cached-tile JPEG decode split between host and GPU, bit-exact with Pillow.

reference call site: ``_tiles_from_cache_file`` (src/stamp/preprocessing/tiling.py:380-406) opens every cached
tile with ``Image.open`` + ``load`` on the CPU.  Here the host only does what is sequential by construction
(marker parsing + Huffman decoding into quantised DCT coefficients, ``stamp_jpeg_entropy_decode``, one call per
tile, GIL released, so a thread pool scales); dequantisation, inverse DCT, chroma up-sampling and colour conversion
(two thirds of libjpeg's decode time) run in ``stamp_jpeg_decode_coefs_u8`` on the GPU and the tiles are born in
HBM, where the tissue filter / Macenko / the tile encoder take them from.  Same bytes over PCIe as decoded RGB
(int16 coefficients of a 4:2:0 tile = 1.5 x H x W x 2 B).
"""

from __future__ import annotations

import ctypes as C
from collections.abc import Sequence
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import torch
from torch import Tensor

from . import _lib


class StampJpegInfo(C.Structure):
    _fields_ = [("width", C.c_int), ("height", C.c_int), ("n_comp", C.c_int), ("h", C.c_int * 3), ("v", C.c_int * 3),
                ("mcus_x", C.c_int), ("mcus_y", C.c_int), ("quant", (C.c_uint16 * 64) * 3)]


def _bind() -> C.CDLL:
    lib = _lib.load()
    if not getattr(lib, "_jpeg_bound", False):
        lib.stamp_jpeg_read_header.restype = C.c_int
        lib.stamp_jpeg_read_header.argtypes = [C.c_char_p, C.c_size_t, C.POINTER(StampJpegInfo)]
        lib.stamp_jpeg_coef_count.restype = C.c_size_t
        lib.stamp_jpeg_coef_count.argtypes = [C.POINTER(StampJpegInfo)]
        lib.stamp_jpeg_entropy_decode.restype = C.c_int
        lib.stamp_jpeg_entropy_decode.argtypes = [C.c_char_p, C.c_size_t, C.POINTER(StampJpegInfo), C.c_void_p, C.c_void_p]
        lib.stamp_jpeg_workspace_bytes.restype = C.c_size_t
        lib.stamp_jpeg_workspace_bytes.argtypes = [C.POINTER(StampJpegInfo), C.c_int]
        lib.stamp_jpeg_decode_coefs_u8.restype = C.c_int
        lib.stamp_jpeg_decode_coefs_u8.argtypes = [C.POINTER(StampJpegInfo), C.c_void_p, C.c_void_p, C.c_int, C.c_void_p,
                                                   C.c_void_p, C.c_size_t, C.c_void_p]
        lib._jpeg_bound = True
    return lib


def read_header(blob: bytes) -> StampJpegInfo:
    info = StampJpegInfo()
    _lib.check(_bind().stamp_jpeg_read_header(blob, len(blob), C.byref(info)), "stamp_jpeg_read_header")
    return info


def coef_count(info: StampJpegInfo) -> int:
    return int(_bind().stamp_jpeg_coef_count(C.byref(info)))


def entropy_decode(blobs: Sequence[bytes], *, max_workers: int = 8, pin: bool = False,
                   out: tuple[Tensor, Tensor] | None = None) -> tuple[StampJpegInfo, Tensor, Tensor]:
    """Huffman-decode a batch of same-geometry JPEG tiles on the host -> (geometry, coefficients int16
    [B, coef_count], quantisation tables uint16-as-int16 [B, 3, 64]).  No GPU involved.  ``out``: reusable (pinned)
    staging tensors of at least that size -- pinning 150 KB per tile on every call costs more than the decode."""
    if len(blobs) == 0:
        raise ValueError("empty batch")
    lib = _bind()
    info = read_header(blobs[0])
    n = coef_count(info)
    if out is None:
        coef = torch.empty((len(blobs), n), dtype=torch.int16, pin_memory=pin)
        quant = torch.empty((len(blobs), 3, 64), dtype=torch.int16, pin_memory=pin)
    else:
        coef, quant = out[0][: len(blobs)], out[1][: len(blobs)]
        if coef.shape != (len(blobs), n) or quant.shape != (len(blobs), 3, 64) or coef.dtype != torch.int16 \
                or quant.dtype != torch.int16 or not coef.is_contiguous() or not quant.is_contiguous() or coef.is_cuda:
            raise ValueError(f"staging tensors must be contiguous int16 host tensors [>= {len(blobs)}, {n}] / [.., 3, 64]")
    cptr, qptr = coef.data_ptr(), quant.data_ptr()

    def one(i: int) -> int:
        return lib.stamp_jpeg_entropy_decode(blobs[i], len(blobs[i]), C.byref(info), cptr + i * n * 2, qptr + i * 384)

    if max_workers > 1 and len(blobs) > 1:
        with ThreadPoolExecutor(max_workers=max_workers) as ex:
            codes = list(ex.map(one, range(len(blobs))))
    else:
        codes = [one(i) for i in range(len(blobs))]
    for i, code in enumerate(codes):
        _lib.check(code, f"stamp_jpeg_entropy_decode (tile {i})")
    return info, coef, quant


def decode_coefficients(info: StampJpegInfo, coef: Tensor, quant: Tensor, out: Tensor | None = None) -> Tensor:
    """Coefficients on the device -> uint8 ``[B, H, W, 3]`` RGB tiles (``stamp_jpeg_decode_coefs_u8``)."""
    if not coef.is_cuda or not quant.is_cuda:
        raise RuntimeError("decode_coefficients runs on a CUDA device only (no CPU fallback)")
    lib = _bind()
    B = coef.shape[0]
    if out is None:
        out = torch.empty((B, info.height, info.width, 3), dtype=torch.uint8, device=coef.device)
    ws = torch.empty(lib.stamp_jpeg_workspace_bytes(C.byref(info), B), dtype=torch.uint8, device=coef.device)
    code = lib.stamp_jpeg_decode_coefs_u8(C.byref(info), coef.data_ptr(), quant.data_ptr(), B, out.data_ptr(),
                                          ws.data_ptr(), ws.numel(), torch.cuda.current_stream().cuda_stream)
    _lib.check(code, "stamp_jpeg_decode_coefs_u8")
    return out


def decode_jpeg_tiles(blobs: Sequence[bytes], device: torch.device | str = "cuda", *, max_workers: int = 8) -> Tensor:
    """JPEG files of one geometry (the tiles of a cache) -> uint8 ``[B, H, W, 3]`` on ``device``; identical to
    ``np.asarray(Image.open(blob).convert("RGB"))`` for every tile."""
    device = torch.device(device)
    if device.type != "cuda":
        raise RuntimeError("decode_jpeg_tiles targets a CUDA device (no CPU fallback; use Pillow for host decoding)")
    info, coef, quant = entropy_decode(blobs, max_workers=max_workers, pin=True)
    return decode_coefficients(info, coef.to(device, non_blocking=True), quant.to(device, non_blocking=True))
