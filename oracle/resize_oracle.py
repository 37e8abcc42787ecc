"""CPU oracle of the tile resampling (TEST INFRASTRUCTURE: only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline leg may import this; the product path is stamp_b200/resize.py + csrc/resize.cu).

Restates Pillow's 8-bit resampling, the code torchvision's ``transforms.Resize`` runs on PIL images at the
reference's call site src/stamp/preprocessing/extractor/gigapath.py:20-27.  Pillow is a third-party dependency of
the reference (pyproject: torchvision -> pillow; 12.2.0 in this image) and not part of /root/reference; its
algorithm (src/libImaging/Resample.c): ``precompute_coeffs`` (filter evaluated in double at
``(x + xmin - center + 0.5) / filterscale``, taps normalised to sum 1), ``normalize_coeffs_8bpc`` (coefficients to
int with 22 fractional bits, rounded half away from zero), ``ImagingResampleHorizontal_8bpc`` then
``ImagingResampleVertical_8bpc`` (int32 accumulators starting at 2^21, ``>> 22``, clip to 0..255; the intermediate
image is uint8).  Pinned: tests/test_resize_cpu.py compares it bit for bit with ``PIL.Image.resize`` and with the
torchvision transforms themselves (both importable here and on the GPU box).
"""

from __future__ import annotations

import math

import numpy as np

PRECISION_BITS = 32 - 8 - 2


def bicubic_filter(x: float, a: float = -0.5) -> float:
    x = -x if x < 0.0 else x
    if x < 1.0:
        return ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    if x < 2.0:
        return (((x - 5) * x + 8) * x - 4) * a
    return 0.0


def bilinear_filter(x: float) -> float:
    x = -x if x < 0.0 else x
    return 1.0 - x if x < 1.0 else 0.0


FILTERS = {"bicubic": (bicubic_filter, 2.0), "bilinear": (bilinear_filter, 1.0)}


def coefficients(in_size: int, out_size: int, filter: str = "bicubic"):
    fn, support = FILTERS[filter]
    scale = filterscale = in_size / out_size
    filterscale = max(filterscale, 1.0)
    support *= filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    kk = np.zeros((out_size, ksize), np.int64)
    bounds = np.zeros((out_size, 2), np.int64)
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        xmin = max(int(center - support + 0.5), 0)
        n = min(int(center + support + 0.5), in_size) - xmin
        w = [fn((x + xmin - center + 0.5) * (1.0 / filterscale)) for x in range(n)]
        ww = 0.0
        for v in w:
            ww += v
        w = [v / ww for v in w] if ww != 0.0 else w
        kk[xx, :n] = [int(-0.5 + v * (1 << PRECISION_BITS)) if v < 0 else int(0.5 + v * (1 << PRECISION_BITS)) for v in w]
        bounds[xx] = (xmin, n)
    return kk, bounds


def _pass(img: np.ndarray, out_size: int, axis: int, filter: str) -> np.ndarray:
    img = np.moveaxis(img, axis, 0)
    kk, bounds = coefficients(img.shape[0], out_size, filter)
    out = np.empty((out_size, *img.shape[1:]), np.uint8)
    for o in range(out_size):
        lo, n = bounds[o]
        acc = np.tensordot(kk[o, :n], img[lo:lo + n].astype(np.int64), axes=(0, 0)) + (1 << (PRECISION_BITS - 1))
        out[o] = np.clip(acc >> PRECISION_BITS, 0, 255)
    return np.moveaxis(out, 0, axis)


def resize(img: np.ndarray, out_h: int, out_w: int, filter: str = "bicubic") -> np.ndarray:
    """uint8 [H, W, C] -> uint8 [out_h, out_w, C]; horizontal pass first, like ImagingResample."""
    if img.shape[1] != out_w:
        img = _pass(img, out_w, 1, filter)
    if img.shape[0] != out_h:
        img = _pass(img, out_h, 0, filter)
    return img


def resize_center_crop(tiles: np.ndarray, size: int, crop: int, filter: str = "bicubic") -> np.ndarray:
    """Batch version of Resize(size) + CenterCrop(crop) for square-or-not uint8 tiles [B, H, W, 3]."""
    out = []
    for t in tiles:
        h, w = t.shape[:2]
        rh, rw = (int(size * h / w), size) if w <= h else (size, int(size * w / h))
        r = resize(t, rh, rw, filter)
        top, left = int(round((rh - crop) / 2.0)), int(round((rw - crop) / 2.0))
        out.append(r[top:top + crop, left:left + crop])
    return np.stack(out)
