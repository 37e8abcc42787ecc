"""ORACLE (test infrastructure, not product code) -- CPU restatement of STAMP's tissue-texture filter.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline leg may import this module.

Follows ``_has_enough_texture`` (src/stamp/preprocessing/tiling.py:279-291):

    tile_grayscale = tile.convert("L")                        # PIL, ITU-R 601-2 luma, integer arithmetic
    edges = cv2.Canny(np.array(tile_grayscale), 40, 100)      # OpenCV, aperture 3, L1 gradient
    edge_score = np.array(edges).mean() / 255
    return edge_score >= cutoff

The arithmetic lives in two third-party libraries (Pillow ``ImagingConvert`` rgb2l, OpenCV ``cv::Canny``;
opencv-python 4.13 / Pillow 12.2 in the build container, where ``oracle/make_golden_texture.py`` pins this
restatement bit-for-bit against both on seeded tiles and writes tests/golden/texture_*.npz).

Canny as OpenCV computes it for 8-bit input, ``L2gradient=False``:
  1. dx, dy: 3x3 Sobel, 16-bit, BORDER_REPLICATE;  mag = |dx| + |dy|, zero outside the image;
  2. pixels with mag > low are kept if they are a local maximum along the gradient direction quantised with
     the fixed-point tangents of 22.5 / 67.5 degrees (TG22 = round(tan(22.5 deg) * 2^15) = 13573):
       |dy| * 2^15 <  |dx| * TG22                   : mag >  left       and mag >= right
       |dy| * 2^15 >  |dx| * TG22 + |dx| * 2^16     : mag >  above      and mag >= below
       otherwise (diagonal, s = sign(dx * dy))       : mag >  (above, x - s) and mag > (below, x + s)
     kept pixels with mag > high are strong, the others weak;
  3. hysteresis: weak pixels 8-connected (through weak pixels) to a strong one become edges.
"""

from __future__ import annotations

import numpy as np

TG22 = 13573
CANNY_SHIFT = 15


def rgb_to_l(rgb: np.ndarray) -> np.ndarray:
    """Pillow ``Image.convert("L")`` on uint8 RGB [..., 3]: (R*19595 + G*38470 + B*7471 + 0x8000) >> 16."""
    r, g, b = (rgb[..., i].astype(np.int64) for i in range(3))
    return ((r * 19595 + g * 38470 + b * 7471 + 0x8000) >> 16).astype(np.uint8)


def sobel3(gray: np.ndarray) -> tuple[np.ndarray, np.ndarray]:
    p = np.pad(gray.astype(np.int32), 1, mode="edge")
    H, W = gray.shape
    s = lambda dy, dx: p[1 + dy:1 + dy + H, 1 + dx:1 + dx + W]
    gx = (s(-1, 1) + 2 * s(0, 1) + s(1, 1)) - (s(-1, -1) + 2 * s(0, -1) + s(1, -1))
    gy = (s(1, -1) + 2 * s(1, 0) + s(1, 1)) - (s(-1, -1) + 2 * s(-1, 0) + s(-1, 1))
    return gx, gy


def canny(gray: np.ndarray, low: int = 40, high: int = 100) -> np.ndarray:
    """``cv2.Canny(gray, low, high)`` for uint8 ``gray`` [H, W]; returns uint8 {0, 255}."""
    H, W = gray.shape
    gx, gy = sobel3(gray)
    mag = np.zeros((H + 2, W + 2), dtype=np.int32)
    mag[1:-1, 1:-1] = np.abs(gx) + np.abs(gy)
    m = mag[1:-1, 1:-1]
    nb = lambda dy, dx: mag[1 + dy:1 + dy + H, 1 + dx:1 + dx + W]
    x = np.abs(gx).astype(np.int64)
    y = np.abs(gy).astype(np.int64) << CANNY_SHIFT
    tg22x = x * TG22
    tg67x = tg22x + (x << (CANNY_SHIFT + 1))
    horiz = y < tg22x
    vert = ~horiz & (y > tg67x)
    diag = ~horiz & ~vert
    s_neg = (gx ^ gy) < 0                      # s = -1 where the signs differ
    keep_h = (m > nb(0, -1)) & (m >= nb(0, 1))
    keep_v = (m > nb(-1, 0)) & (m >= nb(1, 0))
    # s = +1: compare (above, x-1) and (below, x+1);  s = -1: (above, x+1) and (below, x-1)
    keep_d = np.where(s_neg, (m > nb(-1, 1)) & (m > nb(1, -1)), (m > nb(-1, -1)) & (m > nb(1, 1)))
    keep = (m > low) & ((horiz & keep_h) | (vert & keep_v) | (diag & keep_d))
    strong = keep & (m > high)
    # hysteresis: flood fill from the strong pixels through kept pixels, 8-connected
    edge = np.zeros((H + 2, W + 2), dtype=bool)
    cand = np.zeros((H + 2, W + 2), dtype=bool)
    edge[1:-1, 1:-1] = strong
    cand[1:-1, 1:-1] = keep
    while True:
        grow = np.zeros_like(edge)
        for dy in (-1, 0, 1):
            for dx in (-1, 0, 1):
                if dy or dx:
                    grow[1:-1, 1:-1] |= edge[1 + dy:1 + dy + H, 1 + dx:1 + dx + W]
        new = grow & cand & ~edge
        if not new.any():
            break
        edge |= new
    return (edge[1:-1, 1:-1].astype(np.uint8)) * 255


def edge_score(rgb_tile: np.ndarray, low: int = 40, high: int = 100) -> float:
    """``_has_enough_texture``'s score for one uint8 RGB tile [H, W, 3]."""
    return float(np.array(canny(rgb_to_l(rgb_tile), low, high)).mean() / 255)


def has_enough_texture(rgb_tile: np.ndarray, cutoff: float) -> bool:
    return bool(edge_score(rgb_tile) >= cutoff)
