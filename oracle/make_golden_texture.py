"""Pins oracle/texture_oracle.py against the libraries the reference calls (Pillow convert("L"), cv2.Canny)
and writes tests/golden/texture_tiles.npz.  Run in the build container only (needs opencv-python + Pillow):

    python oracle/make_golden_texture.py
"""

from __future__ import annotations

import sys
from pathlib import Path

import cv2
import numpy as np
from PIL import Image

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from oracle import texture_oracle as to  # noqa: E402

OUT = Path(__file__).resolve().parent.parent / "tests" / "golden"


def reference_edges(rgb: np.ndarray) -> np.ndarray:
    """Exactly the reference's calls (tiling.py:283-285)."""
    gray = Image.fromarray(rgb).convert("L")
    return cv2.Canny(np.array(gray), 40, 100)


def make_tiles(rng: np.random.Generator, n: int, size: int) -> np.ndarray:
    """Tissue-like, flat, noisy and hard-edged tiles: smooth blobs + texture noise of varying strength."""
    tiles = np.empty((n, size, size, 3), dtype=np.uint8)
    yy, xx = np.mgrid[0:size, 0:size]
    for i in range(n):
        base = np.full((size, size, 3), rng.integers(120, 245), dtype=np.float64)
        for _ in range(rng.integers(0, 12)):
            cx, cy, r = rng.uniform(0, size, 2).tolist() + [rng.uniform(3, 40)]
            blob = np.exp(-((xx - cx) ** 2 + (yy - cy) ** 2) / (2 * r * r))
            base -= blob[..., None] * rng.uniform(10, 120, 3)
        if i % 5 == 0:      # hard rectangles: long straight edges, ties in the non-maximum suppression
            x0, y0 = rng.integers(0, size // 2, 2)
            base[y0:y0 + size // 3, x0:x0 + size // 2] -= rng.uniform(30, 90)
        base += rng.normal(0, rng.choice([0.0, 1.0, 4.0, 12.0]), base.shape)
        tiles[i] = np.clip(base, 0, 255).astype(np.uint8)
    return tiles


def main() -> None:
    rng = np.random.default_rng(20240229)
    checked = 0
    for size in (224, 64, 33):
        for tile in make_tiles(rng, 24, size):
            ref = reference_edges(tile)
            assert np.array_equal(np.array(Image.fromarray(tile).convert("L")), to.rgb_to_l(tile))
            assert np.array_equal(ref, to.canny(to.rgb_to_l(tile))), "oracle differs from cv2.Canny"
            checked += 1
    for _ in range(20):     # pure noise: every branch of the direction quantisation, dense hysteresis
        tile = rng.integers(0, 256, (48, 48, 3), dtype=np.uint8)
        assert np.array_equal(reference_edges(tile), to.canny(to.rgb_to_l(tile)))
        checked += 1
    tiles = make_tiles(np.random.default_rng(7), 16, 224)
    tiles[3] = rng.integers(0, 256, (224, 224, 3), dtype=np.uint8)
    tiles[4] = 255
    edges = np.stack([reference_edges(t) for t in tiles])
    scores = np.array([np.array(e).mean() / 255 for e in edges])
    small = make_tiles(np.random.default_rng(8), 6, 40)
    small_edges = np.stack([reference_edges(t) for t in small])
    OUT.mkdir(parents=True, exist_ok=True)
    np.savez_compressed(OUT / "texture_tiles.npz", tiles=tiles, edges=np.packbits(edges > 0, axis=-1),
                        scores=scores, small=small, small_edges=np.packbits(small_edges > 0, axis=-1))
    print(f"oracle == cv2.Canny on {checked} tiles; golden scores", np.round(scores, 4))


if __name__ == "__main__":
    main()
