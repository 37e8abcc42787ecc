"""The texture-filter oracle against golden edge maps produced by the libraries the reference calls
(Pillow convert("L") + cv2.Canny, written by oracle/make_golden_texture.py in the build container)."""

from pathlib import Path

import numpy as np

from oracle import texture_oracle as to

GOLD = Path(__file__).resolve().parent / "golden" / "texture_tiles.npz"


def load_texture_golden():
    z = np.load(GOLD)
    tiles, small = z["tiles"], z["small"]
    edges = np.unpackbits(z["edges"], axis=-1)[..., : tiles.shape[2]].astype(bool)
    small_edges = np.unpackbits(z["small_edges"], axis=-1)[..., : small.shape[2]].astype(bool)
    return tiles, edges, z["scores"], small, small_edges


def test_texture_oracle_matches_cv2_golden():
    tiles, edges, scores, small, small_edges = load_texture_golden()
    for i in (0, 1, 3, 4, 9):          # a blob tile, a textured one, pure noise, blank, nearly blank
        got = to.canny(to.rgb_to_l(tiles[i])) > 0
        assert np.array_equal(got, edges[i]), i
        assert to.edge_score(tiles[i]) == scores[i]
    for t, e in zip(small, small_edges):
        assert np.array_equal(to.canny(to.rgb_to_l(t)) > 0, e)
    assert to.has_enough_texture(tiles[3], 0.02) and not to.has_enough_texture(tiles[4], 0.02)
