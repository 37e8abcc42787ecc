"""stamp_tile_texture_u8 on the GPU: Canny edge maps bit-exact with cv2.Canny on Pillow's grayscale
(golden fixtures from the real libraries) and with the oracle on ragged tile sizes and pure noise."""

import numpy as np
import pytest
import torch

from oracle import texture_oracle as to
from test_texture_cpu import load_texture_golden

pytestmark = pytest.mark.gpu


def test_edges_and_scores_bit_exact_vs_cv2_golden(cuda_device):
    from stamp_b200.tiling import canny_edge_counts, edge_scores, has_enough_texture

    tiles, edges, scores, small, small_edges = load_texture_golden()
    t = torch.from_numpy(tiles).to(cuda_device)
    counts, e = canny_edge_counts(t, return_edges=True)
    assert np.array_equal(e.cpu().numpy() > 0, edges)
    assert set(np.unique(e.cpu().numpy()).tolist()) <= {0, 255}
    assert np.array_equal(counts.cpu().numpy(), edges.reshape(len(tiles), -1).sum(1))
    assert np.array_equal(edge_scores(t).cpu().numpy(), scores)          # identical doubles
    for cutoff in (0.0, 0.005, 0.02, 0.2):
        assert np.array_equal(has_enough_texture(t, cutoff).cpu().numpy(), scores >= cutoff)
    s = torch.from_numpy(small).to(cuda_device)
    _, es = canny_edge_counts(s, return_edges=True)
    assert np.array_equal(es.cpu().numpy() > 0, small_edges)


@pytest.mark.parametrize("h,w", [(33, 47), (224, 224), (5, 3), (240, 200), (1, 1)])
def test_ragged_sizes_and_noise_match_oracle(cuda_device, h, w):
    from stamp_b200.tiling import canny_edge_counts

    rng = np.random.default_rng(h * 1000 + w)
    tiles = rng.integers(0, 256, (3, h, w, 3), dtype=np.uint8)
    tiles[1] = (tiles[1].astype(np.int32) // 8 + 100).astype(np.uint8)      # low contrast: mostly weak pixels
    yy, xx = np.mgrid[0:h, 0:w]
    tiles[2] = ((np.sin(xx / 3.0) + np.cos(yy / 5.0)) * 50 + 128).astype(np.uint8)[..., None]   # long thin edges
    counts, e = canny_edge_counts(torch.from_numpy(tiles).to(cuda_device), return_edges=True)
    for i in range(3):
        ref = to.canny(to.rgb_to_l(tiles[i]))
        assert np.array_equal(e[i].cpu().numpy(), ref), i
        assert int(counts[i]) == int((ref > 0).sum())


def test_texture_errors(cuda_device):
    from stamp_b200._lib import StampB200Error
    from stamp_b200.tiling import canny_edge_counts

    with pytest.raises(RuntimeError):
        canny_edge_counts(torch.zeros(1, 8, 8, 3, dtype=torch.uint8))
    with pytest.raises(TypeError):
        canny_edge_counts(torch.zeros(1, 8, 8, 3, device=cuda_device))
    with pytest.raises(StampB200Error):
        canny_edge_counts(torch.zeros(1, 512, 512, 3, dtype=torch.uint8, device=cuda_device))   # > one SM's smem
    assert canny_edge_counts(torch.zeros(0, 8, 8, 3, dtype=torch.uint8, device=cuda_device)).numel() == 0
