"""Tile resampling on the GPU (csrc/resize.cu through stamp_b200.resize) against the Pillow-pinned oracle: bit-exact,
and the GigaPath tile encoder (Resize(256, BICUBIC) + CenterCrop(224) + ViT-g/16 blocks) against the fp32 oracle."""

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _tiles(n, h, w, seed):
    rng = np.random.default_rng(seed)
    t = rng.integers(0, 256, (n, h, w, 3), dtype=np.uint8)
    t[0] = np.where(rng.random((h, w, 3)) < 0.5, 0, 255)           # maximal overshoot of the bicubic lobes
    t[-1] = 255
    return t


@pytest.mark.parametrize("h,w,size,crop,filter", [
    (224, 224, 256, 224, "bicubic"),       # the GigaPath transform
    (224, 224, 112, None, "bicubic"),      # down-sampling: 9 taps
    (200, 260, 256, 224, "bicubic"),       # non-square tile, smaller edge -> 256
    (97, 131, 64, 33, "bilinear"),         # output row of 99 bytes: the byte-wise store path
    (224, 224, 224, 224, "bicubic"),       # identity sizes still go through both passes' rounding
])
def test_resize_matches_pillow_oracle(cuda_device, h, w, size, crop, filter):
    from oracle import resize_oracle as ro
    from stamp_b200.resize import resize_center_crop, resized_shape

    tiles = _tiles(5, h, w, seed=h + w)
    got = resize_center_crop(torch.from_numpy(tiles).to(cuda_device), size, crop, filter).cpu().numpy()
    rh, rw = resized_shape(h, w, size)
    if crop is None:
        want = np.stack([ro.resize(t, rh, rw, filter) for t in tiles])
    else:
        want = np.stack([ro.resize(t, rh, rw, filter)[int(round((rh - crop) / 2.0)):, int(round((rw - crop) / 2.0)):]
                         [:crop, :crop] for t in tiles])
    assert got.shape == want.shape and np.array_equal(got, want)


def test_resize_h_and_e_tiles_and_batch_indexing(cuda_device):
    """H&E-like tiles against the oracle, and a 600-tile batch through size-independent properties: resampling
    commutes with a permutation of the batch, constant tiles stay constant."""
    from oracle import resize_oracle as ro
    from oracle import vit_oracle as vo
    from stamp_b200.resize import resize_center_crop

    tiles = vo.synthetic_tiles(6, seed=2)
    got = resize_center_crop(tiles.to(cuda_device), 256, 224)
    assert np.array_equal(got.cpu().numpy(), ro.resize_center_crop(tiles.numpy(), 256, 224))
    big = tiles.to(cuda_device).repeat(100, 1, 1, 1)
    big += torch.arange(600, device=cuda_device, dtype=torch.uint8).view(-1, 1, 1, 1)      # wraps: 600 distinct tiles
    big[17] = 93
    out = resize_center_crop(big, 256, 224)
    perm = torch.randperm(600, generator=torch.Generator().manual_seed(0)).to(cuda_device)
    assert torch.equal(resize_center_crop(big[perm].contiguous(), 256, 224), out[perm])
    assert bool((out[17] == 93).all())
    assert np.array_equal(out[[5, 311, 599]].cpu().numpy(),
                          ro.resize_center_crop(big[[5, 311, 599]].cpu().numpy(), 256, 224))


def test_gigapath_blocks_match_oracle(cuda_device):
    """gigapath.py:14-35: the transform's resampling + ViT-g/16 at full width (dim 1536, 24 heads of 64, SwiGLU 8192,
    197 tokens), two blocks deep, < 1e-3 per tile like every other tile encoder."""
    from dataclasses import replace

    from oracle import resize_oracle as ro
    from oracle import vit_oracle as vo
    from stamp_b200.vit import GIGAPATH_ARCH, TileEncoder

    cfg = replace(vo.GIGAPATH, depth=2)
    w = vo.make_weights(cfg, seed=1234)
    tiles = vo.synthetic_tiles(4, seed=3)
    with torch.no_grad():
        ref = vo.forward(w, cfg, torch.from_numpy(ro.resize_center_crop(tiles.numpy(), 256, 224)))
    enc = TileEncoder(replace(GIGAPATH_ARCH, depth=2), w, max_batch=64).to(cuda_device).eval()
    out = enc(tiles.to(cuda_device))
    assert out.shape == ref.shape and torch.isfinite(out).all()
    err = ((out.double().cpu() - ref.double()).norm(dim=1) / ref.double().norm(dim=1)).max().item()
    assert err < 1e-3, err
    assert enc.launches_per_batch() == 3 + 7 * 2 + 1 + 1
