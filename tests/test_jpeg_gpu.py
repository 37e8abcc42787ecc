"""Cached-tile JPEG decode on the GPU (host Huffman decode + csrc/jpeg.cu) against Pillow itself: bit-exact."""

import io
import json
import zipfile

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _jpeg(img: np.ndarray, **kw) -> bytes:
    from PIL import Image

    b = io.BytesIO()
    Image.fromarray(img).save(b, format="jpeg", **kw)
    return b.getvalue()


def _pillow(blob: bytes) -> np.ndarray:
    from PIL import Image

    return np.asarray(Image.open(io.BytesIO(blob)).convert("RGB"))


@pytest.mark.parametrize("h,w,kw", [
    (224, 224, {}),                                   # the reference's cache: Pillow defaults, quality 75, 4:2:0
    (224, 224, dict(quality=95, subsampling=0)),      # 4:4:4
    (256, 256, dict(quality=50)),
    (37, 53, {}),                                     # partial MCUs, odd width and height
    (37, 53, dict(subsampling=0, quality=90)),
    (16, 16, dict(quality=100)),                      # a single MCU: every edge rule of the chroma up-sampling at once
    (224, 224, dict(restart_marker_blocks=5, optimize=True)),
])
def test_gpu_decode_is_pillow_bit_for_bit(cuda_device, h, w, kw):
    from oracle import vit_oracle as vo
    from stamp_b200.jpeg import decode_jpeg_tiles

    rng = np.random.default_rng(h * 1000 + w)
    he = vo.synthetic_tiles(3, seed=h + w, img=max(h, w) if max(h, w) % 2 == 0 else 224).numpy()[:, :h, :w]
    imgs = [he[0], he[1], he[2], rng.integers(0, 256, (h, w, 3), dtype=np.uint8),
            (np.indices((h, w)).sum(0) % 2 * 255).astype(np.uint8)[..., None].repeat(3, -1)]
    blobs = [_jpeg(np.ascontiguousarray(i), **kw) for i in imgs]
    got = decode_jpeg_tiles(blobs, cuda_device, max_workers=2)
    assert got.shape == (len(blobs), h, w, 3) and got.dtype == torch.uint8 and got.is_cuda
    want = np.stack([_pillow(b) for b in blobs])
    assert np.array_equal(got.cpu().numpy(), want)


def test_tile_cache_zip_decoded_on_the_gpu(cuda_device, tmp_path):
    """_tiles_from_cache_file (tiling.py:380-406) with the tiles born in HBM: same pixels, coordinates and order as the
    Pillow reader; a 700-tile cache goes through in batches."""
    from oracle import vit_oracle as vo
    from stamp_b200.tiling import tiles_from_cache_file, tiles_from_cache_file_gpu

    tiles = vo.synthetic_tiles(40, seed=9).numpy()
    path = tmp_path / "slide.abc.zip"
    with zipfile.ZipFile(path, "w") as zf:
        zf.writestr("tiler_params.json", json.dumps({"tile_ext": "jpg", "tile_size_um": 256.0, "tile_size_px": 224}))
        for i in range(700):
            zf.writestr(f"tile_({256.0 * (i % 30)}, {256.0 * (i // 30)}).jpg", _jpeg(np.roll(tiles[i % 40], i, axis=1)))
        zf.writestr("thumbnail.jpg", _jpeg(tiles[0][:64, :64]))          # not a tile: ignored by both readers
    host, coords_h, params_h = tiles_from_cache_file(path, pin_memory=False)
    dev, coords_d, params_d = tiles_from_cache_file_gpu(path, cuda_device, batch=256)
    assert dev.is_cuda and dev.shape == (700, 224, 224, 3) and params_d == params_h
    assert torch.equal(coords_d, coords_h) and torch.equal(dev.cpu(), host)
    # ... and straight into the tile encoder: features from GPU-born tiles == features from the host tiles
    from stamp_b200.extractor import Extractor, extract_slide_features, pil_to_u8_hwc
    from stamp_b200.vit import TileEncoder, VitArch

    cfg = vo.tiny_config(depth=1)
    arch = VitArch(cfg.name, patch=cfg.patch, dim=cfg.dim, depth=cfg.depth, heads=cfg.heads, mlp_hidden=cfg.mlp_hidden)
    ext = Extractor(model=TileEncoder(arch, vo.make_weights(cfg), max_batch=64).to(cuda_device).eval(),
                    transform=pil_to_u8_hwc, identifier="tiny")
    assert torch.equal(extract_slide_features(ext, dev[:100], cuda_device, batch_size=48),
                       extract_slide_features(ext, host[:100], cuda_device, batch_size=48))
    # ... and the pipelined cache -> features path (decode of batch i+1 behind the encoder on batch i, tissue filter in
    # between) keeps exactly the tiles the Pillow + OpenCV path keeps and returns the same features
    from stamp_b200.extractor import extract_cache_features
    from stamp_b200.tiling import has_enough_texture

    blank = np.full((224, 224, 3), 236, dtype=np.uint8)
    path2 = tmp_path / "slide2.zip"
    with zipfile.ZipFile(path2, "w") as zf:
        zf.writestr("tiler_params.json", json.dumps({"tile_ext": "jpg", "tile_size_um": 256.0}))
        for i in range(130):
            img = blank if i % 7 == 3 else np.roll(tiles[i % 40], 3 * i, axis=0)
            zf.writestr(f"tile_({256.0 * (i % 12)}, {256.0 * (i // 12)}).jpg", _jpeg(img))
    host2, coords2, _ = tiles_from_cache_file(path2, pin_memory=False)
    keep = has_enough_texture(host2.to(cuda_device), 0.02).cpu()
    assert 0 < int(keep.sum()) < 130 and not bool(keep[3])
    want = extract_slide_features(ext, host2[keep], cuda_device, batch_size=48)
    got, got_coords, params2 = extract_cache_features(ext, path2, cuda_device, batch_size=48, canny_cutoff=0.02, max_workers=2)
    assert params2["tile_size_um"] == 256.0 and torch.equal(got_coords, coords2[keep]) and torch.equal(got, want)
    every, every_coords, _ = extract_cache_features(ext, path2, cuda_device, batch_size=64, canny_cutoff=None)
    assert every.shape[0] == 130 and torch.equal(every_coords, coords2)
    with zipfile.ZipFile(tmp_path / "png.zip", "w") as zf:
        zf.writestr("tiler_params.json", json.dumps({"tile_ext": "png"}))
    with pytest.raises(ValueError):
        tiles_from_cache_file_gpu(tmp_path / "png.zip", cuda_device)
