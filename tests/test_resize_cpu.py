"""The resampling oracle against Pillow / torchvision (the reference's own transform stack), and the product code's
coefficient tables against the oracle's -- CPU only."""

import numpy as np
import pytest
import torch

from oracle import resize_oracle as ro
from oracle import vit_oracle as vo


@pytest.mark.parametrize("shape,out", [((224, 224), (256, 256)), ((224, 224), (112, 112)), ((97, 131), (64, 200)),
                                       ((224, 224), (224, 224)), ((50, 40), (171, 33))])
@pytest.mark.parametrize("filter", ["bicubic", "bilinear"])
def test_oracle_is_pillow_bit_for_bit(shape, out, filter):
    from PIL import Image

    rng = np.random.default_rng(shape[0] * 7 + out[1])
    for img in (rng.integers(0, 256, (*shape, 3), dtype=np.uint8),
                np.where(rng.random((*shape, 3)) < 0.5, 0, 255).astype(np.uint8)):     # maximal overshoot
        want = np.asarray(Image.fromarray(img).resize((out[1], out[0]), getattr(Image, filter.upper())))
        assert np.array_equal(ro.resize(img, out[0], out[1], filter), want)


def test_oracle_is_the_gigapath_transform():
    """gigapath.py:20-27: Resize(256, BICUBIC) -> CenterCrop(224) on the PIL tile."""
    from PIL import Image
    from torchvision import transforms

    tf = transforms.Compose([transforms.Resize(256, interpolation=transforms.InterpolationMode.BICUBIC),
                             transforms.CenterCrop(224)])
    tiles = vo.synthetic_tiles(3, seed=5).numpy()
    want = np.stack([np.asarray(tf(Image.fromarray(t))) for t in tiles])
    assert np.array_equal(ro.resize_center_crop(tiles, 256, 224), want)
    odd = np.random.default_rng(0).integers(0, 256, (2, 200, 260, 3), dtype=np.uint8)
    want = np.stack([np.asarray(tf(Image.fromarray(t))) for t in odd])
    assert np.array_equal(ro.resize_center_crop(odd, 256, 224), want)


def test_product_tables_equal_oracle_tables():
    from stamp_b200.resize import pillow_resample_tables, resized_shape

    for i, o in [(224, 256), (224, 112), (131, 200), (40, 33), (224, 224)]:
        for f in ("bicubic", "bilinear"):
            k, b = pillow_resample_tables(i, o, f)
            ko, bo = ro.coefficients(i, o, f)
            assert k.dtype == np.int32 and np.array_equal(k, ko) and np.array_equal(b, bo)
    assert resized_shape(200, 260, 256) == (256, 332) and resized_shape(224, 224, 256) == (256, 256)
    assert resized_shape(260, 200, 256) == (332, 256) and resized_shape(10, 10, (3, 4)) == (3, 4)


def test_resize_refuses_cpu():
    from stamp_b200.resize import resize_center_crop

    with pytest.raises(RuntimeError):
        resize_center_crop(torch.zeros(1, 224, 224, 3, dtype=torch.uint8), 256, 224)
