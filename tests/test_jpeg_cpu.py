"""Cached-tile JPEG decode on the CPU: the oracle against Pillow (bit for bit), and the product's host half (marker
parsing + Huffman decoding in libstamp_b200.so, no GPU work) against the oracle's coefficients."""

import io

import numpy as np
import pytest
import torch

from oracle import jpeg_oracle as jo
from oracle import vit_oracle as vo


def _jpeg(img: np.ndarray, **kw) -> bytes:
    from PIL import Image

    b = io.BytesIO()
    Image.fromarray(img).save(b, format="jpeg", **kw)
    return b.getvalue()


def _cases():
    rng = np.random.default_rng(0)
    he = vo.synthetic_tiles(2, seed=1).numpy()
    noise = rng.integers(0, 256, (224, 224, 3), dtype=np.uint8)
    return [("h&e, Pillow defaults (q75, 4:2:0)", he[0], {}),
            ("h&e, q95, 4:4:4", he[1], dict(quality=95, subsampling=0)),
            ("noise, q75", noise, {}),
            ("noise, q30, 64 x 80", noise[:64, :80], dict(quality=30)),
            ("37 x 53 (partial MCUs), 4:2:0", noise[:37, :53], {}),
            ("37 x 53, 4:4:4, q90", noise[:37, :53], dict(quality=90, subsampling=0)),
            ("restart markers", he[0], dict(restart_marker_blocks=7)),
            ("optimised Huffman tables", he[1], dict(optimize=True)),
            ("saturated checkerboard", (np.indices((64, 64)).sum(0) % 2 * 255).astype(np.uint8)[..., None].repeat(3, -1), {})]


@pytest.mark.parametrize("name,img,kw", _cases(), ids=[c[0] for c in _cases()])
def test_oracle_is_pillow_bit_for_bit(name, img, kw):
    from PIL import Image

    data = _jpeg(img, **kw)
    want = np.asarray(Image.open(io.BytesIO(data)).convert("RGB"))
    assert np.array_equal(jo.decode(data), want)


@pytest.mark.parametrize("name,img,kw", _cases(), ids=[c[0] for c in _cases()])
def test_host_entropy_decoder_matches_oracle(name, img, kw):
    from stamp_b200 import jpeg

    data = _jpeg(img, **kw)
    info, coef, quant = jpeg.entropy_decode([data], max_workers=1)
    h, w, comps, q, co = jo.parse(data)
    assert (info.height, info.width, info.n_comp) == (h, w, 3)
    assert [(info.h[i], info.v[i]) for i in range(3)] == [c[1:3] for c in comps]
    assert np.array_equal(coef[0].numpy(), np.concatenate([co[i].reshape(-1) for i in range(3)]))
    for i, c in enumerate(comps):
        assert np.array_equal(quant[0, i].numpy().astype(np.uint16), q[c[3]].astype(np.uint16))


def test_host_decoder_batches_and_rejects_what_it_does_not_read():
    from PIL import Image

    from stamp_b200 import _lib, jpeg

    tiles = vo.synthetic_tiles(9, seed=4).numpy()
    blobs = [_jpeg(t) for t in tiles]
    info, coef, quant = jpeg.entropy_decode(blobs, max_workers=4)
    _, one, _ = jpeg.entropy_decode(blobs[5:6], max_workers=1)
    assert coef.shape == (9, 28 * 28 * 64 + 2 * 14 * 14 * 64) and torch.equal(coef[5], one[0])
    gray = io.BytesIO()
    Image.fromarray(tiles[0][..., 0]).save(gray, format="jpeg")
    prog = _jpeg(tiles[0], progressive=True)
    for bad in (gray.getvalue(), prog, b"\xff\xd8\xff\xd9", blobs[0][: len(blobs[0]) // 3], b"not a jpeg at all"):
        with pytest.raises(_lib.StampB200Error):
            jpeg.entropy_decode([bad], max_workers=1)
    with pytest.raises(_lib.StampB200Error):          # one geometry per batch
        jpeg.entropy_decode([blobs[0], _jpeg(tiles[1][:64, :64])], max_workers=1)
    with pytest.raises(RuntimeError):
        jpeg.decode_jpeg_tiles(blobs, "cpu")


def test_host_decoder_survives_corrupt_files():
    """3 000 random mutations of valid tiles (byte flips, truncations, inserted markers, overwritten segment lengths):
    the Huffman decoder returns a status, it never reads or writes out of bounds (the same corpus generator, 30 000
    inputs, runs clean under AddressSanitizer / UBSan with scripts/jpeg_fuzz_driver.cpp)."""
    from stamp_b200 import _lib, jpeg

    rng = np.random.default_rng(0)
    tiles = vo.synthetic_tiles(2, seed=1).numpy()
    seeds = [_jpeg(np.ascontiguousarray(tiles[0][:64, :80])), _jpeg(np.ascontiguousarray(tiles[1][:37, :53]), subsampling=0),
             _jpeg(np.ascontiguousarray(tiles[0][:48, :48]), restart_marker_blocks=3)]
    ok = bad = 0
    for s in seeds:
        for _ in range(1000):
            a = bytearray(s)
            mode = rng.integers(0, 4)
            if mode == 0:
                for _ in range(rng.integers(1, 6)):
                    a[int(rng.integers(0, min(len(a), 700) if rng.random() < 0.7 else len(a)))] = int(rng.integers(0, 256))
            elif mode == 1:
                a = a[: int(rng.integers(0, len(a)))]
            elif mode == 2:
                i = int(rng.integers(0, len(a)))
                a[i:i] = bytes([0xFF, int(rng.integers(0, 256))])
            else:
                i = int(rng.integers(2, min(len(a) - 2, 650)))
                a[i], a[i + 1] = 0xFF, int(rng.choice([0xC0, 0xC4, 0xDA, 0xDB, 0xDD, 0xD9, 0xC2]))
            try:
                info = jpeg.read_header(bytes(a))
                if jpeg.coef_count(info) > 1 << 24:
                    bad += 1
                    continue
                jpeg.entropy_decode([bytes(a)], max_workers=1)
                ok += 1
            except _lib.StampB200Error:
                bad += 1
    assert ok > 100 and bad > 100, (ok, bad)


def test_host_decoder_plus_oracle_arithmetic_is_pillow_on_random_images():
    """The product's host half (marker parsing + Huffman decoding) followed by the oracle's integer arithmetic equals
    Pillow for random sizes, qualities and chroma layouts -- the CPU-side pin of everything the GPU kernels are fed."""
    from hypothesis import given, settings
    from hypothesis import strategies as st
    from PIL import Image

    from stamp_b200 import jpeg

    @settings(max_examples=40, deadline=None, derandomize=True)
    @given(h=st.integers(1, 70), w=st.integers(2, 70), quality=st.integers(5, 100), sub=st.sampled_from([0, 2]),
           seed=st.integers(0, 10_000), smooth=st.booleans())
    def check(h, w, quality, sub, seed, smooth):
        rng = np.random.default_rng(seed)
        img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        if smooth:
            img = (np.cumsum(np.cumsum(img.astype(np.int64), 0), 1) // (np.arange(1, h + 1)[:, None, None] * np.arange(1, w + 1)[None, :, None])).astype(np.uint8)
        data = _jpeg(np.ascontiguousarray(img), quality=quality, subsampling=sub)
        want = np.asarray(Image.open(io.BytesIO(data)).convert("RGB"))
        info, coef, quant = jpeg.entropy_decode([data], max_workers=1)
        c = coef[0].numpy()
        planes, off = [], 0
        for i in range(3):
            by, bx = info.mcus_y * info.v[i], info.mcus_x * info.h[i]
            n = by * bx * 64
            planes.append(jo.idct_islow(c[off:off + n].reshape(by, bx, 64), quant[0, i].numpy().astype(np.uint16).astype(np.int32)))
            off += n
        if info.h[0] == 2:
            cb, cr = (jo.h2v2_fancy_upsample(p, -(-h // 2), -(-w // 2)) for p in planes[1:])
        else:
            cb, cr = planes[1], planes[2]
        got = jo.ycc_to_rgb(planes[0][:h, :w], cb[:h, :w], cr[:h, :w])
        assert np.array_equal(got, want)

    check()
