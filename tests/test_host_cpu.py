"""CPU suite, part 2: host-side logic (bag plumbing, sharding, the world_size-2 collective path on
gloo, weight packing) -- no GPU, no kernels."""

import os
import socket
import sys
from pathlib import Path

import pytest
import torch
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent


def test_to_fixed_size_bag_matches_reference_semantics():
    from stamp_b200.bags import collate_bags, padding_mask, to_fixed_size_bag

    bag, coords = torch.arange(10.0)[:, None].repeat(1, 3), torch.rand(10, 2)
    b, c, n = to_fixed_size_bag(bag, coords, 16)
    assert b.shape == (16, 3) and n == 10 and torch.equal(b[:10], bag) and (b[10:] == 0).all() and (c[10:] == 0).all()
    b, c, n = to_fixed_size_bag(bag, coords, 4, deterministic=True)
    assert n == 4 and b[:, 0].tolist() == [0.0, 3.0, 6.0, 9.0]  # linspace(0, 9, 4).round()
    g = torch.Generator().manual_seed(0)
    b, c, n = to_fixed_size_bag(bag, coords, 4, generator=g)
    assert n == 4 and len(set(b[:, 0].tolist())) == 4
    # empty bag
    b, c, n = to_fixed_size_bag(torch.zeros(0, 3), torch.zeros(0, 2), 4)
    assert n == 0 and b.shape == (4, 3)
    bags, cs, sizes, tg = collate_bags([(b, c, 3, torch.tensor(1.0)), (b, c, 4, torch.tensor([[0.0, 1.0]]))[:1] * 1
                                        ] if False else [(b, c, 3, torch.tensor([1.0, 0.0])), (b, c, 4, torch.tensor([[0.0, 1.0]]))])
    assert bags.shape == (2, 4, 3) and sizes.tolist() == [3, 4] and tg.shape == (2, 2)
    assert padding_mask(sizes, 4).tolist() == [[False, False, False, True], [False] * 4]


def test_sharding_partitions_are_disjoint_and_complete():
    from stamp_b200.sharding import shard_lpt, shard_round_robin

    slides = [f"s{i:03d}" for i in range(37)]
    g = torch.Generator().manual_seed(7)
    sizes = torch.randint(2000, 10001, (37,), generator=g).tolist()  # configs[2]: U{2000..10000} tiles
    for fn in (lambda r, w: shard_round_robin(slides, r, w), lambda r, w: shard_lpt(slides, sizes, r, w)):
        for w in (1, 2, 8):
            parts = [fn(r, w) for r in range(w)]
            assert sorted(sum(parts, [])) == slides
    # LPT balances unequal slides better than round-robin
    def imbalance(parts):
        loads = [sum(sizes[slides.index(s)] for s in p) for p in parts]
        return max(loads) / (sum(loads) / len(loads))
    rr = imbalance([shard_round_robin(slides, r, 8) for r in range(8)])
    lpt = imbalance([shard_lpt(slides, sizes, r, 8) for r in range(8)])
    assert lpt <= rr and lpt < 1.05


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank: int, world_size: int, port: int, q) -> None:
    sys.path.insert(0, str(ROOT))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import torch.distributed as dist

    from stamp_b200.sharding import (FlatGradAllReducer, all_reduce_flat_sum, gather_to_rank0, shard_round_robin,
                                     sync_alibi_running_mean)
    from stamp_b200.mil import VisionTransformer

    dist.init_process_group("gloo", rank=rank, world_size=world_size)
    torch.manual_seed(0)  # identical replicas
    model = VisionTransformer(dim_output=2, dim_input=16, dim_model=64, n_layers=1, n_heads=2,
                              dim_feedforward=32, dropout=0.0, use_alibi=True)
    for i, p in enumerate(model.parameters()):
        p.grad = torch.full_like(p, float(rank + 1) * (i + 1))
    FlatGradAllReducer(list(model.parameters())).all_reduce_mean()
    ok = all(torch.allclose(p.grad, torch.full_like(p, 1.5 * (i + 1))) for i, p in enumerate(model.parameters()))
    flat = torch.full((1000,), float(rank + 1))
    scale = all_reduce_flat_sum(flat)          # the FusedAdamW.flat_grad exchange of train.data_parallel_step
    ok = ok and scale == 0.5 and torch.allclose(flat * scale, torch.full((1000,), 1.5))
    for n, b in model.named_buffers():
        if n.endswith("running_mean"):
            b.fill_(100.0 * (rank + 1))
    sync_alibi_running_mean(model)
    ok = ok and all(abs(b.item() - 150.0) < 1e-4 for n, b in model.named_buffers() if n.endswith("running_mean"))
    mine = shard_round_robin(list(range(10)), rank, world_size)
    gathered = gather_to_rank0({"rank": rank, "items": mine})
    if rank == 0:
        ok = ok and sorted(sum((g["items"] for g in gathered), [])) == list(range(10))
    else:
        ok = ok and gathered is None
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


def test_world_size_2_gloo_collectives():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = dict(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert results == {0: True, 1: True}


def test_mil_state_dict_keys_match_reference_layout():
    """Key-for-key the reference's state dict -- golden fixtures carry the real one."""
    import numpy as np

    from stamp_b200.mil import VisionTransformer

    for name, alibi in (("mil_alibi_nomask", True), ("mil_mha_nomask", False)):
        z = np.load(ROOT / "tests" / "golden" / f"{name}.npz")
        ref_keys = {k[3:]: z[k].shape for k in z.files if k.startswith("sd/")}
        m = VisionTransformer(dim_output=3, dim_input=64, dim_model=128, n_layers=2, n_heads=2,
                              dim_feedforward=128, dropout=0.25, use_alibi=alibi)
        ours = {k: tuple(v.shape) for k, v in m.state_dict().items()}
        assert ours == {k: tuple(s) for k, s in ref_keys.items()}


def test_vit_weight_packing_swiglu_interleave():
    from stamp_b200.vit import TileEncoder, VitArch, random_state_dict

    arch = VitArch("t", patch=14, dim=64, depth=1, heads=2, mlp_hidden=48, mlp="swiglu", reg_tokens=4)
    sd = random_state_dict(arch)
    sd["blocks.0.mlp.fc1.weight"] = torch.arange(48.0)[:, None].repeat(1, 64)
    enc = TileEncoder(arch, sd)
    w = enc.b0_fc1_w.float()[:, 0].tolist()
    assert w[:6] == [0.0, 24.0, 1.0, 25.0, 2.0, 26.0]  # (x1_j, x2_j) adjacent
    assert enc.patch_w.shape == (64, arch.kpad) and arch.kpad == 592 and (enc.patch_w[:, 588:] == 0).all()
    assert enc.prefix.shape == (5, 64) and enc.pos.shape == (256, 64)


def test_extractor_interface_and_identifiers():
    import dataclasses

    import numpy as np

    from stamp_b200.extractor import Extractor, pil_to_u8_hwc, uni
    from stamp_b200.vit import VitArch

    assert [f.name for f in dataclasses.fields(Extractor)] == ["model", "transform", "identifier"]
    with pytest.raises(TypeError):
        Extractor(None, None, "x")  # keyword-only like the reference dataclass
    t = pil_to_u8_hwc(np.zeros((224, 224, 3), dtype=np.uint8))
    assert t.dtype == torch.uint8 and t.shape == (224, 224, 3)
    import stamp_b200.extractor as ex
    ex.UNI_ARCH = VitArch("uni", dim=64, depth=1, heads=2, mlp_hidden=128)  # tiny stand-in for a CPU test
    e = uni(weights="random")
    assert e.identifier == "uni" and e.transform is pil_to_u8_hwc
    with pytest.raises(RuntimeError):
        e.model(torch.zeros(1, 224, 224, 3, dtype=torch.uint8))  # no CPU fallback


def test_uni2_and_h_optimus_architectures_follow_reference_kwargs():
    """extractor configs the reference spells out (uni2.py:18-32, h_optimus_0.py:14-28)."""
    from stamp_b200.vit import H_OPTIMUS_ARCH, UNI2_ARCH, TileEncoder, VitArch, random_state_dict

    assert (UNI2_ARCH.patch, UNI2_ARCH.dim, UNI2_ARCH.depth, UNI2_ARCH.heads, UNI2_ARCH.reg_tokens) == (14, 1536, 24, 24, 8)
    assert UNI2_ARCH.mlp_hidden == int(1536 * 2.66667 * 2) and UNI2_ARCH.no_embed_class
    assert UNI2_ARCH.n_tokens == 256 + 9 and H_OPTIMUS_ARCH.n_tokens == 256 + 5
    assert H_OPTIMUS_ARCH.mean == (0.707223, 0.578729, 0.703617)
    # no_embed_class packing: prefix rows carry no position, patch rows carry the whole table
    arch = VitArch("t", patch=14, dim=64, depth=1, heads=1, mlp_hidden=128, mlp="swiglu", reg_tokens=2, no_embed_class=True)
    sd = random_state_dict(arch)
    assert sd["pos_embed"].shape == (1, 256, 64)
    enc = TileEncoder(arch, sd)
    assert torch.equal(enc.prefix[0], sd["cls_token"].reshape(-1)) and torch.equal(enc.pos, sd["pos_embed"][0])
    bad = dict(sd, pos_embed=torch.zeros(1, 100, 64))
    with pytest.raises(ValueError):
        TileEncoder(arch, bad)


def test_tile_cache_reader_matches_pillow(tmp_path):
    """tiles_from_cache_file mirrors _tiles_from_cache_file (tiling.py:380-406): same tiles, coordinates, order."""
    import io
    import json
    from zipfile import ZipFile

    import numpy as np
    from PIL import Image

    from stamp_b200.tiling import tiles_from_cache_file

    rng = np.random.default_rng(0)
    for ext, fmt in (("jpg", "JPEG"), ("png", "PNG")):
        tiles = rng.integers(0, 256, (5, 32, 32, 3), dtype=np.uint8)
        coords = [(256.0 * i, 512.5 + i) for i in range(5)]
        path = tmp_path / f"cache_{ext}.zip"
        with ZipFile(path, "w") as zf:
            zf.writestr("tiler_params.json", json.dumps({"tile_size_um": 256.0, "tile_size_px": 32, "tile_ext": ext}))
            zf.writestr("notes.txt", "ignored")
            for t, (x, y) in zip(tiles, coords):
                buf = io.BytesIO()
                Image.fromarray(t).save(buf, format=fmt)
                zf.writestr(f"tile_({x}, {y}).{ext}", buf.getvalue())
        got, c, params = tiles_from_cache_file(path, pin_memory=False)
        assert got.shape == (5, 32, 32, 3) and got.dtype == torch.uint8 and params["tile_size_um"] == 256.0
        assert torch.allclose(c, torch.tensor(coords))
        with ZipFile(path) as zf:          # the reference's decode: PIL on the zip member
            for i, (x, y) in enumerate(coords):
                ref = np.asarray(Image.open(io.BytesIO(zf.read(f"tile_({x}, {y}).{ext}"))).convert("RGB"))
                assert np.array_equal(got[i].numpy(), ref)
        if fmt == "PNG":
            assert np.array_equal(got.numpy(), tiles)


def _reference_functions(path: Path, names: list[str]) -> dict:
    """Executes individual pure-torch functions of a reference file (the file itself needs h5py / lightning):
    their source segments are taken with ``ast`` and run in a namespace that only has torch."""
    import ast

    src = path.read_text()
    tree = ast.parse(src)
    ns: dict = {"torch": torch}
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name in names:
            exec("from __future__ import annotations\n" + ast.get_source_segment(src, node), ns)
    return ns


def test_bag_plumbing_matches_the_reference_functions():
    """bags.to_fixed_size_bag / collate_bags against the reference's own _to_fixed_size_bag / _collate_to_tuple
    (src/stamp/modeling/data.py:811-862, :255-277), executed from the reference file where it is present."""
    ref_file = Path("/root/reference/src/stamp/modeling/data.py")
    if not ref_file.exists():
        pytest.skip("reference checkout not present on this machine")
    from stamp_b200.bags import collate_bags, to_fixed_size_bag

    ref = _reference_functions(ref_file, ["_to_fixed_size_bag", "_collate_to_tuple"])
    g = torch.Generator().manual_seed(5)
    items_ref, items_ours = [], []
    for n, bag_size, det in [(10, 16, False), (100, 16, False), (100, 16, True), (16, 16, False), (1, 4, True), (37, 8, True)]:
        bag, coords = torch.randn(n, 6, generator=g), torch.rand(n, 2, generator=g) * 1000
        torch.manual_seed(n * 31 + bag_size)
        rb, rc, rn = ref["_to_fixed_size_bag"](bag, coords, bag_size, deterministic=det)
        torch.manual_seed(n * 31 + bag_size)                     # the same randperm draw
        ob, oc, on = to_fixed_size_bag(bag, coords, bag_size, deterministic=det)
        assert torch.equal(rb, ob) and torch.equal(rc, oc) and rn == on, (n, bag_size, det)
        if bag_size == 16:
            tgt = torch.tensor([[0.0, 1.0]]) if n == 10 else torch.tensor([1.0, 0.0])
            items_ref.append((rb, rc, rn, tgt))
            items_ours.append((ob, oc, on, tgt))
    for a, b in zip(ref["_collate_to_tuple"](items_ref), collate_bags(items_ours)):
        assert torch.equal(a, b)


def test_brightness_rejection_matches_the_reference_function():
    """tiling.foreground_coords against the reference's _foreground_coords (src/stamp/preprocessing/tiling.py:250-277)
    executed from its source with a stand-in slide object (openslide is not installed)."""
    import numpy as np
    from PIL import Image

    ref_file = Path("/root/reference/src/stamp/preprocessing/tiling.py")
    if not ref_file.exists():
        pytest.skip("reference checkout not present on this machine")
    from stamp_b200.tiling import foreground_coords

    ref = _reference_functions(ref_file, ["_foreground_coords"])
    import numpy.typing as npt

    ref.update(np=np, cast=lambda _t, v: v, npt=npt,
               _XYCoords=lambda x, y: (int(x), int(y)), SlidePixels=int)
    rng = np.random.default_rng(3)
    for (w, h), tile in [((5000, 3100), 448), ((4096, 4096), 512), ((1000, 700), 333)]:
        base = rng.integers(0, 256, size=(h // 20, w // 20, 3), dtype=np.uint8)
        base[: h // 60] = 250                                     # a bright background band

        class Slide:
            dimensions = (w, h)

            @staticmethod
            def get_thumbnail(size):
                img = Image.fromarray(base)
                img.thumbnail(size)                                # what openslide's get_thumbnail ends with
                return img

        for cutoff in (224, 128, None):
            want = list(ref["_foreground_coords"](Slide, tile, cutoff))
            grid = np.ceil(np.array(Slide.dimensions) / tile).astype(np.uint32)
            got = foreground_coords(Slide.dimensions, tile, Slide.get_thumbnail(tuple(grid * 2)), cutoff)
            assert got == want and (cutoff is not None or len(got) == grid[0] * grid[1])
        assert 0 < len(foreground_coords(Slide.dimensions, tile, Slide.get_thumbnail(tuple(grid * 2)), 224)) < grid[0] * grid[1]


def test_eagle_coordinate_alignment_matches_the_reference_function():
    """encoder.align_by_coords against the reference's _align_vir2_to_ctp_by_coords (eagle.py:303-330) executed from source:
    same permutation of the Virchow2 features for shuffled tiles (incl. duplicate coordinates), same errors."""
    import numpy as np

    ref_file = Path("/root/reference/src/stamp/encoding/encoder/eagle.py")
    if not ref_file.exists():
        pytest.skip("reference checkout not present on this machine")
    from collections import defaultdict, deque

    from stamp_b200.encoder import align_by_coords

    ref = _reference_functions(ref_file, ["_align_vir2_to_ctp_by_coords"])
    ref.update(np=np, defaultdict=defaultdict, deque=deque)
    rng = np.random.default_rng(0)
    coords = rng.integers(0, 40, (200, 2)).astype(np.float64) * 256.0 + 1e-7 * rng.standard_normal((200, 2))
    coords[17] = coords[5]                                     # a duplicated coordinate
    perm = rng.permutation(200)
    other_feats = torch.randn(200, 8)
    want_f, want_c = ref["_align_vir2_to_ctp_by_coords"](coords, coords[perm], other_feats, decimals=5)
    got_f, got_c = align_by_coords(coords, coords[perm], other_feats, decimals=5)
    assert torch.equal(got_f, want_f) and np.array_equal(got_c, want_c)
    for bad_other in (coords[perm][:-1], np.vstack([coords[perm], coords[:1]])):
        feats = torch.randn(len(bad_other), 8)
        with pytest.raises(ValueError) as e_ref:
            ref["_align_vir2_to_ctp_by_coords"](coords, bad_other, feats)
        with pytest.raises(ValueError) as e_got:
            align_by_coords(coords, bad_other, feats)
        assert str(e_ref.value) == str(e_got.value)


def test_titan_wrapper_input_preparation():
    """titan.py:47-53 (um -> int64 px) and :131-168 (virtual slide: slides side by side along x)."""
    import numpy as np

    from stamp_b200.encoder import titan_coords_px, titan_virtual_slide

    px = titan_coords_px(np.array([[0.0, 256.0], [511.9, 1024.2]]), mpp=0.5)
    assert px.dtype == torch.int64 and px.tolist() == [[0, 512], [1023, 2048]]
    f1, f2 = torch.ones(2, 4), 2 * torch.ones(3, 4)
    c1 = np.array([[0.0, 0.0], [256.0, 256.0]])
    c2 = np.array([[0.0, 0.0], [512.0, 0.0], [256.0, 256.0]])
    feats, coords = titan_virtual_slide([f1, f2], [c1, c2], tile_size_um=256.0)
    assert feats.shape == (1, 5, 4) and coords[:2].tolist() == c1.tolist()
    assert coords[2:, 0].tolist() == [512.0, 1024.0, 768.0] and coords[2:, 1].tolist() == [0.0, 0.0, 256.0]
    assert c2[0, 0] == 0.0                                    # inputs are not modified in place


def test_crossval_splits_follow_the_reference_splitter():
    """crossval.py:373-423: StratifiedKFold(shuffle=True, random_state=0) over the patient ids."""
    import numpy as np
    from sklearn.model_selection import StratifiedKFold

    from stamp_b200.sharding import crossval_splits, folds_for_rank

    patients = [f"pat{i:03d}" for i in range(64)]
    labels = ["MSI" if i % 4 == 0 else "MSS" for i in range(64)]
    splits = crossval_splits(patients, labels, 5)
    ref = StratifiedKFold(n_splits=5, shuffle=True, random_state=0).split(np.array(patients), np.array(labels))
    for (tr, te), (rtr, rte) in zip(splits, ref):
        assert tr == np.array(patients)[rtr].tolist() and te == np.array(patients)[rte].tolist()
    tests = [set(te) for _, te in splits]
    assert set().union(*tests) == set(patients) and sum(len(t) for t in tests) == 64      # a partition
    assert all(abs(sum(labels[patients.index(p)] == "MSI" for p in te) - 16 / 5) < 1.01 for _, te in splits)
    assert len(crossval_splits(patients, None, 4)) == 4
    assert [folds_for_rank(5, r, 8) for r in range(8)] == [[0], [1], [2], [3], [4], [], [], []]
    assert [folds_for_rank(5, r, 2) for r in range(2)] == [[0, 2, 4], [1, 3]]


def test_vals_to_im_matches_the_reference_function():
    """heatmaps/__init__.py:142-156, executed from the reference file where it is present."""
    ref_file = Path("/root/reference/src/stamp/heatmaps/__init__.py")
    if not ref_file.exists():
        pytest.skip("reference checkout not present on this machine")
    from torch import Tensor

    from stamp_b200.heatmaps import vals_to_im

    ns = _reference_functions(ref_file, ["_vals_to_im"])
    ns["Tensor"] = Tensor
    g = torch.Generator().manual_seed(3)
    cells = torch.randperm(7 * 5, generator=g)[:20]
    coords = torch.stack([cells % 7, cells // 7], dim=-1)
    for scores in (torch.rand(20, 3, generator=g), torch.rand(20, generator=g)):
        assert torch.equal(vals_to_im(scores, coords), ns["_vals_to_im"](scores, coords))
    with pytest.raises(RuntimeError):
        from stamp_b200.heatmaps import ranked_tiles

        ranked_tiles(torch.rand(5), 2, 2)


def test_task_aware_folds_match_the_reference_get_splits():
    """crossval.stratification_labels + sharding.crossval_splits against the reference's own ``_get_splits``
    (src/stamp/modeling/crossval.py:373-423, executed from its source with the splitter ``categorical_crossval_`` picks,
    :90-96) for the three tasks: identical train / test patient sets per fold."""
    ref_file = Path("/root/reference/src/stamp/modeling/crossval.py")
    if not ref_file.exists():
        pytest.skip("reference checkout not present on this machine")
    from collections import namedtuple
    from typing import Any, cast

    import numpy as np
    from sklearn.model_selection import KFold, StratifiedKFold

    from stamp_b200.crossval import stratification_labels
    from stamp_b200.sharding import crossval_splits

    ns = _reference_functions(ref_file, ["_get_splits"])
    Split = namedtuple("Split", "train_patients test_patients")
    Splits = namedtuple("Splits", "splits")
    ns.update(np=np, cast=cast, Any=Any, _Split=Split, _Splits=Splits)
    Data = namedtuple("Data", "ground_truth")
    ids = [f"pat{i:03d}" for i in range(53)]
    cases = {"classification": ["MSI" if i % 3 == 0 else "MSS" for i in range(53)],
             "survival": [(float(100 - i), int(i % 4 != 0)) for i in range(53)],
             "regression": [float(i) * 0.1 for i in range(53)]}
    for task, labels in cases.items():
        spliter = KFold if task == "regression" else StratifiedKFold
        ref = ns["_get_splits"](patient_to_data={p: Data(l) for p, l in zip(ids, labels)}, n_splits=4, spliter=spliter, task=task)
        ours = crossval_splits(ids, stratification_labels(labels, task), 4)
        assert len(ref.splits) == len(ours) == 4
        for r, (tr, te) in zip(ref.splits, ours):
            assert set(tr) == r.train_patients and set(te) == r.test_patients, task
    with pytest.raises(ValueError):
        stratification_labels([1, 2], "ranking")
