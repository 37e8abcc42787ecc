"""Macenko stain normalisation and slide-level pooling parity on the GPU (through the C ABI).

Macenko: uint8 images within 1 LSB of the fp64 NumPy oracle, stain vectors within 1e-4
(SURVEY.md 8c); properties at full size: idempotence of the pass-through, determinism.
CHIEF / EAGLE: attention scores and pooled embedding vs the fp32/fp64 oracle, top-k indices
bit-exact (north_star)."""

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _tiles(n, seed, img=224):
    from oracle import vit_oracle as vo

    return vo.synthetic_tiles(n, seed=seed, img=img)


@pytest.mark.parametrize("n_tiles,tpf", [(8, None), (12, 4), (5, 2), (3, 1)])
def test_macenko_matches_oracle(cuda_device, n_tiles, tpf):
    from oracle import macenko_oracle as mo
    from stamp_b200.macenko import macenko_normalize

    tiles = _tiles(n_tiles, seed=20 + n_tiles)
    ref, he_ref, maxc_ref, valid_ref = mo.normalize(tiles.numpy(), tiles_per_fit=tpf)
    out, fit = macenko_normalize(tiles.to(cuda_device), tiles_per_fit=tpf, return_fit=True)
    assert out.dtype == torch.uint8 and out.shape == tiles.shape
    assert np.array_equal(fit.valid.cpu().numpy(), valid_ref)
    he = fit.stain_matrix.cpu().numpy()
    assert np.abs(he - he_ref).max() < 1e-4, np.abs(he - he_ref).max()
    assert np.abs(fit.max_conc.cpu().numpy() - maxc_ref).max() / np.abs(maxc_ref).max() < 1e-4
    diff = np.abs(out.cpu().numpy().astype(np.int16) - ref.astype(np.int16))
    assert diff.max() <= 1, diff.max()
    assert (diff > 0).mean() < 0.02  # only truncation-boundary pixels may differ


def test_macenko_no_tissue_passthrough_and_mixed_groups(cuda_device):
    from oracle import macenko_oracle as mo
    from stamp_b200.macenko import macenko_normalize

    tiles = _tiles(4, seed=31)
    tiles[2:] = 250  # background: every OD < beta -> no tissue pixels in groups 2, 3
    ref, _, _, valid_ref = mo.normalize(tiles.numpy(), tiles_per_fit=1)
    out, fit = macenko_normalize(tiles.to(cuda_device), tiles_per_fit=1, return_fit=True)
    assert fit.valid.cpu().tolist() == valid_ref.tolist() == [True, True, False, False]
    assert torch.equal(out[2:].cpu(), tiles[2:])
    assert np.abs(out.cpu().numpy().astype(np.int16) - ref.astype(np.int16)).max() <= 1


def test_macenko_deterministic_and_non_224(cuda_device):
    from oracle import macenko_oracle as mo
    from stamp_b200.macenko import macenko_normalize

    tiles = _tiles(6, seed=33, img=96).to(cuda_device)
    a, b = macenko_normalize(tiles), macenko_normalize(tiles)
    assert torch.equal(a, b)
    ref, *_ = mo.normalize(tiles.cpu().numpy())
    assert np.abs(a.cpu().numpy().astype(np.int16) - ref.astype(np.int16)).max() <= 1


def test_macenko_batch_scale_property(cuda_device):
    """Full extraction batch (192 tiles, 29 MB): fit is pooled, so normalising a batch made of the
    same tile set twice gives the same stain matrix as the set itself (percentiles are replication
    invariant up to interpolation)."""
    from stamp_b200.macenko import macenko_normalize

    base = _tiles(16, seed=35).to(cuda_device)
    tiles = base.repeat(12, 1, 1, 1)
    _, fit_big = macenko_normalize(tiles, return_fit=True)
    _, fit_small = macenko_normalize(base, return_fit=True)
    assert (fit_big.stain_matrix - fit_small.stain_matrix).abs().max().item() < 2e-3


@pytest.mark.parametrize("n", [1, 5, 300, 5000])
def test_gated_attention_pool_matches_oracle(cuda_device, n):
    from oracle import chief_oracle as co
    from stamp_b200.encoder import GatedAttentionPool

    sd = co.init_state_dict(seed=1)
    g = torch.Generator().manual_seed(n)
    x = torch.randn(n, 768, generator=g)
    ref = co.forward(sd, x.double())
    out = GatedAttentionPool(sd).to(cuda_device)(x.to(cuda_device))
    a, a_ref = out["attention_raw"].cpu().double(), ref["attention_raw"]
    assert a.shape == a_ref.shape
    assert (a - a_ref).abs().max().item() < 2e-5 * max(1.0, a_ref.abs().max().item())  # fp32-grade scores
    p, p_ref = out["WSI_feature"].cpu().double(), ref["WSI_feature"]
    assert ((p - p_ref).norm() / p_ref.norm()).item() < 1e-4


def test_eagle_topk_indices_bit_exact(cuda_device):
    from oracle import chief_oracle as co
    from stamp_b200.encoder import EagleB200

    sd = co.init_state_dict(seed=2)
    g = torch.Generator().manual_seed(7)
    feats, agg = torch.randn(4000, 768, generator=g), torch.randn(4000, 1280, generator=g)
    emb_ref, idx_ref = co.eagle_embedding(sd, feats, agg)
    enc = EagleB200(sd)
    emb = enc._generate_slide_embedding(feats, cuda_device, agg_feats=agg)
    attn = enc.model(feats.to(cuda_device))["attention_raw"].squeeze(0)
    from stamp_b200.encoder import topk

    _, idx = topk(attn.contiguous(), 25)
    assert torch.equal(idx.cpu(), idx_ref)            # same tiles, same order
    assert np.allclose(emb, emb_ref.numpy(), rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("n,k,largest", [(3, 2, True), (3, 2, False), (1000, 25, True), (50000, 100, True), (70, 64, False), (1024, 1024, True)])
def test_topk_matches_torch(cuda_device, n, k, largest):
    from stamp_b200.encoder import topk

    g = torch.Generator().manual_seed(n + k)
    s = torch.randn(n, generator=g).to(cuda_device)
    val, idx = topk(s, k, largest)
    rv, ri = torch.topk(s, k, largest=largest)
    assert torch.equal(val, rv)
    assert torch.equal(s[idx], rv)
    assert idx.unique().numel() == k


def test_topk_heatmap_fixture_and_ties(cuda_device):
    """tests/test_heatmaps.py:15-59 of the reference: scores [0.1, 0.9, 0.3] -> top (0.9, 0.3),
    bottom (0.1, 0.3); ties resolve to the lower index."""
    from stamp_b200.encoder import topk

    s = torch.tensor([0.1, 0.9, 0.3], device=cuda_device)
    v, i = topk(s, 2, True)
    assert i.tolist() == [1, 2] and v.tolist() == pytest.approx([0.9, 0.3])
    v, i = topk(s, 2, False)
    assert i.tolist() == [0, 2]
    t = torch.tensor([1.0, 2.0, 2.0, 2.0, 0.5], device=cuda_device)
    assert topk(t, 3, True)[1].tolist() == [1, 2, 3]
    assert topk(torch.zeros(40, device=cuda_device), 5, True)[1].tolist() == [0, 1, 2, 3, 4]


def test_chief_and_eagle_match_reference_model_golden(cuda_device):
    """The CUDA pooling path against outputs of the reference's own CHIEFModel / EAGLE lines (chief_pool.npz)."""
    from oracle import chief_oracle as co
    from stamp_b200.encoder import EagleB200, GatedAttentionPool, topk
    from test_oracle_cpu import load_chief_golden

    seed, cases = load_chief_golden()
    sd = co.init_state_dict(seed=seed)
    pool = GatedAttentionPool(sd).to(cuda_device)
    for name, c in cases.items():
        out = pool(c["x"].to(cuda_device))
        a, a_ref = out["attention_raw"].cpu(), c["attention_raw"]
        assert (a - a_ref).abs().max().item() < 2e-5 * max(1.0, a_ref.abs().max().item()), name
        p, p_ref = out["WSI_feature"].cpu().double().flatten(), c["wsi"].double().flatten()
        assert ((p - p_ref).norm() / p_ref.norm()).item() < 1e-4, name
        k = min(25, c["x"].shape[0])
        _, idx = topk(out["attention_raw"].squeeze(0).contiguous(), k)
        assert torch.equal(idx.cpu(), c["topk"].long()), name            # margins checked when the fixture was written
        eagle = EagleB200(sd)
        emb = eagle._generate_slide_embedding(c["x"], cuda_device, agg_feats=c["agg"])
        assert np.allclose(emb, c["eagle"].numpy(), rtol=1e-5, atol=1e-6), name
        # patient level (eagle.py:122-134): the slides' features concatenated -> the same embedding for a split bag
        h = c["x"].shape[0] // 2
        pat = eagle._generate_patient_embedding([c["x"][:h], c["x"][h:]], cuda_device,
                                                agg_feats_list=[c["agg"][:h], c["agg"][h:]])
        assert np.allclose(pat, emb, rtol=1e-6, atol=1e-7), name
        with pytest.raises(ValueError):
            eagle._generate_patient_embedding([c["x"]], cuda_device)


def test_standalone_eagle_feature_file_walk(cuda_device, tmp_path):
    """eagle.py:136-300 through the stand-alone EagleB200: ctranspath and Virchow2 feature files paired by name, the
    Virchow2 tiles stored in another order and re-aligned by coordinates, wrong extractors skipped; slide- and
    patient-level embeddings equal the direct calls."""
    import numpy as np

    from stamp_b200 import encoder as E
    from stamp_b200 import features, h5lite

    if E.BOUND_TO_REFERENCE:
        pytest.skip("the reference's own Eagle (h5py) is in use")
    g = torch.Generator().manual_seed(5)
    sd = {"attention_net.0.weight": torch.randn(512, 768, generator=g) * 0.04, "attention_net.0.bias": torch.zeros(512),
          "attention_net.3.attention_a.0.weight": torch.randn(256, 512, generator=g) * 0.05,
          "attention_net.3.attention_a.0.bias": torch.zeros(256),
          "attention_net.3.attention_b.0.weight": torch.randn(256, 512, generator=g) * 0.05,
          "attention_net.3.attention_b.0.bias": torch.zeros(256),
          "attention_net.3.attention_c.weight": torch.randn(1, 256, generator=g) * 0.06,
          "attention_net.3.attention_c.bias": torch.zeros(1)}
    enc = E.EagleB200(sd)
    ctp_dir, vir_dir = tmp_path / "ctranspath", tmp_path / "virchow2"
    data = {}
    for name, n in (("s1", 90), ("s2", 40), ("bad", 30)):
        ctp = torch.randn(n, 768, generator=g).half()
        vir = torch.randn(n, 2560, generator=g).half()
        cells = torch.randperm(400, generator=g)[:n]
        coords = torch.stack([(cells % 20).float(), (cells // 20).float()], -1).numpy() * 256.0
        perm = torch.randperm(n, generator=g)                      # the Virchow2 file lists the tiles in another order
        features.write_tile_features(ctp_dir / f"{name}.h5", ctp, coords, extractor="chief-ctranspath", tile_size_um=256.0,
                                     tile_size_px=224)
        features.write_tile_features(vir_dir / f"{name}.h5", vir[perm], coords[perm.numpy()],
                                     extractor="uni" if name == "bad" else "virchow2-0a1b2c3d", tile_size_um=256.0, tile_size_px=224)
        data[name] = (ctp, vir)
    enc.encode_slides_(tmp_path / "out", ctp_dir, cuda_device, agg_feat_dir=vir_dir)
    out = tmp_path / "out" / "eagle-slide"
    assert sorted(p.name for p in out.iterdir()) == ["s1.h5", "s2.h5"]             # "bad": wrong aggregation extractor
    for name in ("s1", "s2"):
        ctp, vir = data[name]
        want = enc._generate_slide_embedding(ctp, cuda_device, vir)
        with h5lite.File(out / f"{name}.h5") as h5:
            assert np.array_equal(h5["feats"][()], want) and h5.attrs["encoder"] == "eagle"
    (tmp_path / "slides.csv").write_text("PATIENT,FILENAME\np1,s1.h5\np1,s2.h5\np2,missing.h5\n")
    enc.encode_patients_(tmp_path / "out", ctp_dir, tmp_path / "slides.csv", "PATIENT", "FILENAME", cuda_device,
                         agg_feat_dir=vir_dir)
    pat = tmp_path / "out" / "eagle-pat"
    assert [p.name for p in pat.iterdir()] == ["p1.h5"]
    want = enc._generate_patient_embedding([data["s1"][0], data["s2"][0]], cuda_device, [data["s1"][1], data["s2"][1]])
    with h5lite.File(pat / "p1.h5") as h5:
        assert np.array_equal(h5["feats"][()], want) and h5.attrs["feat_type"] == "patient"
    with pytest.raises(ValueError):
        enc.encode_slides_(tmp_path / "out2", ctp_dir, cuda_device)
