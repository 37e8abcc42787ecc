"""The whole hot path composed on the GPU with synthetic data (BASELINE configs[0] plumbing, on the device):

  uint8 tiles -> tissue-texture filter -> Macenko -> tile encoder (small ViT, all kernels of the ViT path)
  -> fp16 features + coordinates -> fixed-size bags -> MIL training steps (FusedAdamW + OneCycleLR)
  -> patient-level deploy -> grad-CAM + top-k tiles

Checks the dtype / shape contracts between the stages and that nothing leaves the device except through the
documented interfaces; the arithmetic of every stage has its own parity test."""

import pytest
import torch

pytestmark = pytest.mark.gpu


def test_tiles_to_heatmap_pipeline(cuda_device):
    from oracle import vit_oracle as vo
    from stamp_b200 import train as T
    from stamp_b200.bags import collate_bags, to_fixed_size_bag
    from stamp_b200.deploy import predict_patients
    from stamp_b200.encoder import topk
    from stamp_b200.extractor import Extractor, extract_slide_features, pil_to_u8_hwc
    from stamp_b200.macenko import macenko_normalize
    from stamp_b200.mil import VisionTransformer
    from stamp_b200.tiling import has_enough_texture
    from stamp_b200.vit import TileEncoder, VitArch

    torch.manual_seed(0)
    cfg = vo.tiny_config(depth=2)
    arch = VitArch(cfg.name, patch=cfg.patch, dim=cfg.dim, depth=cfg.depth, heads=cfg.heads,
                   mlp_hidden=cfg.mlp_hidden, mlp=cfg.mlp, reg_tokens=cfg.reg_tokens)
    ext = Extractor(model=TileEncoder(arch, vo.make_weights(cfg), max_batch=32).to(cuda_device).eval(),
                    transform=pil_to_u8_hwc, identifier="tiny")

    slides = {}
    for pid in range(6):
        tiles = vo.synthetic_tiles(40, seed=pid)                       # H&E-like uint8 [40, 224, 224, 3]
        tiles[-4:] = 240                                                # four blank background tiles
        dev_tiles = tiles.to(cuda_device)
        keep = has_enough_texture(dev_tiles, 0.02)
        assert keep.dtype == torch.bool and int(keep.sum()) == 36 and not bool(keep[-4:].any())
        kept = macenko_normalize(dev_tiles[keep].contiguous())
        assert kept.dtype == torch.uint8 and kept.shape == (36, 224, 224, 3)
        feats = extract_slide_features(ext, kept.cpu(), cuda_device, batch_size=16)
        assert feats.dtype == torch.float16 and feats.shape == (36, arch.dim) and not feats.is_cuda
        cells = torch.randperm(100, generator=torch.Generator().manual_seed(pid))[:36]
        coords = torch.stack([(cells % 10).float(), (cells // 10).float()], dim=-1) * 256.0
        slides[f"p{pid}"] = (feats, coords, pid % 2)

    model = VisionTransformer(dim_output=2, dim_input=arch.dim, dim_model=128, n_layers=2, n_heads=2,
                              dim_feedforward=128, dropout=0.1, use_alibi=True).to(cuda_device).train()
    opt, sched = T.configure_optimizers(model, total_steps=6, max_lr=1e-3)
    items = []
    for feats, coords, label in slides.values():
        bag, c, size = to_fixed_size_bag(feats.float().to(cuda_device), coords.to(cuda_device), bag_size=32)
        items.append((bag, c, size, torch.nn.functional.one_hot(torch.tensor(label), 2).float().to(cuda_device)))
    batch = collate_bags(items)
    assert batch[0].shape == (6, 32, arch.dim) and batch[3].shape == (6, 2)
    losses = [float(T.data_parallel_step(model, opt, batch, torch.tensor([1.0, 1.2], device=cuda_device), sched))
              for _ in range(6)]
    assert all(torch.isfinite(torch.tensor(losses)))      # convergence itself: test_loss_decreases_on_planted_signal

    preds = predict_patients(model, list(slides), ((f, c) for f, c, _ in slides.values()), cuda_device)
    assert list(preds) == list(slides)
    assert all(p.shape == (2,) and abs(float(p.sum()) - 1.0) < 1e-5 for p in preds.values())

    feats, coords, _ = slides["p0"]
    cam = T.gradcam_per_category(model, feats.float().to(cuda_device), coords.to(cuda_device))
    assert cam.shape == (36, 2) and torch.allclose(cam.sum(0), torch.ones(2, device=cuda_device), atol=1e-4)
    vals, idx = topk(cam[:, 1].contiguous(), 8)
    # (after a few steps on 6 slides the map is nearly flat and fp32 ties are common: compare the selected VALUES;
    # tie order is covered where the reference fixes it, test_heatmaps_gpu / test_kernels_gpu)
    col = cam[:, 1].cpu()
    assert idx.shape == (8,) and len(set(idx.tolist())) == 8
    assert torch.equal(vals.cpu(), col.topk(8).values) and torch.equal(col[idx.cpu().long()], vals.cpu())


def test_crossval_on_resident_bags(cuda_device):
    """categorical_crossval_ (src/stamp/modeling/crossval.py:48-370) at small scale: 40 patients with a planted signal
    (mean shift on 10 % of the tiles of class 1), 4 folds, bags drawn to 64 of ~100 tiles, test fold as validation set,
    early stopping bookkeeping, best checkpoint restored, predictions for every patient exactly once."""
    from stamp_b200.crossval import Patient, class_weights, crossval
    from stamp_b200.sharding import crossval_splits

    g = torch.Generator().manual_seed(0)
    pats = []
    for i in range(40):
        n = 90 + int(torch.randint(0, 30, (1,), generator=g))
        f = torch.randn(n, 64, generator=g)
        if i % 2:
            f[: n // 10] += 1.5
        cells = torch.randperm(400, generator=g)[:n]
        c = torch.stack([(cells % 20).float(), (cells // 20).float()], dim=-1) * 256.0
        pats.append(Patient(f"p{i:02d}", f.half().to(cuda_device), c.to(cuda_device), i % 2))
    w = class_weights([0, 0, 0, 1], 2, "cpu")
    assert torch.allclose(w, torch.tensor([0.25, 0.75]))           # inverse frequencies, normalised
    res = crossval(pats, n_splits=4, n_classes=2, dim_input=64, mode="fold_per_gpu",
                   model_params=dict(dim_model=128, n_heads=2, dim_feedforward=128, dropout=0.0, use_alibi=True),
                   bag_size=64, batch_size=8, max_epochs=6, patience=2, max_lr=3e-3, seed=1)
    assert [r.fold for r in res] == [0, 1, 2, 3]
    splits = crossval_splits([p.pid for p in pats], [p.label for p in pats], 4)
    seen = []
    for r, (_, te) in zip(res, splits):
        assert r.test_patients == te and r.probs.shape == (len(te), 2)
        assert torch.allclose(r.probs.sum(1), torch.ones(len(te)), atol=1e-5)
        assert 1 <= r.epochs_run <= 6 and 0 <= r.best_epoch < r.epochs_run
        assert r.train_steps == r.epochs_run * 4                   # 30 training patients / batch 8 -> 4 steps per epoch
        best = min(h["validation_loss"] for h in r.history)
        assert abs(r.history[r.best_epoch]["validation_loss"] - best) < 1e-12
        seen += te
    assert sorted(seen) == sorted(p.pid for p in pats)              # every patient predicted exactly once
    # the planted signal is learnt: held-out accuracy well above chance over the four folds
    label = {p.pid: p.label for p in pats}
    hits = sum(int(r.probs[i].argmax()) == label[pid] for r in res for i, pid in enumerate(r.test_patients))
    assert hits >= 30, hits


def test_prefetch_to_device_keeps_order_and_values(cuda_device):
    """The host->device feed of the training loop: batches arrive on the device in order, bit-identical, None entries
    pass through, and the generator drains its queue at the end."""
    from stamp_b200.bags import prefetch_to_device

    g = torch.Generator().manual_seed(5)
    batches = [(torch.randn(3, 50, 16, generator=g).half(), torch.rand(3, 50, 2, generator=g), None,
                torch.eye(2)[torch.arange(3) % 2]) for _ in range(5)]
    seen = 0
    for i, (bags, coords, sizes, targets) in enumerate(prefetch_to_device(iter(batches), cuda_device, depth=3)):
        assert bags.is_cuda and coords.is_cuda and targets.is_cuda and sizes is None
        assert torch.equal(bags.cpu(), batches[i][0]) and torch.equal(coords.cpu(), batches[i][1])
        seen += 1
    assert seen == 5
    assert list(prefetch_to_device(iter([]), cuda_device)) == []


def test_predict_bags_graph_replay_matches_eager(cuda_device):
    """deploy.predict_bags captures a repeated bag shape as a CUDA graph: same probabilities as the eager path, from
    host and device bags, also after the weights change (the capture is dropped with the packed weights)."""
    from stamp_b200 import deploy
    from stamp_b200.mil import VisionTransformer, bump_weights_epoch

    torch.manual_seed(3)
    model = VisionTransformer(dim_output=3, dim_input=64, dim_model=128, n_layers=2, n_heads=2, dim_feedforward=128,
                              dropout=0.0, use_alibi=True).to(cuda_device).eval()
    g = torch.Generator().manual_seed(9)
    bags = [(torch.randn(300, 64, generator=g).half(), torch.rand(300, 2, generator=g) * 5000) for _ in range(9)]
    bags += [(torch.randn(77, 64, generator=g).half(), torch.rand(77, 2, generator=g) * 5000)]      # an odd one out
    eager = deploy.predict_bags(model, iter(bags), cuda_device, graphs=False)
    deploy._GRAPHS.clear(); deploy._SEEN.clear()
    replay = deploy.predict_bags(model, iter(bags), cuda_device)                 # captures from the 4th bag on
    assert len(deploy._GRAPHS) == 3 and torch.allclose(replay, eager, atol=1e-6)
    dev_bags = [(f.to(cuda_device), c.to(cuda_device)) for f, c in bags]
    assert torch.allclose(deploy.predict_bags(model, iter(dev_bags), cuda_device), eager, atol=1e-6)
    with torch.no_grad():
        model.mlp_head[0].bias[0].add_(1.0)          # (a shift of ALL logits would not move the softmax)
    bump_weights_epoch()
    after = deploy.predict_bags(model, iter(bags), cuda_device)
    assert torch.allclose(after, deploy.predict_bags(model, iter(bags), cuda_device, graphs=False), atol=1e-6)
    assert (after - eager).abs().max() > 1e-3


def test_feature_files_between_extraction_and_training(cuda_device, tmp_path):
    """BASELINE configs[0] end to end with the files in between: tiles -> ``extract_to_feature_files`` (``.h5`` per slide
    with the reference's datasets / attributes, written by the background thread) -> ``load_cohort_to_device`` (one
    HBM-resident fp16 tensor) -> deploy on the resident bags; the file round trip is bit-exact."""
    from oracle import vit_oracle as vo
    from stamp_b200 import features, h5lite
    from stamp_b200.deploy import predict_bags
    from stamp_b200.extractor import Extractor, extract_slide_features, extract_to_feature_files, pil_to_u8_hwc
    from stamp_b200.mil import VisionTransformer
    from stamp_b200.vit import TileEncoder, VitArch

    torch.manual_seed(0)
    cfg = vo.tiny_config(depth=2)
    arch = VitArch(cfg.name, patch=cfg.patch, dim=cfg.dim, depth=cfg.depth, heads=cfg.heads,
                   mlp_hidden=cfg.mlp_hidden, mlp=cfg.mlp, reg_tokens=cfg.reg_tokens)
    ext = Extractor(model=TileEncoder(arch, vo.make_weights(cfg), max_batch=32).to(cuda_device).eval(),
                    transform=pil_to_u8_hwc, identifier="tiny")
    slides = []
    for s in range(4):
        n = 10 + 7 * s
        cells = torch.randperm(64, generator=torch.Generator().manual_seed(s))[:n]
        coords = torch.stack([(cells % 8).float(), (cells // 8).float()], dim=-1) * 256.0
        slides.append((f"cohort/slide_{s}.svs", vo.synthetic_tiles(n, seed=s), coords))
    slides.append(("cohort/empty.svs", torch.zeros((0, 224, 224, 3), dtype=torch.uint8), torch.zeros((0, 2))))
    written = extract_to_feature_files(ext, slides, tmp_path, device=cuda_device, batch_size=16, code_hash="0a1b2c3d")
    assert [p.relative_to(tmp_path).as_posix() for p in written] == [f"tiny-0a1b2c3d/cohort/slide_{s}.h5" for s in range(4)]
    assert extract_to_feature_files(ext, slides, tmp_path, device=cuda_device, code_hash="0a1b2c3d") == []  # all exist
    direct = extract_slide_features(ext, slides[2][1], cuda_device, batch_size=16)
    with h5lite.File(written[2]) as h5:
        assert h5["feats"].dtype == torch.zeros(1).half().numpy().dtype and h5.attrs["extractor"] == "tiny"
        assert h5.attrs["tile_size_px"] == 224 and h5.attrs["feat_type"] == "tile"
        assert torch.equal(torch.from_numpy(h5["feats"][()]), direct)
        assert torch.equal(torch.from_numpy(features.get_coords(h5).coords_um), slides[2][2])
    cohort = features.load_cohort_to_device({"p0": written[:2], "p1": written[2:]}, cuda_device)
    assert cohort.feats.is_cuda and cohort.offsets == [0, 10 + 17, 10 + 17 + 24 + 31]
    f1, c1 = cohort.bag(1)
    assert torch.equal(f1[:24].cpu(), direct) and torch.equal(c1[:24].cpu(), slides[2][2])
    model = VisionTransformer(dim_output=2, dim_input=arch.dim, dim_model=128, n_layers=2, n_heads=2,
                              dim_feedforward=128, dropout=0.0, use_alibi=True).to(cuda_device).eval()
    probs = predict_bags(model, (cohort.bag(i) for i in range(len(cohort))), cuda_device)
    host = predict_bags(model, (features.read_bag(ps) for ps in (written[:2], written[2:])), cuda_device)
    assert probs.shape == (2, 2) and torch.allclose(probs, host, atol=1e-6)
