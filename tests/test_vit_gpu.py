"""Tile-encoder parity on the GPU: stamp_vit_forward (through the TileEncoder module) against the
fp32 CPU oracle on identical seeded synthetic tiles and weights.

Tolerance (north_star): feature vectors within 1e-3 relative, per tile: ||f - f_ref|| / ||f_ref||."""

import pytest
import torch

pytestmark = pytest.mark.gpu


def _per_tile_rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return ((a - b).norm(dim=1) / b.norm(dim=1)).max().item()


def _run(cfg_o, n_tiles, device, seed=3, max_batch=256):
    from oracle import vit_oracle as vo
    from stamp_b200.vit import TileEncoder, VitArch

    w = vo.make_weights(cfg_o, seed=1234)
    tiles = vo.synthetic_tiles(n_tiles, seed=seed, img=cfg_o.img)
    with torch.no_grad():
        ref = vo.forward(w, cfg_o, tiles)
    arch = VitArch(cfg_o.name, img=cfg_o.img, patch=cfg_o.patch, dim=cfg_o.dim, depth=cfg_o.depth,
                   heads=cfg_o.heads, mlp_hidden=cfg_o.mlp_hidden, mlp=cfg_o.mlp,
                   reg_tokens=cfg_o.reg_tokens, ln_eps=cfg_o.ln_eps, no_embed_class=cfg_o.no_embed_class,
                   mean=cfg_o.mean, std=cfg_o.std)
    enc = TileEncoder(arch, w, max_batch=max_batch).to(device).eval()
    out = enc(tiles.to(device))
    assert out.dtype == torch.float16 and out.shape == ref.shape
    assert torch.isfinite(out).all()
    return _per_tile_rel(out.float(), ref)


def test_tiny_no_embed_class_and_custom_normalisation(cuda_device):
    """UNI2-h / H-optimus style: position table on the patch tokens only, 8 register tokens, own mean / std."""
    from oracle import vit_oracle as vo

    cfg = vo.tiny_config(mlp="swiglu", reg_tokens=8, patch=14, depth=3, no_embed_class=True,
                         mean=vo.H_OPTIMUS.mean, std=vo.H_OPTIMUS.std)
    err = _run(cfg, 5, cuda_device)
    assert err < 1e-3, err


def test_uni2_and_h_optimus_blocks_match_oracle(cuda_device):
    """Full-width UNI2-h / H-optimus blocks (dim 1536, 24 heads of 64, SwiGLU 8192), two blocks deep."""
    from dataclasses import replace

    from oracle import vit_oracle as vo

    for full in (vo.UNI2, vo.H_OPTIMUS):
        err = _run(replace(full, depth=2), 4, cuda_device)
        assert err < 1e-3, (full.name, err)


@pytest.mark.parametrize("mlp,reg,patch", [("gelu", 0, 16), ("swiglu", 4, 14), ("gelu", 2, 14)])
def test_tiny_vit_matches_oracle(cuda_device, mlp, reg, patch):
    from oracle import vit_oracle as vo

    err = _run(vo.tiny_config(mlp=mlp, reg_tokens=reg, patch=patch, depth=3), 5, cuda_device)
    assert err < 1e-3, err


def test_tiny_vit_batch_chunking(cuda_device):
    from oracle import vit_oracle as vo

    err = _run(vo.tiny_config(depth=2), 7, cuda_device, max_batch=3)
    assert err < 1e-3, err


def test_uni_vit_l16_matches_oracle(cuda_device):
    """Full UNI architecture (ViT-L/16, 24 blocks) on 4 synthetic H&E-like tiles."""
    from oracle import vit_oracle as vo

    err = _run(vo.UNI, 4, cuda_device)
    print("ViT-L/16 max per-tile relative error:", err)
    assert err < 1e-3, err


def test_virchow2_vit_h14_matches_oracle(cuda_device):
    """Full Virchow2 architecture (ViT-H/14, 32 blocks, SwiGLU, 4 register tokens) on 2 tiles."""
    from oracle import vit_oracle as vo

    err = _run(vo.VIRCHOW2, 2, cuda_device)
    print("ViT-H/14 max per-tile relative error:", err)
    assert err < 1e-3, err


def _run_in_full_batch(cfg_o, n_oracle, batch, device):
    """The benched batch: `n_oracle` oracle-checked tiles spread over a batch of `batch` different tiles (the GEMM
    wave shape, CTA-pair tiles and attention grid of the bench), same < 1e-3 per-tile bound."""
    from oracle import vit_oracle as vo
    from stamp_b200.vit import TileEncoder, VitArch

    w = vo.make_weights(cfg_o, seed=1234)
    tiles = vo.synthetic_tiles(batch, seed=17, img=cfg_o.img)
    pos = torch.linspace(0, batch - 1, n_oracle).round().long()
    with torch.no_grad():
        ref = vo.forward(w, cfg_o, tiles[pos])
    arch = VitArch(cfg_o.name, img=cfg_o.img, patch=cfg_o.patch, dim=cfg_o.dim, depth=cfg_o.depth,
                   heads=cfg_o.heads, mlp_hidden=cfg_o.mlp_hidden, mlp=cfg_o.mlp,
                   reg_tokens=cfg_o.reg_tokens, ln_eps=cfg_o.ln_eps, no_embed_class=cfg_o.no_embed_class,
                   mean=cfg_o.mean, std=cfg_o.std)
    enc = TileEncoder(arch, w, max_batch=batch).to(device).eval()
    out = enc(tiles.to(device))
    assert out.shape == (batch, cfg_o.dim) and torch.isfinite(out).all()
    # every tile of the batch is a different image: no two feature rows may coincide (a mis-indexed batch would)
    assert torch.unique(out.float().cpu(), dim=0).shape[0] == batch
    return _per_tile_rel(out[pos.to(device)].float(), ref)


def test_uni_vit_l16_at_benched_batch_192(cuda_device):
    from oracle import vit_oracle as vo

    err = _run_in_full_batch(vo.UNI, 4, 192, cuda_device)
    print("ViT-L/16 @ batch 192, max per-tile relative error:", err)
    assert err < 1e-3, err


def test_uni_vit_l16_at_benched_batch_384(cuda_device):
    """The batch bench.py runs by default (296 row tiles of 256: four per CTA pair)."""
    from oracle import vit_oracle as vo

    err = _run_in_full_batch(vo.UNI, 4, 384, cuda_device)
    print("ViT-L/16 @ batch 384, max per-tile relative error:", err)
    assert err < 1e-3, err


def test_virchow2_vit_h14_at_benched_batch_96(cuda_device):
    from oracle import vit_oracle as vo

    err = _run_in_full_batch(vo.VIRCHOW2, 2, 96, cuda_device)
    print("ViT-H/14 @ batch 96, max per-tile relative error:", err)
    assert err < 1e-3, err


def test_virchow_full_cls_and_mean_patch_token(cuda_device):
    """virchow_full.py:24-35: [class token | mean of the other tokens] after the final norm (2 x dim values), Virchow-v1
    shaped blocks (patch 14, SwiGLU, no register tokens) three deep at full width; and the class-token half equals the
    plain class-token extractor's output of the same weights."""
    from dataclasses import replace

    from oracle import vit_oracle as vo
    from stamp_b200.vit import VIRCHOW_ARCH, VIRCHOW_FULL_ARCH, TileEncoder

    cfg = vo.VitConfig("virchow-d3", patch=14, dim=1280, depth=3, heads=16, mlp_hidden=6832, mlp="swiglu")
    w = vo.make_weights(cfg, seed=1234)
    tiles = vo.synthetic_tiles(5, seed=8)
    with torch.no_grad():
        tok = vo.forward_tokens(w, cfg, vo.transform_u8(tiles, torch.float32, cfg.mean, cfg.std))
        ref = torch.cat([tok[:, 0], tok[:, 1:].mean(1)], dim=-1)
    full = TileEncoder(replace(VIRCHOW_FULL_ARCH, depth=3), w, max_batch=3).to(cuda_device).eval()
    out = full(tiles.to(cuda_device))
    assert out.shape == (5, 2560) and out.dtype == torch.float16 and torch.isfinite(out).all()
    assert _per_tile_rel(out[:, :1280].float(), ref[:, :1280]) < 1e-3
    assert _per_tile_rel(out[:, 1280:].float(), ref[:, 1280:]) < 1e-3
    cls = TileEncoder(replace(VIRCHOW_ARCH, depth=3), w, max_batch=8).to(cuda_device).eval()(tiles.to(cuda_device))
    # (the class-token-only encoder skips the other tokens in its last block: same arithmetic on the rows it keeps)
    assert _per_tile_rel(cls.float(), out[:, :1280].float()) < 2e-4
    assert full.launches_per_batch() == 3 + 7 * 3 + 1 + 1


def test_dinobloom_vit_s14_matches_oracle(cuda_device):
    """dinobloom.py:30-78: DINOv2 ViT-S/14 at full depth (dim 384 = one and a half 256-column GEMM tiles, 6 heads of 64,
    257 tokens), weights handed over with the checkpoint's ``backbone.`` prefix."""
    from oracle import vit_oracle as vo
    from stamp_b200.extractor import dino_bloom

    cfg = vo.VitConfig("dinobloom", patch=14, dim=384, depth=12, heads=6, mlp_hidden=1536)
    w = vo.make_weights(cfg, seed=1234)
    tiles = vo.synthetic_tiles(6, seed=12)
    with torch.no_grad():
        ref = vo.forward(w, cfg, tiles)
    teacher = {f"backbone.{k}": v for k, v in w.items()}
    teacher["dino_head.mlp.0.weight"] = torch.zeros(4, 4)
    ext = dino_bloom(teacher, max_batch=4)
    assert ext.identifier == "dino-bloom"
    out = ext.model.to(cuda_device).eval()(tiles.to(cuda_device))
    assert out.shape == (6, 384) and _per_tile_rel(out.float(), ref) < 1e-3


def test_tile_encoder_refuses_cpu():
    from oracle import vit_oracle as vo
    from stamp_b200.vit import TileEncoder, VitArch

    cfg = vo.tiny_config(depth=1)
    enc = TileEncoder(VitArch("t", dim=cfg.dim, depth=1, heads=cfg.heads, mlp_hidden=cfg.mlp_hidden),
                      vo.make_weights(cfg))
    with pytest.raises(RuntimeError):
        enc(torch.zeros(1, 224, 224, 3, dtype=torch.uint8))


def test_macenko_then_virchow2_chain_matches_oracle_chain(cuda_device):
    """BASELINE configs[2] pipeline on one batch: Macenko over the tile batch -> Virchow2-style
    encoder (patch 14, SwiGLU, 4 register tokens; depth reduced for CPU-oracle speed)."""
    import numpy as np

    from oracle import macenko_oracle as mo
    from oracle import vit_oracle as vo
    from stamp_b200.macenko import macenko_normalize
    from stamp_b200.vit import TileEncoder, VitArch

    cfg = vo.VitConfig("virchow2-d3", patch=14, dim=256, depth=3, heads=4, mlp_hidden=1376, mlp="swiglu", reg_tokens=4)
    w = vo.make_weights(cfg, seed=99)
    tiles = vo.synthetic_tiles(6, seed=41)
    norm_ref, *_ = mo.normalize(tiles.numpy())
    with torch.no_grad():
        feats_ref = vo.forward(w, cfg, torch.from_numpy(norm_ref))
    arch = VitArch(cfg.name, patch=14, dim=256, depth=3, heads=4, mlp_hidden=1376, mlp="swiglu", reg_tokens=4)
    enc = TileEncoder(arch, w).to(cuda_device).eval()
    norm = macenko_normalize(tiles.to(cuda_device))
    feats = enc(norm)
    assert np.abs(norm.cpu().numpy().astype(np.int16) - norm_ref.astype(np.int16)).max() <= 1
    # features after a <= 1 LSB difference on <2 % of the input bytes + fp16 operands
    assert _per_tile_rel(feats.float(), feats_ref) < 2e-3
