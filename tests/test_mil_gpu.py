"""MIL aggregator parity on the GPU, through the drop-in VisionTransformer module (C-ABI underneath).

1. Golden fixtures = outputs of the reference module itself on seeded inputs (tests/golden), loaded
   through load_state_dict -> also pins state-dict compatibility with reference checkpoints.
2. Default-size model (1024 -> 512, 8 heads, 2 layers, ff 512) against the CPU oracle at
   N in {64, 512, 4096}.
Tolerance (north_star): logits within 1e-3 relative (||d|| / ||ref|| per bag)."""

import numpy as np
import pytest
import torch

from test_oracle_cpu import GOLDEN, GOLDEN_LONG, load_golden, load_golden_long

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


@pytest.mark.parametrize("path", GOLDEN_LONG, ids=[p.stem for p in GOLDEN_LONG])
def test_mil_long_bag_matches_reference_golden(cuda_device, path):
    """The benched shape (one bag of 4096 x 1024, default-size model, trained running mean) against the logits
    of the reference module itself: pins the tcgen05 long-bag attention kernel to vision_tranformer.py:42-74
    (ALiBi) and :218-228 (nn.MultiheadAttention) directly, not through the oracle."""
    sd, bags, coords, ref_logits = load_golden_long(path)
    model = _model_from_sd(sd, 8, cuda_device)
    with torch.inference_mode():
        out = model(bags.to(cuda_device), coords=coords.to(cuda_device), mask=None)
    err = _rel_per_bag(out, ref_logits)
    print(path.stem, "max per-bag relative error vs the reference module", err)
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


@pytest.mark.parametrize("use_alibi", [True, False])
def test_long_bag_attention_kernel_variants_agree(cuda_device, use_alibi):
    """Long bags run on the third-generation tcgen05 kernel (attention_mil_v3.cu: distance tiles by TMA, P through
    tensor memory, lazy accumulator rescaling; mode 1).  Bit 3 (modes 9, 41) forces a rescale whenever a row
    maximum grows (exercises the TMEM read-modify-write path), bit 5 (33, 41) selects the second generation, an
    independent implementation that recomputes the distances in the kernel; all must match the oracle.  Logit scales
    are blown up so that maxima do grow along the bag."""
    from oracle import mil_oracle
    from stamp_b200 import _lib

    sd = mil_oracle.init_state_dict(dim_input=1024, dim_output=3, use_alibi=use_alibi, seed=12, running_mean=9000.0)
    for k in list(sd):          # peaky attention: row maxima keep growing, the lazy threshold (2^8) is crossed
        if "query_encoders" in k and k.endswith("weight"):
            sd[k] = sd[k] * 6.0
        if k.endswith(".in_proj_weight"):
            sd[k] = torch.cat([sd[k][:512] * 6.0, sd[k][512:]])      # the query rows only
    bags, coords = mil_oracle.synthetic_bag(1200, 1024, seed=5, batch=2)
    order = bags[0].norm(dim=1).argsort()           # ascending norms: later tiles tend to score higher
    bags[0], coords[0] = bags[0][order], coords[0][order]
    with torch.no_grad():
        ref = mil_oracle.forward(sd, bags, coords, None, n_heads=8)
    model = _model_from_sd(sd, 8, cuda_device)
    lib = _lib.load()
    outs = {}
    try:
        for mode in (1, 9, 33, 41):
            lib.stamp_b200_attention_tc_enable(mode)
            with torch.inference_mode():
                outs[mode] = model(bags.to(cuda_device), coords=coords.to(cuda_device), mask=None).cpu()
    finally:
        lib.stamp_b200_attention_tc_enable(1)
    for mode, out in outs.items():
        err = _rel_per_bag(out, ref)
        print(f"alibi={use_alibi} mode={mode}: max per-bag relative error {err:.2e}")
        # fp16 q / k with 6x larger logits: the softmax side is noisier than in the 1e-3 default-size tests
        assert err < (1e-3 if use_alibi else 4e-3), (mode, err)
    for mode in (9, 33, 41):       # the kernels agree with each other far below that
        assert _rel_per_bag(outs[mode], outs[1]) < 3e-4, mode


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
    # margin-checked fixture: the largest k <= 10 whose ranks 1..k+1 are separated by more than the numerical
    # noise of the scores; the margin itself is asserted (a fixture without one would not test the contract)
    gap = pr.sort(descending=True).values
    margins = gap[:11] - gap[1:12]
    k = max(kk for kk in range(1, 11) if margins[:kk].min() > 1e-4)
    assert k >= 5, f"fixture has no score margin among its top tiles: {margins.tolist()}"
    assert torch.equal(pr.topk(k).indices, po.topk(k).indices)


@pytest.mark.parametrize("use_alibi", [False, True])
@pytest.mark.parametrize("dims", [dict(dim_input=456, n_heads=4, head=33, dim_ff=135, n_layers=3, batch=6, n=75, C=3),
                                  dict(dim_input=457, n_heads=5, head=34, dim_ff=135, n_layers=3, batch=7, n=76, C=4),
                                  dict(dim_input=25, n_heads=8, head=64, dim_ff=512, n_layers=2, batch=3, n=300, C=2),
                                  dict(dim_input=64, n_heads=3, head=20, dim_ff=96, n_layers=1, batch=2, n=40, C=2)])
def test_mil_shapes_outside_the_kernel_envelope_run_zero_padded(cuda_device, dims, use_alibi):
    """The reference's own unit tests build models the kernels have no native shape for (tests/test_model.py: heads of
    33 / 34 columns, 456 / 457 input features, 135 hidden units; tests/test_train_deploy.py: 25 input features).  They
    run with zero-padded weights (LayerNorm statistics and softmax scale on the real sizes), masked and unmasked,
    and match the oracle like any other shape."""
    from oracle import mil_oracle

    H, hd = dims["n_heads"], dims["head"]
    # (ALiBi: two layers, like every other 1e-3 check of that variant -- its unscaled distance term costs accuracy per
    #  layer; the reference's odd-shape tests are use_alibi=False, three layers)
    sd = mil_oracle.init_state_dict(dim_input=dims["dim_input"], dim_output=dims["C"], dim_model=H * hd, n_heads=H,
                                    n_layers=min(dims["n_layers"], 2) if use_alibi else dims["n_layers"],
                                    dim_feedforward=dims["dim_ff"], use_alibi=use_alibi,
                                    seed=31, running_mean=3000.0)
    bags, coords = mil_oracle.synthetic_bag(dims["n"], dims["dim_input"], seed=7, batch=dims["batch"])
    g = torch.Generator().manual_seed(3)
    mask = torch.arange(dims["n"])[None, :] >= torch.randint(1, dims["n"], (dims["batch"], 1), generator=g)
    model = _model_from_sd(sd, H, cuda_device)
    for m in (None, mask):
        with torch.no_grad():
            ref = mil_oracle.forward(sd, bags, coords, m, n_heads=H)
            out = model(bags.to(cuda_device), coords=coords.to(cuda_device), mask=None if m is None else m.to(cuda_device))
        assert out.shape == (dims["batch"], dims["C"])
        err = _rel_per_bag(out, ref)
        print(f"alibi={use_alibi} {dims} masked={m is not None}: max per-bag relative error {err:.2e}")
        assert err < 1e-3, err
    with torch.inference_mode():      # determinism of two evaluations (the reference's test_inference_reproducibility)
        a = model(bags.to(cuda_device), coords=coords.to(cuda_device), mask=mask.to(cuda_device))
        b = model(bags.to(cuda_device), coords=coords.to(cuda_device), mask=mask.to(cuda_device))
    assert torch.equal(a, b)


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
    # gradients exist for the mask=None branch only (tests/test_mil_train_gpu.py); masked forwards with
    # autograd enabled must refuse rather than silently drop the graph
    with pytest.raises(NotImplementedError):
        m(x.to(cuda_device), coords=c.to(cuda_device), mask=torch.zeros(1, 5, dtype=torch.bool, device=cuda_device))


def test_mil_full_scale_permutation_invariance(cuda_device):
    """BASELINE configs[4] scale (50k-tile bag, S = 50 001): the dense reference cannot run this with
    ALiBi (>= 50 GB of S x S temporaries), so parity is checked through a size-independent property:
    the aggregator is a set function of (feature, coordinate) pairs -- permuting the tiles must not
    change the logits beyond summation-order noise -- and spot rows against a blocked fp64 restatement."""
    from oracle import mil_oracle

    n = 50_000
    sd = mil_oracle.init_state_dict(dim_input=768, dim_output=2, seed=21, running_mean=9000.0)
    g = torch.Generator().manual_seed(5)
    feats = torch.randn(1, n, 768, generator=g).half().float()
    cells = torch.randperm(250 * 200, generator=g)[:n]
    coords = torch.stack([(cells % 250).float(), (cells // 250).float()], -1)[None] * 256.0
    model = _model_from_sd(sd, 8, cuda_device)
    perm = torch.randperm(n, generator=g)
    with torch.inference_mode():
        a = model(feats.to(cuda_device), coords=coords.to(cuda_device), mask=None)
        b = model(feats[:, perm].to(cuda_device), coords=coords[:, perm].to(cuda_device), mask=None)
    assert torch.isfinite(a).all()
    assert _rel_per_bag(a, b.cpu()) < 1e-3

    # spot check: first-layer ALiBi attention output of the class token and two tiles, one head, in fp64
    d, H, hd = 512, 8, 64
    x = torch.nn.functional.gelu(torch.nn.functional.linear(
        feats[0].double(), sd["project_features.0.weight"].double(), sd["project_features.0.bias"].double()))
    x = torch.cat([sd["class_token"].double()[None], x])
    c = torch.cat([torch.zeros(1, 2, dtype=torch.float64), coords[0].double()])
    p = "transformer.layers.0.0."
    xn = torch.nn.functional.layer_norm(x, (d,), sd[p + "norm.weight"].double(), sd[p + "norm.bias"].double(), 1e-5)
    h = 3
    lin = lambda name: torch.nn.functional.linear(xn, sd[p + f"mhsa.{name}_encoders.{h}.weight"].double(),
                                                  sd[p + f"mhsa.{name}_encoders.{h}.bias"].double())
    q, k, v = lin("query"), lin("key"), lin("value")
    slope = (sd[p + f"mhsa.attentions.{h}.bias_scale"] / sd[p + f"mhsa.attentions.{h}.scale_distance.running_mean"]).double()
    rows = torch.tensor([0, 1, 31337])
    w = torch.softmax(q[rows] @ k.T / 8.0, -1) - slope * (c[rows, None] - c[None]).norm(dim=-1)
    ref = w @ v

    from stamp_b200 import ops
    qkv = torch.cat([torch.cat([lin_(xn) for lin_ in [
        (lambda t, nm=nm, hh=hh: torch.nn.functional.linear(t, sd[p + f"mhsa.{nm}_encoders.{hh}.weight"].double(),
                                                             sd[p + f"mhsa.{nm}_encoders.{hh}.bias"].double()))
        for hh in range(H)]], -1) for nm in ("query", "key", "value")], -1)
    slopes = torch.cat([sd[p + f"mhsa.attentions.{hh}.bias_scale"] / sd[p + f"mhsa.attentions.{hh}.scale_distance.running_mean"]
                        for hh in range(H)]).float()
    out = ops.attention(qkv[None].to(cuda_device, torch.float16), H, coords=c[None].float().to(cuda_device).contiguous(),
                        slope=slopes.to(cuda_device))
    got = out[0, rows][:, h * hd:(h + 1) * hd].double().cpu()
    assert ((got - ref).norm() / ref.norm()).item() < 2e-3  # fp16 q/k/v operands at S = 50 001


def test_predict_bags_pipeline_matches_per_bag_forward(cuda_device):
    """deploy._predict mirror: ragged bags, fp16 (feature-file dtype) and fp32 inputs, probabilities per patient."""
    from oracle import mil_oracle
    from stamp_b200.deploy import predict_bags, predict_patients

    sd = mil_oracle.init_state_dict(dim_input=64, dim_output=3, dim_model=128, n_heads=2, dim_feedforward=128, seed=2,
                                    running_mean=7000.0)
    model = _model_from_sd(sd, 2, cuda_device)
    sizes = [5, 300, 64, 1, 129, 700, 33]
    bags = [mil_oracle.synthetic_bag(n, 64, seed=40 + n) for n in sizes]
    host = [(f[0].half() if i % 2 == 0 else f[0], c[0]) for i, (f, c) in enumerate(bags)]   # mixed dtypes
    probs = predict_bags(model, iter(host), cuda_device)
    assert probs.shape == (len(sizes), 3) and torch.allclose(probs.sum(1), torch.ones(len(sizes)), atol=1e-5)
    for i, (f, c) in enumerate(bags):
        ref = torch.softmax(mil_oracle.forward(sd, f, c, None), dim=1)[0]
        assert torch.allclose(probs[i], ref, atol=2e-4), (i, probs[i], ref)
    named = predict_patients(model, [f"p{i}" for i in range(len(sizes))], iter(host), cuda_device)
    # (predict_patients batches the bags raggedly: short bags meet another attention kernel than in predict_bags)
    assert list(named) == [f"p{i}" for i in range(len(sizes))]
    for i, (f, c) in enumerate(bags):
        ref = torch.softmax(mil_oracle.forward(sd, f, c, None), dim=1)[0]
        assert torch.allclose(named[f"p{i}"], ref, atol=2e-4) and torch.allclose(named[f"p{i}"], probs[i], atol=2e-4)
    assert predict_bags(model, iter([]), cuda_device).shape == (0, 3)
    with pytest.raises(RuntimeError):
        predict_bags(model, iter(host), "cpu")


@pytest.mark.parametrize("use_alibi", [True, False])
def test_ragged_batch_equals_per_bag_forwards(cuda_device, use_alibi):
    """Bags of different lengths through ONE forward (stamp_mil_forward_ragged): the dense layers run over the rows of
    all bags, the long-bag attention kernel over each bag; rows and bags are independent, so every bag's logits equal
    its own batch-1 forward bit for bit (lengths on both sides of the tile sizes, 1-tile bags, > 256 and <= 256)."""
    from stamp_b200.deploy import predict_bags, predict_bags_ragged
    from stamp_b200.mil import VisionTransformer

    torch.manual_seed(7)
    model = VisionTransformer(dim_output=3, dim_input=96, dim_model=128, n_layers=2, n_heads=2, dim_feedforward=256,
                              dropout=0.0, use_alibi=use_alibi).to(cuda_device).eval()
    if use_alibi:
        for layer in model.transformer.layers:
            for a in layer[0].mhsa.attentions:
                a.scale_distance.running_mean.fill_(3000.0)
    assert model.supports_ragged()
    g = torch.Generator().manual_seed(1)
    lengths = [700, 1, 63, 64, 65, 127, 128, 129, 1000, 300, 255, 256, 257, 2, 513]
    bags = [(torch.randn(n, 96, generator=g).half(), torch.rand(n, 2, generator=g) * 20000) for n in lengths]
    tokens, coords, seq, s_max = model.pack_ragged(bags, pin=False)
    assert s_max == 1001 and seq.tolist()[-1] == sum(lengths) + len(lengths)
    with torch.inference_mode():
        ragged = model.forward_ragged(tokens.to(cuda_device), coords.to(cuda_device), seq.to(cuda_device), s_max)
        single = torch.cat([model(f.to(cuda_device)[None], coords=c.to(cuda_device)[None], mask=None) for f, c in bags])
    assert ragged.shape == (len(lengths), 3) and torch.isfinite(ragged).all()
    # (bags of <= 256 tokens take another attention kernel in the batch-1 path: same math, different tile order)
    long = torch.tensor([n + 1 > 256 for n in lengths], device=cuda_device)
    assert torch.equal(ragged[long], single[long])
    err = ((ragged - single).norm(dim=1) / single.norm(dim=1)).max().item()
    assert err < 1e-3, err
    # the deploy loops: ragged batches (several groups: max_rows 1500) == per-bag streams, host and device bags
    p_ref = predict_bags(model, iter(bags), cuda_device, graphs=False)
    p_rag = predict_bags_ragged(model, iter(bags), cuda_device, max_rows=1500)
    # (probabilities: identical for the long bags, within the kernels' tolerance for the short ones, see above)
    longh = long.cpu()
    assert p_rag.shape == p_ref.shape and torch.equal(p_rag[longh], p_ref[longh]) and torch.allclose(p_rag, p_ref, atol=5e-4)
    dev_bags = [(f.to(cuda_device), c.to(cuda_device)) for f, c in bags]
    p_dev = predict_bags_ragged(model, iter(dev_bags), cuda_device, max_rows=4000, max_bags=4)
    assert torch.equal(p_dev[longh], p_ref[longh]) and torch.allclose(p_dev, p_ref, atol=5e-4)
