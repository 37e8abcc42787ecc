"""MIL training step on the GPU (C-ABI: stamp_mil_train_forward/backward, stamp_cross_entropy,
stamp_adamw_step, stamp_pairwise_dist_mean) against

1. the reference's own training step (tests/golden/mil_train_step.npz, written from the reference
   module by oracle/make_golden_train.py): logits, loss, every parameter gradient, running means;
2. the fp64 oracle with the kernels' own dropout masks (dropout sites active, p = 0.25 / 0.5);
3. the fp64 oracle at the default model size (1024 -> 512, 8 heads, 2 layers), ragged bag lengths;
4. torch.optim.AdamW / F.cross_entropy / torch.cdist as checkers of the small kernels.

Tolerance: the training path computes in bf16 (BASELINE.json configs[3] "training bf16"; 8-bit
significand) with fp32 accumulation: logits within 2e-2 relative, per-tensor gradients within
5e-2 of the tensor norm and cosine > 0.998 against fp32/fp64 references."""

import numpy as np
import pytest
import torch

from oracle import mil_oracle
from test_mil_train_cpu import TRAIN_GOLDENS, load_train_golden

pytestmark = pytest.mark.gpu

LOGIT_TOL = 2e-2
GRAD_TOL = 5e-2
COS_TOL = 0.998
VERBOSE = bool(int(__import__('os').environ.get('STAMP_TEST_VERBOSE', '0')))


def _model(sd, n_heads, device, dropout=0.0, p_ff=0.0):
    from stamp_b200.mil import VisionTransformer

    d_model, d_in = sd["project_features.0.weight"].shape
    n_layers = 1 + max(int(k.split(".")[2]) for k in sd if k.startswith("transformer.layers."))
    m = VisionTransformer(dim_output=sd["mlp_head.0.weight"].shape[0], dim_input=d_in, dim_model=d_model,
                          n_layers=n_layers, n_heads=n_heads,
                          dim_feedforward=sd["transformer.layers.0.1.1.weight"].shape[0], dropout=dropout,
                          use_alibi=any(".query_encoders." in k for k in sd))
    m.load_state_dict(sd, strict=True)
    for _, ff in m.transformer.layers:
        ff[3].p = p_ff
        ff[5].p = p_ff
    return m.to(device).train()


def _check_grads(model, ref_grads, what, grad_tol=GRAD_TOL, term_scale=None):
    """term_scale: per-tensor magnitude of the per-bag terms a gradient is the sum of -- the floor of the relative error for
    losses whose per-bag cotangents cancel (the Cox loss: they sum to zero)."""
    worst = (0.0, None)
    named = {k: p.grad for k, p in model.named_parameters()}
    assert set(named) == set(ref_grads)
    assert all(g is not None for g in named.values())
    # the H scalar bias_scale parameters of a layer are compared as one [H] vector: each is a signed sum
    # over all query/key pairs, and a head whose sum happens to cancel has no meaningful relative error
    ours, refs = {}, {}
    for k, ref in ref_grads.items():
        if k.endswith(".bias_scale"):
            key = k.split(".mhsa.")[0] + ".mhsa.attentions.*.bias_scale"
            ours.setdefault(key, []).append(named[k].double().cpu().flatten())
            refs.setdefault(key, []).append(ref.double().flatten())
        else:
            ours[k], refs[k] = [named[k].double().cpu().flatten()], [ref.double().flatten()]
    scale = max(float(v.norm()) for v in ref_grads.values())
    failures = []
    for k in refs:
        g, ref = torch.cat(ours[k]), torch.cat(refs[k])
        err = float((g - ref).norm())
        floor = 1e-5 * scale
        if term_scale is not None and k in term_scale:
            floor = max(floor, term_scale[k])
        if ".key_encoders." in k and k.endswith(".bias"):
            # analytically zero (softmax is invariant to the per-query shift a key bias adds): both sides
            # hold round-off only; bound it by the same head's key-weight gradient
            floor = float(ref_grads[k[:-4] + "weight"].norm())
        rel = err / max(float(ref.norm()), floor)
        cos = float(torch.dot(g, ref) / (g.norm() * ref.norm())) if float(ref.norm()) > 10 * floor else 1.0
        if VERBOSE:
            print(f"    {rel:.3e} cos {cos:.5f} {k}" + (f"  ours {g.tolist()} ref {ref.tolist()}" if g.numel() <= 8 else ""))
        if cos <= COS_TOL or rel >= grad_tol:
            failures.append((k, rel, cos))
        if rel > worst[0]:
            worst = (rel, k)
    assert not failures, (what, failures)
    print(f"[{what}] worst per-tensor gradient error {worst[0]:.3e} ({worst[1]})")


@pytest.mark.parametrize("name", TRAIN_GOLDENS)
def test_training_step_matches_reference_golden(cuda_device, name):
    from stamp_b200 import train as T

    g = load_train_golden(name)
    model = _model(g["sd"], g["n_heads"], cuda_device)
    bags, coords = g["bags"].to(cuda_device), g["coords"].to(cuda_device)
    batch = (bags, coords, None, g["targets"].to(cuda_device))
    loss = T.training_step(model, batch, g["class_weights"].to(cuda_device))
    loss.backward()
    # running means: the reference's training-mode side effect
    for k, v in model.state_dict().items():
        if "scale_distance" in k:
            assert torch.allclose(v.cpu(), g["after"][k], rtol=1e-4), k
    with torch.no_grad():
        model.eval()
        # same weights, updated running means, through the fp16 inference path: sanity of the golden itself
        ev = model(bags, coords=coords, mask=None).cpu()
        model.train()
    rel = ((ev - g["logits"]).norm(dim=1) / g["logits"].norm(dim=1)).max()
    assert rel < 1e-3, float(rel)
    loss = loss.detach()
    assert abs(float(loss) - float(g["loss"])) < LOGIT_TOL * max(1.0, float(g["loss"]))
    print(f"loss {float(loss):.5f} vs reference {float(g['loss']):.5f}")
    _check_grads(model, g["grads"], name)


def test_train_forward_logits_match_golden(cuda_device):
    g = load_train_golden()
    model = _model(g["sd"], g["n_heads"], cuda_device)
    out = model(g["bags"].to(cuda_device), coords=g["coords"].to(cuda_device), mask=None)
    assert out.requires_grad
    rel = ((out.detach().cpu() - g["logits"]).norm(dim=1) / g["logits"].norm(dim=1)).max()
    print(f"training-forward logits rel err {float(rel):.3e}")
    assert rel < LOGIT_TOL


def test_dropout_sites_match_oracle_with_kernel_masks(cuda_device):
    from stamp_b200 import train as T

    g = load_train_golden()
    p_proj, p_ff = 0.25, 0.5
    model = _model(g["sd"], g["n_heads"], cuda_device, dropout=p_proj, p_ff=p_ff)
    bags, coords = g["bags"].to(cuda_device), g["coords"].to(cuda_device)
    torch.manual_seed(31337)
    loss = T.training_step(model, (bags, coords, None, g["targets"].to(cuda_device)), g["class_weights"].to(cuda_device))
    loss.backward()
    torch.manual_seed(31337)
    seed = int(torch.randint(0, 2 ** 62, (1,)).item())
    B, n, _ = bags.shape
    d, ff, M = 128, 128, B * (n + 1)
    masks = {0: T.dropout_keep_mask(seed, 0, B * n * d, p_proj, cuda_device).cpu().bool()}
    for l in range(2):
        masks[1 + 2 * l] = T.dropout_keep_mask(seed, 1 + 2 * l, M * ff, p_ff, cuda_device).cpu().bool()
        masks[2 + 2 * l] = T.dropout_keep_mask(seed, 2 + 2 * l, M * d, p_ff, cuda_device).cpu().bool()
    assert abs(float(masks[0].float().mean()) - (1 - p_proj)) < 0.01
    assert abs(float(masks[1].float().mean()) - (1 - p_ff)) < 0.01
    assert not torch.equal(masks[1], masks[3])   # sites draw independent masks
    _, ref_loss, ref_grads, _ = mil_oracle.train_grads(g["sd"], g["bags"], g["coords"], g["targets"],
                                                       g["class_weights"], drop_masks=masks, p_proj=p_proj, p_ff=p_ff)
    assert abs(float(loss) - float(ref_loss)) < LOGIT_TOL * max(1.0, float(ref_loss))
    _check_grads(model, ref_grads, "dropout")
    # a different seed gives a different loss (the masks are live)
    model.zero_grad()
    torch.manual_seed(1)
    loss2 = T.training_step(model, (bags, coords, None, g["targets"].to(cuda_device)), g["class_weights"].to(cuda_device))
    assert abs(float(loss2) - float(loss)) > 1e-4


@pytest.mark.parametrize("n_tiles,batch", [(1, 2), (63, 1), (64, 2), (65, 1), (300, 1), (512, 2), (777, 1)])
def test_default_size_gradients_match_oracle(cuda_device, n_tiles, batch):
    from stamp_b200 import train as T

    sd = mil_oracle.init_state_dict(dim_input=1024, dim_output=2, seed=3)
    bags, coords = mil_oracle.synthetic_bag(n_tiles, 1024, seed=10 + n_tiles, batch=batch, signal=True)
    targets = torch.nn.functional.one_hot(torch.arange(batch) % 2, 2).float()
    model = _model(sd, 8, cuda_device)
    loss = T.training_step(model, (bags.to(cuda_device), coords.to(cuda_device), None, targets.to(cuda_device)), None)
    loss.backward()
    _, ref_loss, ref_grads, _ = mil_oracle.train_grads(sd, bags, coords, targets, None)
    assert abs(float(loss) - float(ref_loss)) < LOGIT_TOL * max(1.0, float(ref_loss)), (float(loss), float(ref_loss))
    _check_grads(model, ref_grads, f"default N={n_tiles}")


def test_benched_shape_gradients_match_oracle(cuda_device):
    """The shape bench.py and BASELINE configs[3] run: 8 bags of 4096 x 1024 per GPU, default model.  All eight bags
    go through the kernels; ONE of them carries the loss at a time (the other targets are all-zero rows, whose
    cross-entropy term and gradient vanish exactly), so the fp32 CPU oracle differentiates one 4097-token bag per
    case: the first bag (class 0) and the sixth (class 1).  The running mean is updated from the distances of all
    eight bags, as the reference does.  In a third case both bags carry the loss: the cross-entropy gradient is
    linear in the target rows, so the oracle gradient is the sum of the two single-bag ones.  With opposite labels
    the ALiBi-dominated per-bag gradients nearly cancel in every bias-like sum (measured: the net is ~1/10 of the
    parts), which amplifies bf16 noise relative to the NET gradient by that factor -- for the reference under bf16
    autocast just the same -- so this case is measured against the scale of what is summed."""
    from stamp_b200 import train as T

    B, N = 8, 4096
    label = {0: 0, 5: 1}
    sd = mil_oracle.init_state_dict(dim_input=1024, dim_output=2, seed=5)
    bags, coords = mil_oracle.synthetic_bag(N, 1024, seed=4242, batch=B, signal=True)
    sd2 = mil_oracle.running_mean_update(sd, coords)
    model = _model(sd, 8, cuda_device)
    oracle = {}
    for live in [(0,), (5,), (0, 5)]:
        targets = torch.zeros(B, 2)
        for b in live:
            targets[b, label[b]] = 1.0
        model.load_state_dict(sd)          # the training-mode forward moves the running means: start over
        model.zero_grad()
        loss = T.training_step(model, (bags.to(cuda_device), coords.to(cuda_device), None, targets.to(cuda_device)), None)
        loss.backward()
        torch.cuda.synchronize()
        for k, v in model.state_dict().items():
            if "scale_distance" in k:
                assert torch.allclose(v.cpu(), sd2[k], rtol=1e-4), k
        if len(live) == 1:
            b = live[0]
            params = {k: v.detach().clone().requires_grad_(True) for k, v in sd2.items() if "scale_distance" not in k}
            # one bag at a time: the S x S intermediates of one bag are ~5 GB in fp32
            logits = mil_oracle.forward({**sd2, **params}, bags[b:b + 1], coords[b:b + 1], None, exact_dist=True)
            term = mil_oracle.cross_entropy(logits, targets[b:b + 1], None) / B
            term.backward()
            oracle[b] = ({k: v.grad.detach() for k, v in params.items()}, float(term.detach()))
            ref_grads, ref_loss = oracle[b]
            assert abs(float(loss.detach()) - ref_loss) < LOGIT_TOL * max(1.0, ref_loss), (float(loss.detach()), ref_loss)
            _check_grads(model, ref_grads, f"benched shape 8 x 4096, bag {b} carries the loss")
        else:
            (g0, l0), (g5, l5) = oracle[0], oracle[5]
            assert abs(float(loss.detach()) - (l0 + l5)) < LOGIT_TOL * max(1.0, l0 + l5)
            worst = 0.0
            for k, p in model.named_parameters():
                ref = (g0[k] + g5[k]).double()
                err = float((p.grad.double().cpu() - ref).norm())
                scale = float(g0[k].double().norm() + g5[k].double().norm())
                if ".key_encoders." in k and k.endswith(".bias"):
                    scale = float(g0[k[:-4] + "weight"].double().norm() + g5[k[:-4] + "weight"].double().norm())
                worst = max(worst, err / max(scale, 1e-30))
                assert err < GRAD_TOL * scale, (k, err / scale)
            print(f"[benched shape, two bags with opposite labels] worst error / summed scale {worst:.3e}")


@pytest.mark.parametrize("use_alibi", [False, True])
def test_input_width_outside_the_envelope_trains_zero_padded(cuda_device, use_alibi):
    """tests/test_train_deploy.py of the reference trains on 25-dimensional features: the training path pads the
    input width (bags and projection weight) with zeros; every gradient, including the projection's, matches."""
    from stamp_b200 import train as T

    sd = mil_oracle.init_state_dict(dim_input=25, dim_output=2, dim_model=128, n_heads=2, dim_feedforward=128, seed=6,
                                    use_alibi=use_alibi)
    bags, coords = mil_oracle.synthetic_bag(32, 25, seed=9, batch=8, signal=True)
    targets = torch.nn.functional.one_hot(torch.arange(8) % 2, 2).float()
    model = _model(sd, 2, cuda_device)
    loss = T.training_step(model, (bags.to(cuda_device), coords.to(cuda_device), None, targets.to(cuda_device)), None)
    loss.backward()
    _, ref_loss, ref_grads, _ = mil_oracle.train_grads(sd, bags, coords, targets, None, n_heads=2)
    assert abs(float(loss.detach()) - float(ref_loss)) < LOGIT_TOL * max(1.0, float(ref_loss))
    assert model.project_features[0].weight.grad.shape == (128, 25)
    _check_grads(model, ref_grads, f"dim_input 25, alibi={use_alibi}")


def test_default_size_mha_variant_matches_oracle(cuda_device):
    """use_alibi=False (the reference's default backbone): long bag -> tcgen05 forward / backward, plain softmax."""
    from stamp_b200 import train as T

    sd = mil_oracle.init_state_dict(dim_input=1024, dim_output=2, seed=4, use_alibi=False)
    bags, coords = mil_oracle.synthetic_bag(400, 1024, seed=21, batch=2, signal=True)
    targets = torch.nn.functional.one_hot(torch.arange(2) % 2, 2).float()
    model = _model(sd, 8, cuda_device)
    loss = T.training_step(model, (bags.to(cuda_device), coords.to(cuda_device), None, targets.to(cuda_device)), None)
    loss.backward()
    _, ref_loss, ref_grads, _ = mil_oracle.train_grads(sd, bags, coords, targets, None, n_heads=8)
    assert abs(float(loss.detach()) - float(ref_loss)) < LOGIT_TOL * max(1.0, float(ref_loss))
    _check_grads(model, ref_grads, "mha default N=400")


def test_small_model_long_bag_gradients_match_oracle(cuda_device):
    """dim_input 64 / dim_model 128 with enough tokens for the tcgen05 weight-gradient kernel: its 128 x 128
    tiles are only partly inside dW here (Kin = 64, Nout = 128 / 384), the rest comes from TMA zero fill."""
    from stamp_b200 import train as T

    sd = mil_oracle.init_state_dict(dim_input=64, dim_output=3, dim_model=128, n_heads=2, dim_feedforward=128, seed=9)
    bags, coords = mil_oracle.synthetic_bag(333, 64, seed=77, batch=3, signal=True)
    targets = torch.nn.functional.one_hot(torch.arange(3) % 3, 3).float()
    model = _model(sd, 2, cuda_device)
    loss = T.training_step(model, (bags.to(cuda_device), coords.to(cuda_device), None, targets.to(cuda_device)), None)
    loss.backward()
    _, ref_loss, ref_grads, _ = mil_oracle.train_grads(sd, bags, coords, targets, None)
    assert abs(float(loss.detach()) - float(ref_loss)) < LOGIT_TOL * max(1.0, float(ref_loss))
    _check_grads(model, ref_grads, "small model N=333")


def test_feature_gradient_and_gradcam_match_oracle(cuda_device):
    """d logits / d feats (heatmaps' jacrev, src/stamp/heatmaps/__init__.py:36-56) and the class-activation map."""
    from stamp_b200 import train as T

    sd = mil_oracle.init_state_dict(dim_input=64, dim_output=3, dim_model=128, n_heads=2, dim_feedforward=128,
                                    seed=5, running_mean=9000.0)
    feats, coords = mil_oracle.synthetic_bag(150, 64, seed=2)
    model = _model(sd, 2, cuda_device).eval()
    cam = T.gradcam_per_category(model, feats[0].to(cuda_device), coords[0].to(cuda_device)).cpu()
    x = feats[0].double().requires_grad_(True)
    sd64 = {k: v.double() for k, v in sd.items()}
    jac = torch.autograd.functional.jacobian(
        lambda b: mil_oracle.forward(sd64, b[None], coords.double(), None, exact_dist=True)[0], x)
    ref = torch.softmax((x.detach() * jac).mean(-1).abs(), dim=-1).T
    assert cam.shape == ref.shape == (150, 3)
    # the Jacobian rows themselves (the cam's softmax over tiles is forgiving)
    xg = feats.to(cuda_device).requires_grad_(True)
    out = model(xg, coords=coords.to(cuda_device), mask=None)
    for c in range(3):
        (j,) = torch.autograd.grad(out[0, c], xg, retain_graph=True)
        jrel = float((j[0].double().cpu() - jac[c]).norm() / jac[c].norm())
        print(f"d logit[{c}] / d feats rel err {jrel:.3e}")
        assert jrel < GRAD_TOL
    assert torch.allclose(cam.sum(0), torch.ones(3), atol=1e-4)
    rel = float((cam.double() - ref).norm() / ref.norm())
    print(f"gradcam rel err {rel:.3e}")
    assert rel < GRAD_TOL
    # the reference's own call (heatmaps/__init__.py:41-52): torch.func.jacrev straight through the module
    from torch.func import jacrev
    fd, cd = feats[0].to(cuda_device), coords[0].to(cuda_device)
    jf = jacrev(lambda b: model.forward(b.unsqueeze(0), coords=cd.unsqueeze(0), mask=None).squeeze(0))(fd)
    assert jf.shape == (3, 150, 64)
    jfrel = float((jf.double().cpu() - jac).norm() / jac.norm())
    print(f"jacrev Jacobian rel err {jfrel:.3e}")
    assert jfrel < GRAD_TOL
    cam_ref_style = torch.softmax((fd * jf).mean(-1).abs(), dim=-1).permute(-1, -2)
    assert torch.allclose(cam_ref_style.cpu(), cam, atol=1e-6)
    # eval mode: buffers untouched, no dropout
    assert float(model.transformer.layers[0][0].mhsa.attentions[0].scale_distance.items_so_far) == 1.0


def test_pairwise_dist_mean_matches_cdist(cuda_device):
    from stamp_b200 import train as T

    _, coords = mil_oracle.synthetic_bag(300, 8, seed=4, batch=3)
    c = torch.cat([torch.zeros(3, 1, 2), coords], dim=1).double()
    ref = torch.cdist(c, c).mean()
    out = T.pairwise_dist_mean(coords.to(cuda_device))
    assert abs(float(out) - float(ref)) < 1e-5 * float(ref)


def test_cross_entropy_kernel_matches_torch(cuda_device):
    from stamp_b200 import train as T

    g = torch.Generator().manual_seed(0)
    logits = (3 * torch.randn(37, 5, generator=g)).to(cuda_device).requires_grad_(True)
    targets = torch.softmax(torch.randn(37, 5, generator=g), dim=1).to(cuda_device)   # soft targets
    w = torch.rand(5, generator=g).to(cuda_device) + 0.5
    loss = T.cross_entropy(logits, targets, w)
    (3.0 * loss).backward()
    ref_l = logits.detach().clone().requires_grad_(True)
    ref = torch.nn.functional.cross_entropy(ref_l, targets, weight=w)
    (3.0 * ref).backward()
    assert torch.allclose(loss, ref, rtol=1e-5, atol=1e-6)
    assert torch.allclose(logits.grad, ref_l.grad, rtol=1e-4, atol=1e-6)


def test_fused_adamw_matches_torch_adamw(cuda_device):
    from stamp_b200 import train as T

    g = torch.Generator().manual_seed(1)
    shapes = [(128, 64), (128,), (3, 128), (1,)]
    ours = [torch.nn.Parameter(torch.randn(s, generator=g).to(cuda_device)) for s in shapes]
    theirs = [torch.nn.Parameter(p.detach().clone()) for p in ours]
    opt = T.FusedAdamW(ours, lr=1e-3)
    ref = torch.optim.AdamW(theirs, lr=1e-3)
    sched = torch.optim.lr_scheduler.OneCycleLR(opt, total_steps=6, max_lr=1e-2)
    rsched = torch.optim.lr_scheduler.OneCycleLR(ref, total_steps=6, max_lr=1e-2)
    for _ in range(5):
        for p, q in zip(ours, theirs):
            gr = torch.randn(p.shape, generator=g).to(cuda_device)
            p.grad.copy_(gr)       # grads are views of the flat buffer
            q.grad = gr.clone()
        opt.step(); ref.step(); sched.step(); rsched.step()
        opt.zero_grad()
    assert float(opt.flat_grad.abs().sum()) == 0.0
    for p, q in zip(ours, theirs):
        assert torch.allclose(p, q, rtol=1e-5, atol=1e-6)


def test_golden_adamw_step(cuda_device):
    """Parameters after the reference's first AdamW step.  The first Adam update is lr * g / (|g| + eps'):
    where |g| is far above round-off it is +-lr regardless of small gradient errors."""
    from stamp_b200 import train as T

    g = load_train_golden()
    model = _model(g["sd"], g["n_heads"], cuda_device)
    opt = T.FusedAdamW(model.parameters(), lr=1e-3)
    loss = T.training_step(model, (g["bags"].to(cuda_device), g["coords"].to(cuda_device), None,
                                   g["targets"].to(cuda_device)), g["class_weights"].to(cuda_device))
    loss.backward()
    opt.step()
    checked = total = 0
    for k, p in model.named_parameters():
        ref_after, ref_g = g["after"][k], g["grads"][k]
        sel = ref_g.abs() > 1e-6
        diff = (p.detach().cpu() - ref_after).abs()
        bad = (diff[sel] > 5e-5).float().mean() if sel.any() else torch.tensor(0.0)
        assert bad < 0.02, (k, float(bad))     # sign flips of near-zero gradients under bf16 only
        checked += int(sel.sum()); total += ref_g.numel()
    assert checked > 0.9 * total
    # inference after the step sees the new weights (pack cache invalidation through the flat buffer)
    with torch.no_grad():
        model.eval()
        a = model(g["bags"].to(cuda_device), coords=g["coords"].to(cuda_device), mask=None)
    sd_after = {k: v for k, v in g["after"].items()}
    ref = mil_oracle.forward(sd_after, g["bags"], g["coords"], None)
    assert ((a.cpu() - ref).norm(dim=1) / ref.norm(dim=1)).max() < 2e-2


@pytest.mark.parametrize("use_alibi", [False, True])
def test_eval_forward_after_optimizer_step_uses_new_weights(cuda_device, use_alibi):
    """validate -> train -> validate: the packed-weight cache of the inference path must be rebuilt after
    FusedAdamW.step (it writes the parameters through raw pointers, which no version counter sees).  With
    use_alibi=False -- the reference's default -- no running-mean buffer changes between the two evaluations, so
    nothing else invalidates the cache."""
    from stamp_b200 import train as T

    sd = mil_oracle.init_state_dict(dim_input=64, dim_output=2, dim_model=128, n_heads=2, dim_feedforward=128,
                                    seed=17, use_alibi=use_alibi)
    bags, coords = mil_oracle.synthetic_bag(90, 64, seed=3, batch=2, signal=True)
    targets = torch.nn.functional.one_hot(torch.arange(2) % 2, 2).float()
    model = _model(sd, 2, cuda_device)
    opt = T.FusedAdamW(model.parameters(), lr=5e-2)
    dev = lambda t: t.to(cuda_device)
    model.eval()
    with torch.no_grad():
        before = model(dev(bags), coords=dev(coords), mask=None).cpu()      # packs (and caches) the weights
    model.train()
    T.training_step(model, (dev(bags), dev(coords), None, dev(targets)), None).backward()
    opt.step()
    model.eval()
    with torch.no_grad():
        after = model(dev(bags), coords=dev(coords), mask=None).cpu()
    sd_now = {k: v.detach().cpu() for k, v in model.state_dict().items()}
    ref = mil_oracle.forward(sd_now, bags, coords, None, n_heads=2)
    assert (after - before).abs().max() > 1e-3, "the step at lr 5e-2 must move the logits"
    stale = ((before - ref).norm(dim=1) / ref.norm(dim=1)).max()
    rel = ((after - ref).norm(dim=1) / ref.norm(dim=1)).max()
    # a stale cache leaves the logits where they were (far from the new weights' oracle); a rebuilt one is at the
    # inference path's accuracy (the big step leaves the ALiBi variant slightly above its usual 1e-3)
    assert stale > 10 * rel and rel < 3e-3, (float(stale), float(rel))


@pytest.mark.parametrize("use_alibi", [False, True])
def test_packed_order_optimizer_receives_gradients_directly(cuda_device, use_alibi):
    """FusedAdamW(model=...) lays parameters and gradients out in the kernels' packed order: the backward adds into
    flat_grad itself.  Same seeds, same dropout masks: the gradients equal the ones autograd routes for a twin model
    without that optimiser, they accumulate over two backwards, and detaching one .grad falls back to autograd."""
    import copy

    from stamp_b200 import train as T

    sd = mil_oracle.init_state_dict(dim_input=64, dim_output=2, dim_model=128, n_heads=2, dim_feedforward=128,
                                    seed=23, use_alibi=use_alibi)
    bags, coords = mil_oracle.synthetic_bag(150, 64, seed=5, batch=3, signal=True)
    targets = torch.nn.functional.one_hot(torch.arange(3) % 2, 2).float()
    dev = lambda t: t.to(cuda_device)
    batch = (dev(bags), dev(coords), None, dev(targets))
    model = _model(sd, 2, cuda_device, dropout=0.0 if not use_alibi else 0.1, p_ff=0.1).train()
    twin = copy.deepcopy(model).train()
    opt = T.FusedAdamW(model.parameters(), lr=1e-3, model=model)
    assert opt.owns(model) and not opt.owns(twin)
    names = [n for n, _ in model.named_parameters()]

    def grads_of(m, n_backwards):
        torch.manual_seed(99)                     # the dropout seed is drawn from torch's generator
        for _ in range(n_backwards):
            T.training_step(m, batch, None).backward()
        return {n: p.grad.detach().clone() for n, p in m.named_parameters()}

    def compare(got, want, what):
        for n in names:
            scale = want[n].abs().max().clamp_min(1e-6)
            err = float((got[n] - want[n]).abs().max() / scale)
            # same kernels, same inputs: only the order of the fp32 atomics differs
            assert err < 2e-3, (what, n, err)

    want1 = grads_of(twin, 1)
    got1 = grads_of(model, 1)
    assert all(p.grad.data_ptr() >= opt.flat_grad.data_ptr() for p in model.parameters())
    compare(got1, want1, "one backward")
    twin.zero_grad()
    opt.zero_grad()
    assert float(opt.flat_grad.abs().max()) == 0.0
    compare(grads_of(model, 2), grads_of(twin, 2), "two accumulated backwards")
    # a detached gradient: autograd routes again, relink() folds the result back
    opt.zero_grad()
    twin.zero_grad()
    model.class_token.grad = None
    assert not opt.owns(model)
    got = grads_of(model, 1)
    opt.relink()
    assert opt.owns(model)
    compare(got, grads_of(twin, 1), "fallback through autograd")
    before = model.class_token.detach().clone()
    opt.step()
    assert float((model.class_token.detach() - before).abs().max()) > 0


def test_fused_adamw_state_dict_relink_and_groups(cuda_device):
    """Checkpoint / resume keeps the Adam moments and the bias-correction step; gradients detached from the flat
    buffer (model.zero_grad(set_to_none=True), stray .grad tensors) are folded back in; one group only."""
    from stamp_b200 import train as T

    g = torch.Generator().manual_seed(2)
    shapes = [(64, 32), (64,), (1,)]
    mk = lambda: [torch.nn.Parameter(torch.randn(s, generator=torch.Generator().manual_seed(7 + i)).to(cuda_device))
                  for i, s in enumerate(shapes)]
    a, b, r = mk(), mk(), [torch.nn.Parameter(p.detach().clone()) for p in mk()]
    oa, ob, ref = T.FusedAdamW(a, lr=1e-2), T.FusedAdamW(b, lr=1e-2), torch.optim.AdamW(r, lr=1e-2)
    grads = [[torch.randn(s, generator=g).to(cuda_device) for s in shapes] for _ in range(4)]

    def feed(params, gs, detach):
        for p, x in zip(params, gs):
            if detach:
                p.grad = None               # what model.zero_grad() does by default
                p.grad = x.clone()          # autograd then allocates a fresh tensor
            else:
                p.grad.copy_(x)

    for i in range(2):
        feed(a, grads[i], detach=False); oa.step(); oa.zero_grad()
        feed(r, grads[i], detach=True); ref.step()
    state = oa.state_dict()
    assert "fused" in state and state["fused"]["step"] == 2
    with torch.no_grad():
        for p, q in zip(b, a):
            p.copy_(q)
    ob.load_state_dict(state)               # resume in a fresh optimiser
    for i in range(2, 4):
        feed(b, grads[i], detach=True); ob.step()          # detached gradients: relink folds them in
        feed(r, grads[i], detach=True); ref.step()
    for p, q in zip(b, r):
        assert torch.allclose(p, q, rtol=1e-5, atol=1e-6)
        assert p.grad.data_ptr() >= ob.flat_grad.data_ptr()          # re-linked into the flat buffer
    with pytest.raises(ValueError):
        ob.add_param_group({"params": [torch.nn.Parameter(torch.zeros(4, device=cuda_device))]})
    with pytest.raises(KeyError):
        ob.load_state_dict(ref.state_dict())
    b[0].data = b[0].data.clone()           # parameter moved out of the flat buffer: refuse to step
    with pytest.raises(RuntimeError):
        ob.step()


def test_model_deepcopy_and_train_mode_forward_without_grad(cuda_device):
    import copy

    sd = mil_oracle.init_state_dict(dim_input=64, dim_output=2, dim_model=128, n_heads=2, dim_feedforward=128, seed=19)
    bags, coords = mil_oracle.synthetic_bag(40, 64, seed=4, batch=1)
    model = _model(sd, 2, cuda_device).eval()
    with torch.no_grad():
        a = model(bags.to(cuda_device), coords=coords.to(cuda_device), mask=None)
    twin = copy.deepcopy(model)             # the ctypes caches are dropped, not pickled
    assert twin._packed is None and twin._workspace is None
    with torch.no_grad():
        assert torch.equal(a, twin(bags.to(cuda_device), coords=coords.to(cuda_device), mask=None))
    # train mode under no_grad behaves like the reference module: the running mean moves
    twin.train()
    rm0 = twin.transformer.layers[0][0].mhsa.attentions[0].scale_distance.running_mean.clone()
    with torch.no_grad():
        twin(bags.to(cuda_device), coords=coords.to(cuda_device), mask=None)
    assert not torch.equal(rm0, twin.transformer.layers[0][0].mhsa.attentions[0].scale_distance.running_mean)
    with pytest.raises(NotImplementedError):
        twin(bags.to(cuda_device), coords=coords.to(cuda_device), mask=torch.zeros(1, 40, dtype=torch.bool, device=cuda_device))


def test_loss_decreases_on_planted_signal(cuda_device):
    """configs[3]-style loop at small scale: 5 % of the tiles of class-1 bags carry a mean shift."""
    from stamp_b200 import train as T
    from stamp_b200.mil import VisionTransformer

    torch.manual_seed(0)
    model = VisionTransformer(dim_output=2, dim_input=64, dim_model=128, n_layers=2, n_heads=2, dim_feedforward=128,
                              dropout=0.0, use_alibi=True).to(cuda_device).train()
    steps = 30
    opt, sched = T.configure_optimizers(model, total_steps=steps, max_lr=3e-3)
    B, n = 16, 200
    losses = []
    for it in range(steps):
        feats, coords = mil_oracle.synthetic_bag(n, 64, seed=100 + it, batch=B)
        y = torch.arange(B) % 2
        feats[y == 1, : n // 20] += 1.5
        batch = (feats.to(cuda_device), coords.to(cuda_device), None,
                 torch.nn.functional.one_hot(y, 2).float().to(cuda_device))
        opt.zero_grad()
        loss = T.training_step(model, batch, None)
        loss.backward()
        opt.step()
        sched.step()
        losses.append(float(loss))
    print("losses", [round(v, 3) for v in losses[::5]])
    assert np.isfinite(losses).all()
    assert np.mean(losses[-5:]) < 0.6 * np.mean(losses[:5])


def test_stale_checkpoint_is_refused(cuda_device):
    g = load_train_golden()
    model = _model(g["sd"], g["n_heads"], cuda_device)
    bags, coords = g["bags"].to(cuda_device), g["coords"].to(cuda_device)
    a = model(bags, coords=coords, mask=None).sum()
    model(bags, coords=coords, mask=None)
    with pytest.raises(RuntimeError):
        a.backward()


def test_gradcam_matches_reference_heatmap_code_golden(cuda_device):
    """mil_gradcam_trained_scale.npz was written by the reference's own _gradcam_per_category (jacrev over the
    reference module, oracle/make_golden_gradcam.py) for the state dict / bag of mil_alibi_trained_scale.npz."""
    import numpy as np

    from stamp_b200 import train as T
    from test_oracle_cpu import ROOT, load_golden

    sd, bags, coords, _mask, _logits, n_heads = load_golden(ROOT / "tests" / "golden" / "mil_alibi_trained_scale.npz")
    z = np.load(ROOT / "tests" / "golden" / "mil_gradcam_trained_scale.npz")
    cam_ref, scores_ref = torch.from_numpy(z["cam"]), torch.from_numpy(z["scores"])
    model = _model(sd, n_heads, cuda_device).eval()
    feats, c = bags[0].to(cuda_device), coords[0].to(cuda_device)
    cam = T.gradcam_per_category(model, feats, c).cpu()
    assert cam.shape == cam_ref.shape
    assert torch.allclose(cam, cam_ref, rtol=1e-3, atol=1e-7)
    # the pre-softmax scores are what carries the information (bf16 backward: 5e-2 of the row norm)
    x = feats.clone().requires_grad_(True)
    logits = model(x[None], coords=c[None], mask=None)[0]
    for k in range(3):
        (j,) = torch.autograd.grad(logits[k], x, retain_graph=True)
        s = (x.detach() * j).mean(-1).abs().cpu()
        rel = float((s - scores_ref[k]).norm() / scores_ref[k].norm())
        print(f"class {k}: heatmap score rel err {rel:.3e}")
        assert rel < GRAD_TOL
