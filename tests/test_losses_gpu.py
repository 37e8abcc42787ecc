"""Losses of the regression and survival tasks (csrc/losses.cu behind stamp_b200.train) against the reference:
tests/golden/cox_loss.npz = ``neg_partial_log_likelihood`` of the reference's models/cox.py itself (loss and autograd
gradient, oracle/make_golden_cox.py); tests/golden/mil_train_step_{regression,survival}.npz = one training step of the
reference module with LitTileRegressor's / LitTileSurvival's loss (oracle/make_golden_train.py)."""

from pathlib import Path

import numpy as np
import pytest
import torch

GOLD = Path(__file__).parent / "golden"


def _cox_cases():
    z = np.load(GOLD / "cox_loss.npz")
    for name in sorted({k.split("/")[0] for k in z.files}):
        t = lambda k: torch.from_numpy(z[f"{name}/{k}"])      # noqa: E731
        yield name, t("log_hz"), t("time"), t("event"), str(z[f"{name}/ties"]), float(z[f"{name}/loss"]), t("grad")


def test_cox_oracle_matches_reference_golden():
    from oracle.cox_oracle import neg_partial_log_likelihood

    names = []
    for name, s, t, e, ties, loss, grad in _cox_cases():
        s64 = s.double().requires_grad_(True)
        got = neg_partial_log_likelihood(s64, t, e, ties)
        got.backward()
        assert abs(got.item() - loss) < 3e-6 * max(1.0, abs(loss)), name
        assert (s64.grad.float() - grad).abs().max().item() < 5e-6, name
        names.append(name)
    assert {"distinct_times", "ties_efron", "ties_breslow", "all_tied", "large_scores", "one_event"} <= set(names)
    s = torch.randn(5, dtype=torch.float64, requires_grad=True)
    assert neg_partial_log_likelihood(s, torch.rand(5), torch.zeros(5)).item() == 0.0     # no events (cox.py:209-214)


def test_losses_refuse_the_cpu():
    from stamp_b200 import train as T

    with pytest.raises(RuntimeError):
        T.neg_partial_log_likelihood(torch.randn(4), torch.rand(4), torch.ones(4))
    with pytest.raises(RuntimeError):
        T.l1_loss(torch.randn(4, 1), torch.randn(4, 1))
    with pytest.raises(ValueError):
        T.neg_partial_log_likelihood(torch.randn(4), torch.rand(4), torch.ones(4), ties_method="exact")


@pytest.mark.gpu
def test_cox_loss_kernel_matches_reference_golden(cuda_device):
    from stamp_b200 import train as T

    for name, s, t, e, ties, loss, grad in _cox_cases():
        sd = s.to(cuda_device).requires_grad_(True)
        got = T.neg_partial_log_likelihood(sd, t.to(cuda_device), e.to(cuda_device), ties_method=ties)
        (2.0 * got).backward()
        assert abs(got.item() - loss) < 3e-6 * max(1.0, abs(loss)), (name, got.item(), loss)
        assert (sd.grad.cpu() / 2.0 - grad).abs().max().item() < 5e-6, name
    sd = torch.randn(9, device=cuda_device, requires_grad=True)
    zero = T.neg_partial_log_likelihood(sd, torch.rand(9, device=cuda_device), torch.zeros(9, device=cuda_device))
    zero.backward()
    assert zero.item() == 0.0 and sd.grad.abs().max().item() == 0.0


@pytest.mark.gpu
@pytest.mark.parametrize("n,distinct,ties", [(4096, 40, "efron"), (8192, 8192, "efron"), (1500, 7, "breslow"), (1, 1, "efron")])
def test_cox_loss_kernel_matches_oracle_at_cohort_sizes(cuda_device, n, distinct, ties):
    from oracle.cox_oracle import neg_partial_log_likelihood
    from stamp_b200 import train as T

    g = torch.Generator().manual_seed(n)
    s = torch.randn(n, generator=g) * 1.5
    t = torch.randint(0, distinct, (n,), generator=g).float() if distinct < n else torch.randperm(n, generator=g).float()
    e = torch.rand(n, generator=g) < 0.6
    e[0] = True
    s64 = s.double().requires_grad_(True)
    want = neg_partial_log_likelihood(s64, t, e, ties)
    want.backward()
    sd = s.to(cuda_device).requires_grad_(True)
    got = T.neg_partial_log_likelihood(sd, t.to(cuda_device), e.to(cuda_device), ties_method=ties)
    got.backward()
    assert abs(got.item() - want.item()) < 2e-6 * max(1.0, abs(want.item())), (got.item(), want.item())
    assert ((sd.grad.cpu().double() - s64.grad).norm() / s64.grad.norm().clamp_min(1e-30)).item() < 1e-5 or n == 1
    with pytest.raises(RuntimeError):
        T.neg_partial_log_likelihood(torch.zeros(8193, device=cuda_device), torch.zeros(8193, device=cuda_device),
                                     torch.ones(8193, device=cuda_device))


@pytest.mark.gpu
def test_l1_loss_kernel_matches_torch(cuda_device):
    from stamp_b200 import train as T

    g = torch.Generator().manual_seed(3)
    for shape in ((64, 1), (5000, 1), (1, 1)):
        p, y = torch.randn(*shape, generator=g), torch.randn(*shape, generator=g)
        y[0] = p[0]                                                     # sign(0) = 0
        pr = p.clone().requires_grad_(True)
        want = torch.nn.functional.l1_loss(pr, y)
        want.backward()
        pd = p.to(cuda_device).requires_grad_(True)
        got = T.l1_loss(pd, y.to(cuda_device))
        got.backward()
        assert abs(got.item() - want.item()) < 1e-6
        assert torch.equal(pd.grad.cpu(), pr.grad)


@pytest.mark.gpu
@pytest.mark.parametrize("task", ["regression", "survival"])
def test_task_training_step_matches_reference_golden(cuda_device, task):
    from stamp_b200 import train as T
    from test_mil_train_cpu import load_train_golden
    from test_mil_train_gpu import LOGIT_TOL, _check_grads, _model

    g = load_train_golden(f"mil_train_step_{task}")
    model = _model(g["sd"], g["n_heads"], cuda_device)
    batch = (g["bags"].to(cuda_device), g["coords"].to(cuda_device), None, g["targets"].to(cuda_device))
    loss = (T.regression_step if task == "regression" else T.survival_step)(model, batch)
    loss.backward()
    print(f"{task}: loss {loss.item():.5f} vs reference {float(g['loss']):.5f}")
    assert abs(loss.item() - float(g["loss"])) < LOGIT_TOL * max(1.0, float(g["loss"]))
    ours = torch.cat([p.grad.double().cpu().flatten() for _, p in model.named_parameters()])
    ref = torch.cat([g["grads"][k].double().flatten() for k, _ in model.named_parameters()])
    cos = float(torch.dot(ours, ref) / (ours.norm() * ref.norm()))
    rel = float((ours - ref).norm() / ref.norm())
    print(f"{task}: whole gradient relative error {rel:.3e}, cos {cos:.5f}")
    assert cos > 0.998 and rel < 5e-2
    term_scale = None
    if task == "survival":
        # The per-bag cotangents of the Cox loss sum to zero, so gradients of parameters every bag sees alike (the biases of
        # the residual stream) are small differences of large per-bag terms: their error is measured against the size of
        # those terms, sum_b |dloss/dscore_b| * |dscore_b/dparam| (from per-bag backward passes of this module).
        saved = {k: p.grad.clone() for k, p in model.named_parameters()}
        scores = g["logits"].squeeze(-1).to(cuda_device).requires_grad_(True)
        y = g["targets"].to(cuda_device)
        T.neg_partial_log_likelihood(scores, y[:, 0], y[:, 1]).backward()
        term_scale = {k: 0.0 for k in saved}
        for b in range(scores.numel()):
            for p in model.parameters():
                p.grad = None
            model(batch[0], coords=batch[1], mask=None)[b, 0].backward()
            for k, p in model.named_parameters():
                term_scale[k] += abs(scores.grad[b].item()) * p.grad.norm().item()
        for k, p in model.named_parameters():
            p.grad = saved[k]
    _check_grads(model, g["grads"], task, term_scale=term_scale)


def test_concordance_index_pair_rules():
    """c-index of the survival validation (models/__init__.py:662-694, lifelines' pair rules): against a brute-force loop
    over the pairs, ties in time and score included."""
    import math

    from stamp_b200.crossval import concordance_index

    g = torch.Generator().manual_seed(0)
    for n in (1, 2, 7, 40, 200):
        s = torch.randint(0, 5, (n,), generator=g).float()
        t = torch.randint(0, 6, (n,), generator=g).float()
        e = torch.rand(n, generator=g) < 0.6
        pairs = good = 0.0
        for i in range(n):
            for j in range(n):
                if i != j and e[j] and (t[j] < t[i] or (t[j] == t[i] and not e[i])):     # j died before i left the study
                    pairs += 1
                    good += 1.0 if s[j] > s[i] else (0.5 if s[j] == s[i] else 0.0)
        got = concordance_index(s, t, e)
        assert (pairs == 0 and math.isnan(got)) or abs(got - good / pairs) < 1e-12
    assert concordance_index(torch.tensor([3.0, 2.0, 1.0]), torch.tensor([1.0, 2.0, 3.0]), torch.ones(3)) == 1.0
    assert concordance_index(torch.tensor([1.0, 2.0, 3.0]), torch.tensor([1.0, 2.0, 3.0]), torch.ones(3)) == 0.0


@pytest.mark.gpu
@pytest.mark.parametrize("task", ["regression", "survival"])
def test_crossval_and_deploy_of_the_other_tasks(cuda_device, task):
    """Cross-validation (crossval.py:48-370) and deployment (deploy.py:390-456) of the regression and survival tasks at
    small scale: a planted signal (the share of shifted tiles sets the target / the hazard) is learnt on held-out folds,
    the monitored metric picks the checkpoint, predictions come back as the reference's ``_predict`` shapes them."""
    from stamp_b200.crossval import Patient, concordance_index, crossval
    from stamp_b200.deploy import predict_patients
    from stamp_b200.mil import VisionTransformer

    g = torch.Generator().manual_seed(1)
    pats, truth = [], {}
    for i in range(36):
        n = 90 + int(torch.randint(0, 30, (1,), generator=g))
        level = (i % 6) / 5.0                                                 # 0 .. 1
        f = torch.randn(n, 64, generator=g)
        f[: int(n * 0.4 * level)] += 1.5
        cells = torch.randperm(400, generator=g)[:n]
        c = torch.stack([(cells % 20).float(), (cells // 20).float()], dim=-1) * 256.0
        label = 2.0 * level - 1.0 if task == "regression" else (10.0 - 8.0 * level + 0.01 * i, float(i % 4 != 0))
        truth[f"p{i:02d}"] = level
        pats.append(Patient(f"p{i:02d}", f.half().to(cuda_device), c.to(cuda_device), label))
    res = crossval(pats, n_splits=3, dim_input=64, mode="fold_per_gpu", task=task,
                   model_params=dict(dim_model=128, n_heads=2, dim_feedforward=128, dropout=0.0, use_alibi=False),
                   bag_size=64, batch_size=12, max_epochs=12, patience=4, max_lr=3e-3, seed=1)
    assert [r.fold for r in res] == [0, 1, 2]
    seen, preds, levels = [], [], []
    for r in res:
        assert r.probs.shape == ((len(r.test_patients), 1) if task == "regression" else (len(r.test_patients),))
        best = min(h["validation_loss"] for h in r.history)
        assert abs(r.history[r.best_epoch]["validation_loss"] - best) < 1e-12
        seen += r.test_patients
        preds += r.probs.flatten().tolist()
        levels += [truth[p] for p in r.test_patients]
    assert sorted(seen) == sorted(p.pid for p in pats)
    corr = torch.corrcoef(torch.tensor([preds, levels]))[0, 1].item()
    print(f"{task}: held-out correlation of the prediction with the planted level {corr:.3f}; "
          f"validation metric per fold {[round(min(h['validation_loss'] for h in r.history), 3) for r in res]}")
    # measured 0.92 (regression) / 0.70 (survival); the reference module trained the same way on the CPU: 0.96 / 0.79;
    # chance level for 36 patients is 0 +- 0.17
    assert corr > (0.6 if task == "regression" else 0.35), corr
    with pytest.raises(ValueError):
        crossval(pats, n_splits=3, dim_input=64, mode="dp_in_fold", task=task)

    # deployment: raw [1] predictions per patient for regression, scalar risk scores for survival
    model = VisionTransformer(dim_output=1, dim_input=64, dim_model=128, n_layers=2, n_heads=2, dim_feedforward=128,
                              dropout=0.0, use_alibi=False).to(cuda_device).eval()
    out = predict_patients(model, [p.pid for p in pats[:5]], [(p.feats.cpu(), p.coords.cpu()) for p in pats[:5]], cuda_device,
                           task=task)
    with torch.inference_mode():
        want = torch.cat([model(p.feats[None], coords=p.coords[None], mask=None).float() for p in pats[:5]]).cpu()
    for i, p in enumerate(pats[:5]):
        assert out[p.pid].shape == (() if task == "survival" else (1,))
        assert abs(out[p.pid].flatten()[0].item() - want[i, 0].item()) < 2e-3 * max(1.0, abs(want[i, 0].item()))
    if task == "survival":
        y = torch.tensor([p.label for p in pats])
        assert 0.0 <= concordance_index(torch.tensor(preds), y[:, 0][[int(s[1:]) for s in seen]], y[:, 1][[int(s[1:]) for s in seen]]) <= 1.0


def test_task_plumbing_on_the_host():
    """What the tasks change outside the kernels: target tensors of a batch (data.py's encoded targets: one-hot classes,
    [B, 1] regression values, [B, 2] = (time, event)), and ``_predict``'s post-processing (deploy.py:440-449)."""
    from stamp_b200.crossval import _targets
    from stamp_b200.deploy import _finish, predict_bags, predict_bags_ragged

    assert torch.equal(_targets([0, 2, 1], "classification", 3, "cpu"), torch.eye(3)[[0, 2, 1]])
    assert torch.equal(_targets([0.5, -1.0], "regression", 1, "cpu"), torch.tensor([[0.5], [-1.0]]))
    assert torch.equal(_targets([(5.0, 1), (3.0, 0)], "survival", 1, "cpu"), torch.tensor([[5.0, 1.0], [3.0, 0.0]]))
    logits = torch.tensor([[1.0, 3.0], [0.0, 0.0]])
    assert torch.allclose(_finish(logits, "classification"), torch.softmax(logits, dim=1))
    assert torch.equal(_finish(logits[:, :1].half(), "regression"), logits[:, :1]) and _finish(logits[:, :1], "survival").dtype == torch.float32
    for fn in (predict_bags, predict_bags_ragged):
        with pytest.raises(RuntimeError):
            fn(None, iter([]), "cpu")
