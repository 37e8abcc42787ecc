"""Generates tests/golden/mil_train_step*.npz by running the REFERENCE module itself in training mode.

Run in the build container only (``python oracle/make_golden_train.py``).  Imports
``/root/reference/src/stamp/modeling/models/vision_tranformer.py`` by file path, puts the model in
``train()`` mode with every nn.Dropout set to p = 0 (random masks cannot be pinned; the dropout sites
are covered separately against the oracle with explicit masks), and records one optimisation step as
``LitTileClassifier`` runs it (src/stamp/modeling/models/__init__.py:239-286, :133-141):

    logits = model(bags, coords=coords, mask=None)
    loss = F.cross_entropy(logits, soft_targets, weight=class_weights); loss.backward()
    torch.optim.AdamW(model.parameters(), lr=1e-3).step()

Stored: inputs, state dict before, logits, loss, every parameter gradient, the running-mean buffers
after the forward, and the parameters after the AdamW step.
"""

from __future__ import annotations

import sys
from pathlib import Path

import numpy as np
import torch
import torch.nn.functional as F

sys.path.insert(0, str(Path(__file__).resolve().parent))
from make_golden import OUT, load_reference  # noqa: E402

DIMS = dict(dim_input=64, dim_model=128, n_layers=2, n_heads=2, dim_feedforward=128, dim_output=3)


def main() -> None:
    ref = load_reference()
    one_case(ref, use_alibi=True, name="mil_train_step", seed=4242)
    # the reference's default backbone (VitModelParams.use_alibi = False, dropout = 0.0): nn.MultiheadAttention
    one_case(ref, use_alibi=False, name="mil_train_step_mha", seed=4343)
    # the other two tasks of the reference on the same backbone with dim_output = 1:
    # LitTileRegressor (models/__init__.py:420-462, l1_loss) and LitTileSurvival (:751-776, cox.py's loss, Efron ties)
    one_case(ref, use_alibi=True, name="mil_train_step_regression", seed=4444, task="regression")
    one_case(ref, use_alibi=False, name="mil_train_step_survival", seed=4545, task="survival")


def load_cox():
    import importlib.util

    spec = importlib.util.spec_from_file_location("ref_cox", "/root/reference/src/stamp/modeling/models/cox.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def one_case(ref, *, use_alibi: bool, name: str, seed: int, task: str = "classification") -> None:
    torch.manual_seed(seed)
    dims = DIMS if task == "classification" else {**DIMS, "dim_output": 1}
    model = ref.VisionTransformer(dropout=0.25 if use_alibi else 0.0, use_alibi=use_alibi, **dims).train()
    for m in model.modules():
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0
    g = torch.Generator().manual_seed(77)
    B, n = (3, 77) if task == "classification" else (6, 45)
    bags = torch.randn(B, n, DIMS["dim_input"], generator=g).half().float()
    cells = torch.stack([torch.randperm(100 * 100, generator=g)[:n] for _ in range(B)])
    coords = torch.stack([(cells % 100).float(), (cells // 100).float()], dim=-1) * 256.0
    if task == "classification":
        targets = F.one_hot(torch.tensor([0, 2, 1]), 3).float()
        class_weights = torch.tensor([0.7, 1.1, 1.6])
    elif task == "regression":
        targets, class_weights = torch.randn(B, 1, generator=g), torch.zeros(1)
    else:                                      # (time, event); two patients share a time: Efron's correction is exercised
        targets = torch.tensor([[5.0, 1.0], [3.0, 1.0], [5.0, 1.0], [8.0, 0.0], [1.0, 1.0], [4.0, 0.0]])
        class_weights = torch.zeros(1)
    before = {k: v.clone() for k, v in model.state_dict().items()}

    opt = torch.optim.AdamW(model.parameters(), lr=1e-3)
    logits = model(bags, coords=coords, mask=None)
    if task == "classification":
        loss = F.cross_entropy(logits, targets, weight=class_weights)
    elif task == "regression":
        loss = F.l1_loss(logits, targets.to(logits).float())
    else:
        import warnings

        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            loss = load_cox().neg_partial_log_likelihood(logits.squeeze(-1), targets[:, 0], targets[:, 1])
    loss.backward()
    grads = {k: p.grad.clone() for k, p in model.named_parameters()}
    opt.step()
    after = model.state_dict()

    arrays = {f"sd/{k}": v.numpy() for k, v in before.items()}
    arrays.update({f"grad/{k}": v.numpy() for k, v in grads.items()})
    arrays.update({f"after/{k}": v.detach().numpy() for k, v in after.items()})
    arrays.update(bags=bags.numpy(), coords=coords.numpy(), targets=targets.numpy(),
                  class_weights=class_weights.numpy(), logits=logits.detach().numpy(),
                  loss=loss.detach().numpy(), n_heads=np.int64(DIMS["n_heads"]))
    arrays["use_alibi"] = np.bool_(use_alibi)
    arrays["task"] = np.array(task)
    np.savez_compressed(OUT / f"{name}.npz", **arrays)
    print(name, "loss", float(loss), "logits", logits[0].tolist())


if __name__ == "__main__":
    main()
