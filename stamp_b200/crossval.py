"""k-fold cross-validation of the MIL aggregator on bags that stay resident in HBM.

Mirrors ``categorical_crossval_`` (src/stamp/modeling/crossval.py:48-370) for tile features; classification by default,
``task="regression"`` (LitTileRegressor: L1 loss, monitor ``validation_loss``) and ``task="survival"`` (LitTileSurvival: Cox
partial likelihood per batch, monitor ``val_cindex``, maximised; src/stamp/modeling/train.py:521-537) on a ``dim_output = 1``
backbone:

* folds from ``_get_splits`` (:373-423; :func:`stamp_b200.sharding.crossval_splits`);
* per fold, ``train_model_`` (src/stamp/modeling/train.py:504-564) on the training patients with the TEST fold as
  the validation set (the reference hands ``test_dl`` to ``valid_dl``, crossval.py:241-252,262-288):
  ``bag_size`` tiles per patient drawn with ``randperm`` (``_to_fixed_size_bag``, data.py:811-862), batches of
  ``batch_size`` shuffled every epoch, AdamW + OneCycleLR (models/__init__.py:133-141), class weights =
  normalised inverse frequencies of the training labels (train.py:567-621), validation = whole bags at batch 1,
  early stopping with ``patience`` and the best ``validation_loss`` checkpoint kept and restored (train.py:528-562);
* ``_predict`` on the test fold with the restored model (deploy.py:390-456): softmax probabilities per patient.

What differs is where the data lives and how the GPUs are used: a fold's features stay on the
device as the fp16 the feature files hold (320 x 4096 x 1024 x 2 B = 2.7 GB), there is no DataLoader, and the five
folds -- independent trainings -- either go one per GPU (``folds_for_rank``: no collective at all) or every fold is
trained data-parallel by all ranks (one all-reduce of the flat gradient buffer per step).
"""

from __future__ import annotations

import time
from collections.abc import Sequence
from dataclasses import dataclass, field

import torch
from torch import Tensor

from . import train as T
from .bags import to_fixed_size_bag
from .mil import VisionTransformer
from .sharding import all_reduce_flat_sum, crossval_splits, folds_for_rank, sync_alibi_running_mean, world


@dataclass
class Patient:
    pid: str
    feats: Tensor        # [N, F] fp16 (as stored) or fp32, on the training device
    coords: Tensor       # [N, 2] fp32
    label: int | float | tuple[float, float]     # class index | regression target | (time, event)


@dataclass
class FoldResult:
    fold: int
    test_patients: list[str]
    probs: Tensor                      # [n_test, C] on the host (regression: [n_test, 1] predictions; survival: [n_test] risk scores)
    history: list[dict] = field(default_factory=list)   # per epoch: training_loss, validation_loss
    best_epoch: int = -1
    epochs_run: int = 0
    seconds: float = 0.0
    train_steps: int = 0


def class_weights(labels: Sequence[int], n_classes: int, device) -> Tensor:
    """train.py:567-621: counts.sum() / counts, normalised to sum 1."""
    counts = torch.bincount(torch.as_tensor(list(labels)), minlength=n_classes).double()
    w = counts.sum() / counts
    return (w / w.sum()).float().to(device)


def _one_hot(labels: Sequence[int], n_classes: int, device) -> Tensor:
    return torch.nn.functional.one_hot(torch.as_tensor(list(labels)), n_classes).float().to(device)


@torch.no_grad()
def evaluate(model: VisionTransformer, patients: Sequence[Patient], n_classes: int, weights: Tensor | None,
             data_parallel: bool = False) -> tuple[float, Tensor]:
    """Whole bags at batch 1 in eval mode: (mean weighted cross-entropy, probabilities [n, C] on the device).
    With ``data_parallel`` the patients are split over the ranks and the results summed (one small all-reduce)."""
    import torch.distributed as dist

    rank, ws = world() if data_parallel else (0, 1)
    was_training = model.training
    model.eval()
    dev = next(model.parameters()).device
    probs = torch.zeros((len(patients), n_classes), device=dev)
    loss = torch.zeros((), device=dev)
    for i in range(rank, len(patients), ws):
        p = patients[i]
        logits = model(p.feats.unsqueeze(0), coords=p.coords.unsqueeze(0), mask=None).float()
        logp = torch.log_softmax(logits, dim=1)
        w = weights if weights is not None else torch.ones(n_classes, device=dev)
        loss += -(w[p.label] * logp[0, p.label])
        probs[i] = torch.softmax(logits, dim=1)[0]
    if ws > 1:
        dist.all_reduce(probs)
        dist.all_reduce(loss)
    model.train(was_training)
    return float(loss) / max(1, len(patients)), probs


def concordance_index(scores: Tensor, times: Tensor, events: Tensor) -> float:
    """``LitSurvivalBase.c_index`` (models/__init__.py:662-694): lifelines' ``concordance_index(times, -scores, events)``.
    A pair is comparable when the sample with the smaller time had an event (equal times: an event against a censored
    sample only); it is concordant when that sample has the higher risk score, tied scores count one half.  NaN when
    nothing is comparable.  (lifelines is not installable offline: this restates its documented pair rules and is
    checked against a brute-force pair loop, not against the library -- unpinned.)"""
    s, t, e = scores.flatten().double(), times.flatten().double(), events.flatten().bool()
    first = e[:, None] & ((t[:, None] < t[None, :]) | ((t[:, None] == t[None, :]) & ~e[None, :]))   # row died before column
    pairs = first.sum()
    if int(pairs) == 0:
        return float("nan")
    ds = s[:, None] - s[None, :]
    good = ((ds > 0) & first).sum().double() + 0.5 * ((ds == 0) & first).sum().double()
    return float(good / pairs.double())


def stratification_labels(labels: Sequence, task: str) -> list | None:
    """``_get_splits`` (crossval.py:373-423) with the splitter ``categorical_crossval_`` picks (:90-96): folds stratified by
    class, by event status for survival, plain ``KFold`` (None) for regression."""
    if task == "classification":
        return list(labels)
    if task == "survival":
        return [int(v[1]) for v in labels]
    if task == "regression":
        return None
    raise ValueError(f"unknown task {task!r}")


def _targets(labels: Sequence, task: str, n_classes: int, device) -> Tensor:
    if task == "classification":
        return _one_hot(labels, n_classes, device)
    return torch.as_tensor([list(v) if task == "survival" else [float(v)] for v in labels], dtype=torch.float32, device=device)


@torch.no_grad()
def evaluate_task(model: VisionTransformer, patients: Sequence[Patient], task: str) -> tuple[float, Tensor]:
    """Validation of the regression / survival tasks, whole bags at batch 1 in eval mode: (the monitored metric as a loss to
    MINIMISE -- mean absolute error, or minus the concordance index --, predictions [n, 1] on the device)."""
    was_training = model.training
    model.eval()
    preds = torch.cat([model(p.feats.unsqueeze(0), coords=p.coords.unsqueeze(0), mask=None).float() for p in patients])
    model.train(was_training)
    y = _targets([p.label for p in patients], task, 1, preds.device)
    if task == "regression":
        return float((preds - y).abs().mean()), preds
    ci = concordance_index(preds.squeeze(-1), y[:, 0], y[:, 1])
    return (-ci if ci == ci else float("inf")), preds


def train_fold(fold: int, train: Sequence[Patient], test: Sequence[Patient], *, n_classes: int, dim_input: int,
               model_params: dict | None = None, bag_size: int = 512, batch_size: int = 64, max_epochs: int = 32,
               patience: int = 16, max_lr: float = 1e-4, div_factor: float = 25.0, seed: int = 0,
               data_parallel: bool = False, task: str = "classification") -> FoldResult:
    """One split of the cross-validation (crossval.py:180-370) on the current CUDA device."""
    if task not in ("classification", "regression", "survival"):
        raise ValueError(f"unknown task {task!r}")
    if task != "classification":
        if data_parallel:
            raise ValueError("regression / survival folds train on one GPU each (mode='fold_per_gpu'): the risk sets of the "
                             "Cox loss span the whole batch")
        n_classes = 1
    rank, ws = world() if data_parallel else (0, 1)
    dev = train[0].feats.device
    torch.manual_seed(seed)                         # identical initial weights and dropout seeds on every rank
    params = dict(dim_model=512, n_layers=2, n_heads=8, dim_feedforward=512, dropout=0.25, use_alibi=False)
    params.update(model_params or {})
    model = VisionTransformer(dim_output=n_classes, dim_input=dim_input, **params).to(dev).train()
    w = class_weights([p.label for p in train], n_classes, dev) if task == "classification" else None
    step = {"classification": lambda b: T.training_step(model, b, w), "regression": lambda b: T.regression_step(model, b),
            "survival": lambda b: T.survival_step(model, b)}[task]
    steps_per_epoch = (len(train) + batch_size - 1) // batch_size
    opt, sched = T.configure_optimizers(model, total_steps=steps_per_epoch * max_epochs, max_lr=max_lr,
                                        div_factor=div_factor)
    gen = torch.Generator(device=dev).manual_seed(seed)          # tile sub-sampling
    order_gen = torch.Generator().manual_seed(seed)              # epoch shuffles (same on every rank)
    best = (float("inf"), -1, None)
    res = FoldResult(fold=fold, test_patients=[p.pid for p in test], probs=torch.empty(0))
    torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    for epoch in range(max_epochs):
        order = torch.randperm(len(train), generator=order_gen).tolist()
        run_loss, n_seen = torch.zeros((), device=dev), 0
        for s in range(0, len(order), batch_size):
            idx = order[s:s + batch_size]
            mine = idx[rank::ws]                      # data parallel: this rank's bags of the global batch
            opt.zero_grad()
            if mine:
                items = [to_fixed_size_bag(train[i].feats, train[i].coords, bag_size, generator=gen) for i in mine]
                bags = torch.stack([b for b, _, _ in items]).float()
                coords = torch.stack([c for _, c, _ in items])
                targets = _targets([train[i].label for i in mine], task, n_classes, dev)
                loss = step((bags, coords, None, targets))
                # the reference's loss is the mean over the GLOBAL batch: weight the local mean by its share
                (loss * (len(mine) / len(idx))).backward()
                run_loss += loss.detach() * len(mine)
                n_seen += len(mine)
            opt.relink()
            if ws > 1:
                all_reduce_flat_sum(opt.flat_grad)    # the one collective of the step; the shares already sum to 1
                sync_alibi_running_mean(model)
            opt.step()
            sched.step()
            res.train_steps += 1
        val_loss, _ = (evaluate(model, test, n_classes, w, data_parallel) if task == "classification"
                       else evaluate_task(model, test, task))
        res.history.append({"epoch": epoch, "training_loss": float(run_loss) / max(1, n_seen),
                            "validation_loss": val_loss})
        res.epochs_run = epoch + 1
        if val_loss < best[0]:                        # ModelCheckpoint(monitor="validation_loss", mode="min")
            best = (val_loss, epoch, {k: v.detach().clone() for k, v in model.state_dict().items()})
        elif epoch - best[1] >= patience:             # EarlyStopping(patience)
            break
    if best[2] is not None:
        model.load_state_dict(best[2])                # train_model_ reloads the best checkpoint (train.py:558-562)
        from .mil import bump_weights_epoch

        bump_weights_epoch()
    res.best_epoch = best[1]
    if task == "classification":
        _, probs = evaluate(model, test, n_classes, None, data_parallel)     # _predict on the held-out fold
    else:
        _, probs = evaluate_task(model, test, task)
        if task == "survival":
            probs = probs.squeeze(-1)                                        # deploy.py:447-448
    torch.cuda.synchronize(dev)
    res.seconds = time.perf_counter() - t0
    res.probs = probs.cpu()
    return res


def crossval(patients: Sequence[Patient], *, n_splits: int = 5, n_classes: int = 1, dim_input: int,
             mode: str = "fold_per_gpu", task: str = "classification", **fold_kw) -> list[FoldResult]:
    """``mode="fold_per_gpu"``: rank r trains folds r, r + world, ... alone (bit-for-bit the single-GPU training of
    that fold up to the order of fp32 atomic reductions; no collective); ``mode="dp_in_fold"``: every fold is trained
    by all ranks, data parallel.  Returns this rank's fold results (all folds in ``dp_in_fold`` mode)."""
    if mode not in ("fold_per_gpu", "dp_in_fold"):
        raise ValueError(mode)
    rank, ws = world()
    by_id = {p.pid: p for p in patients}
    splits = crossval_splits([p.pid for p in patients], stratification_labels([p.label for p in patients], task), n_splits)
    mine = folds_for_rank(n_splits, rank, ws) if mode == "fold_per_gpu" else list(range(n_splits))
    out = []
    for f in mine:
        tr, te = splits[f]
        out.append(train_fold(f, [by_id[i] for i in tr], [by_id[i] for i in te], n_classes=n_classes,
                              dim_input=dim_input, data_parallel=(mode == "dp_in_fold"), task=task, **fold_kw))
    return out
