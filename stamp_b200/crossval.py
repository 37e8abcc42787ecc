"""k-fold cross-validation of the MIL aggregator on bags that stay resident in HBM.

Mirrors ``categorical_crossval_`` (src/stamp/modeling/crossval.py:48-370) for tile features / classification:

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
    label: int


@dataclass
class FoldResult:
    fold: int
    test_patients: list[str]
    probs: Tensor                      # [n_test, C] on the host
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


def train_fold(fold: int, train: Sequence[Patient], test: Sequence[Patient], *, n_classes: int, dim_input: int,
               model_params: dict | None = None, bag_size: int = 512, batch_size: int = 64, max_epochs: int = 32,
               patience: int = 16, max_lr: float = 1e-4, div_factor: float = 25.0, seed: int = 0,
               data_parallel: bool = False) -> FoldResult:
    """One split of the cross-validation (crossval.py:180-370) on the current CUDA device."""
    rank, ws = world() if data_parallel else (0, 1)
    dev = train[0].feats.device
    torch.manual_seed(seed)                         # identical initial weights and dropout seeds on every rank
    params = dict(dim_model=512, n_layers=2, n_heads=8, dim_feedforward=512, dropout=0.25, use_alibi=False)
    params.update(model_params or {})
    model = VisionTransformer(dim_output=n_classes, dim_input=dim_input, **params).to(dev).train()
    w = class_weights([p.label for p in train], n_classes, dev)
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
                targets = _one_hot([train[i].label for i in mine], n_classes, dev)
                loss = T.training_step(model, (bags, coords, None, targets), w)
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
        val_loss, _ = evaluate(model, test, n_classes, w, data_parallel)
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
    _, probs = evaluate(model, test, n_classes, None, data_parallel)     # _predict on the held-out fold
    torch.cuda.synchronize(dev)
    res.seconds = time.perf_counter() - t0
    res.probs = probs.cpu()
    return res


def crossval(patients: Sequence[Patient], *, n_splits: int = 5, n_classes: int, dim_input: int,
             mode: str = "fold_per_gpu", **fold_kw) -> list[FoldResult]:
    """``mode="fold_per_gpu"``: rank r trains folds r, r + world, ... alone (bit-for-bit the single-GPU training of
    that fold up to the order of fp32 atomic reductions; no collective); ``mode="dp_in_fold"``: every fold is trained
    by all ranks, data parallel.  Returns this rank's fold results (all folds in ``dp_in_fold`` mode)."""
    if mode not in ("fold_per_gpu", "dp_in_fold"):
        raise ValueError(mode)
    rank, ws = world()
    by_id = {p.pid: p for p in patients}
    splits = crossval_splits([p.pid for p in patients], [p.label for p in patients], n_splits)
    mine = folds_for_rank(n_splits, rank, ws) if mode == "fold_per_gpu" else list(range(n_splits))
    out = []
    for f in mine:
        tr, te = splits[f]
        out.append(train_fold(f, [by_id[i] for i in tr], [by_id[i] for i in te], n_classes=n_classes,
                              dim_input=dim_input, data_parallel=(mode == "dp_in_fold"), **fold_kw))
    return out
