"""Multi-GPU orchestration: one process per GPU, slides / patients sharded across ranks.

The reference is single-device (``devices=1``: src/stamp/modeling/train.py:541-547,
src/stamp/modeling/deploy.py:402-406) and relies on "skip if the output exists" plus a random
shuffle for multi-machine runs (src/stamp/preprocessing/__init__.py:269-286).  Here:

* tile extraction, slide encoding, MIL deploy: independent units -> deterministic shards, NO
  data-path collective; rank 0 gathers small result tables on the host side;
* MIL training (data parallel over bags): exactly one exchange step per optimizer step -- an
  all-reduce of the flat gradient buffer -- plus the ALiBi ``running_mean`` buffers.

Works with any initialised ``torch.distributed`` backend (NCCL over NVLink on the B200 box, gloo in
the CPU tests).
"""

from __future__ import annotations

from collections.abc import Sequence
from typing import Any, TypeVar

import torch
import torch.distributed as dist
from torch import Tensor, nn

T = TypeVar("T")


def world() -> tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_round_robin(items: Sequence[T], rank: int, world_size: int) -> list[T]:
    """``items[rank::world]`` over a deterministic (sorted by the caller) list."""
    return list(items[rank::world_size])


def shard_lpt(items: Sequence[T], sizes: Sequence[float], rank: int, world_size: int) -> list[T]:
    """Longest-processing-time-first greedy assignment: slides have 2k..10k+ tiles, round-robin
    leaves ranks idle at the end of a cohort.  Deterministic on every rank (stable sort, ties by
    position), no communication needed."""
    if len(items) != len(sizes):
        raise ValueError("items and sizes must have the same length")
    order = sorted(range(len(items)), key=lambda i: (-float(sizes[i]), i))
    loads = [0.0] * world_size
    mine: list[int] = []
    for i in order:
        r = min(range(world_size), key=lambda j: (loads[j], j))
        loads[r] += float(sizes[i])
        if r == rank:
            mine.append(i)
    return [items[i] for i in sorted(mine)]


def gather_to_rank0(obj: Any) -> list[Any] | None:
    """Host-side gather of per-rank result tables (e.g. {patient: probabilities}); rank 0 gets
    the list, the others None.  Not on the data path."""
    rank, ws = world()
    if ws == 1:
        return [obj]
    out: list[Any] | None = [None] * ws if rank == 0 else None
    dist.gather_object(obj, out, dst=0)
    return out


class FlatGradAllReducer:
    """The one collective of data-parallel MIL training: average all gradients with a single
    all-reduce over one flat buffer (3.68 M parameters = 14.7 MB fp32 for the default model --
    latency-bound on NVLink 5, so one bucket, not many)."""

    def __init__(self, params: Sequence[nn.Parameter]) -> None:
        self.params = [p for p in params if p.requires_grad]
        n = sum(p.numel() for p in self.params)
        dev = self.params[0].device if self.params else torch.device("cpu")
        self.flat = torch.zeros(n, dtype=torch.float32, device=dev)

    @torch.no_grad()
    def all_reduce_mean(self) -> None:
        _, ws = world()
        off = 0
        for p in self.params:
            n = p.numel()
            if p.grad is None:
                self.flat[off:off + n].zero_()
            else:
                self.flat[off:off + n].copy_(p.grad.reshape(-1))
            off += n
        if ws > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM)
            self.flat.div_(ws)
        off = 0
        for p in self.params:
            n = p.numel()
            if p.grad is None:
                p.grad = torch.empty_like(p)
            p.grad.copy_(self.flat[off:off + n].view_as(p))
            off += n


@torch.no_grad()
def all_reduce_flat_sum(flat: Tensor) -> float:
    """Sum one flat gradient buffer over the ranks in place (the single collective of a data-parallel
    MIL step when the optimiser already keeps its gradients flat, ``train.FusedAdamW.flat_grad``).
    Returns the factor that turns the sum into the mean; the caller folds it into the optimiser step
    instead of spending another pass over the buffer."""
    _, ws = world()
    if ws > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    return 1.0 / ws


@torch.no_grad()
def sync_alibi_running_mean(model: nn.Module) -> None:
    """Ranks see different bags, so the ``scale_distance.running_mean`` buffers the reference
    updates inside forward (vision_tranformer.py:23-31) would diverge: average them."""
    _, ws = world()
    bufs = [b for n, b in model.named_buffers() if n.endswith("scale_distance.running_mean")]
    if ws == 1 or not bufs:
        return
    flat = torch.cat([b.reshape(-1).float() for b in bufs])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    flat.div_(ws)
    for i, b in enumerate(bufs):
        b.copy_(flat[i:i + 1].to(b.dtype))


def crossval_splits(patients: Sequence[str], labels: Sequence | None, n_splits: int) -> list[tuple[list[str], list[str]]]:
    """The reference's fold assignment (``_get_splits``, src/stamp/modeling/crossval.py:373-423):
    ``StratifiedKFold(n_splits, shuffle=True, random_state=0)`` over the patients with their class labels
    (plain ``KFold`` when ``labels`` is None: regression), so a fold trained here sees exactly the patients the
    reference's fold would.  Returns ``[(train_patients, test_patients)]`` in fold order."""
    import numpy as np
    from sklearn.model_selection import KFold, StratifiedKFold

    pts = np.array(list(patients))
    if labels is None:
        it = KFold(n_splits=n_splits, shuffle=True, random_state=0).split(pts)
    else:
        it = StratifiedKFold(n_splits=n_splits, shuffle=True, random_state=0).split(pts, np.array(list(labels)))
    return [(pts[tr].tolist(), pts[te].tolist()) for tr, te in it]


def folds_for_rank(n_folds: int, rank: int, world_size: int) -> list[int]:
    """Fold-per-GPU cross-validation: folds are independent trainings, so ``folds[rank::world]``
    needs no collective at all and reproduces the sequential reference fold by fold; with more GPUs than folds
    the surplus ranks get nothing (use data parallelism inside a fold instead)."""
    return list(range(n_folds))[rank::world_size]
