"""Host-side bag plumbing that feeds the MIL aggregator.

Mirrors the reference's behaviour, not its code: ``to_fixed_size_bag`` follows
``_to_fixed_size_bag`` (src/stamp/modeling/data.py:811-862: ``randperm(n)[:bag_size]`` sub-sampling,
or equidistant ``linspace().round()`` when deterministic, zero padding to the right, returned size
``min(bag_size, n)``); ``collate_bags`` follows ``_collate_to_tuple`` (:255-277).  Pure index /
copy work on whatever device the features live on -- keeping a fold's bags resident in HBM
removes the per-step DataLoader up-cast.
"""

from __future__ import annotations

import torch
from torch import Tensor


def to_fixed_size_bag(bag: Tensor, coords: Tensor, bag_size: int, deterministic: bool = False,
                      generator: torch.Generator | None = None) -> tuple[Tensor, Tensor, int]:
    n_tiles = bag.shape[0]
    if n_tiles <= bag_size:
        idx = torch.arange(n_tiles, device=bag.device)
    elif deterministic:
        idx = torch.linspace(0, n_tiles - 1, steps=bag_size, device=bag.device).round().long()
    else:
        idx = torch.randperm(n_tiles, device=bag.device, generator=generator)[:bag_size]
    out_bag = bag.new_zeros((bag_size, bag.shape[1]))
    out_coords = coords.new_zeros((bag_size, coords.shape[1]))
    out_bag[: idx.numel()] = bag[idx]
    out_coords[: idx.numel()] = coords[idx]
    return out_bag, out_coords, min(bag_size, n_tiles)


def collate_bags(items: list[tuple[Tensor, Tensor, int, Tensor]]) -> tuple[Tensor, Tensor, Tensor, Tensor]:
    """[(bag [n,F], coords [n,2], bag_size, target)] -> (bags [B,n,F], coords [B,n,2], sizes [B], targets [B,C])."""
    bags = torch.stack([b for b, _, _, _ in items])
    coords = torch.stack([c for _, c, _, _ in items])
    sizes = torch.tensor([s for _, _, s, _ in items])
    targets = []
    for _, _, _, t in items:
        t = torch.as_tensor(t)
        targets.append(t.unsqueeze(0) if t.ndim == 0 else t.reshape(-1))
    return bags, coords, sizes, torch.stack(targets)


def padding_mask(bag_sizes: Tensor, bag_len: int) -> Tensor:
    """mask[b, i] = True for zero-padded tiles (i >= bag_size[b]); what ``_step`` would pass if
    ``use_mask`` were on (src/stamp/modeling/models/__init__.py:244-250)."""
    return torch.arange(bag_len, device=bag_sizes.device)[None, :] >= bag_sizes[:, None]


def prefetch_to_device(batches, device, depth: int = 2):
    """Yield the batches of ``batches`` (tuples of tensors; ``None`` entries pass through) on ``device``, with the
    host->device copies of the next ``depth - 1`` batches running on a side stream while the caller works on the
    current one.  What the reference gets from ``DataLoader(pin_memory=True)`` plus Lightning's synchronous
    ``batch.to(device)`` (src/stamp/modeling/data.py:255-277, train.py:541-547), minus the stall: a 64 MB batch of
    fp16 bags takes ~2.5 ms over PCIe, a training step ~7 ms.  Host tensors are pinned on first use."""
    from collections import deque

    device = torch.device(device)
    if device.type != "cuda":
        raise RuntimeError("prefetch_to_device targets a CUDA device")
    copy_stream = torch.cuda.Stream(device=device)
    queue: deque = deque()

    def enqueue(batch) -> None:
        with torch.cuda.stream(copy_stream):
            moved, keep = [], []
            for t in batch:
                if isinstance(t, Tensor) and not t.is_cuda:
                    src = t if t.is_pinned() else t.pin_memory()
                    keep.append(src)
                    moved.append(src.to(device, non_blocking=True))
                else:
                    moved.append(t)
            ev = torch.cuda.Event()
            ev.record(copy_stream)
        queue.append((tuple(moved), ev, keep))

    it = iter(batches)
    for batch in it:
        enqueue(batch)
        if len(queue) < depth:
            continue
        moved, ev, _keep = queue.popleft()
        cur = torch.cuda.current_stream(device)
        cur.wait_event(ev)
        for t in moved:
            if isinstance(t, Tensor) and t.is_cuda:
                t.record_stream(cur)
        yield moved
    while queue:
        moved, ev, _keep = queue.popleft()
        cur = torch.cuda.current_stream(device)
        cur.wait_event(ev)
        for t in moved:
            if isinstance(t, Tensor) and t.is_cuda:
                t.record_stream(cur)
        yield moved
