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
