"""Heatmap export helpers around the MIL aggregator (``stamp heatmaps``).

Mirrors src/stamp/heatmaps/__init__.py: the class-activation map comes from ``train.gradcam_per_category`` (or the
reference's own ``jacrev`` call on the drop-in module, :36-56), per-tile scores from a batch-of-tiles forward
(:417-427), ``vals_to_im`` arranges per-tile values on the slide grid (:142-156) and ``ranked_tiles`` picks what
``_export_ranked_tiles`` saves (:190-238) with the exact top-k kernel.  Index / copy work stays in torch on the
tensors' device; nothing here falls back to the CPU for arithmetic.
"""

from __future__ import annotations

import torch
from torch import Tensor

from .encoder import topk as _topk


def vals_to_im(scores: Tensor, coords_norm: Tensor) -> Tensor:
    """Arranges per-tile ``scores [N, ...]`` in a 2-D grid according to the integer grid coordinates
    ``coords_norm [N, 2]`` (x, y): returns ``[max_y + 1, max_x + 1, ...]``, zeros where the slide has no tile."""
    size = coords_norm.max(0).values.flip(0) + 1
    im = torch.zeros((*size.tolist(), *scores.shape[1:]), dtype=scores.dtype, device=scores.device)
    flat = im.flatten(end_dim=1)
    flat[coords_norm[:, 1] * im.shape[1] + coords_norm[:, 0]] = scores
    return flat.reshape_as(im)


def ranked_tiles(tile_scores: Tensor, topk: int, bottomk: int) -> dict[str, tuple[Tensor, Tensor]]:
    """The tiles ``_export_ranked_tiles`` writes out: ``{"top": (scores, indices), "bottom": (scores, indices)}``,
    highest first / lowest first, at most ``topk`` / ``bottomk`` of them (ties resolve to the lower tile index)."""
    scores = tile_scores.detach().flatten().float().contiguous()
    if not scores.is_cuda:
        raise RuntimeError("ranked_tiles runs the exact top-k kernel: scores must be on a CUDA device")
    out: dict[str, tuple[Tensor, Tensor]] = {}
    n = scores.numel()
    if n == 0:
        return out
    if min(topk, n) > 0:
        v, i = _topk(scores, min(topk, n), largest=True)
        out["top"] = (v, i)
    if min(bottomk, n) > 0:
        v, i = _topk(scores, min(bottomk, n), largest=False)
        out["bottom"] = (v, i)
    return out
