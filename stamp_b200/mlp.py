"""MLP and Linear aggregators on the B200 primitives -- inference.

Mirror ``MLP`` / ``Linear`` (src/stamp/modeling/models/mlp.py:6-62; ``ModelName.MLP`` / ``ModelName.LINEAR``,
src/stamp/modeling/registry.py): same constructors, same sub-modules and state-dict keys (``mlp.<i>`` / ``fc``), same
``forward(x)`` for ``[B, F]`` feature vectors and ``[B, T, F]`` bags (mean over the tiles first).

  mean over the tiles (:40-41, :57-58)     stamp_bag_mean: HBM-bound, every feature read once (fp32 or fp16 bags, the
                                           fp16 bags of the feature files go in without widening), deterministic
  Linear (+ ReLU) layers                   stamp_sgemm_batched_f32 with the bias / ReLU epilogue: B rows, fp32 -- the
                                           products are a few MFLOP, the precision is the reference's

Any other rank raises ``ValueError`` like the reference.  Dropout is the identity in eval mode; the training step stays
with the reference module (these heads train in seconds on pooled features).
"""

from __future__ import annotations

import ctypes as C

import torch
from torch import Tensor, nn

from . import _lib
from .transmil import _bind as _bind_f32


def _bind() -> C.CDLL:
    lib = _bind_f32()
    if not getattr(lib, "_bagmean_bound", False):
        ll, vp, i = C.c_longlong, C.c_void_p, C.c_int
        lib.stamp_bag_mean_splits.argtypes = [i, i, i, i]
        lib.stamp_bag_mean_splits.restype = i
        lib.stamp_bag_mean.argtypes = [vp, i, ll, ll, i, i, i, vp, ll, vp, i, vp]
        lib.stamp_bag_mean.restype = i
        lib._bagmean_bound = True
    return lib


def bag_mean(x: Tensor) -> Tensor:
    """``x.mean(dim=1)`` of a CUDA ``[B, T, F]`` fp32 / fp16 tensor -> fp32 ``[B, F]`` (stamp_bag_mean)."""
    if not x.is_cuda:
        raise RuntimeError("stamp_b200 bag_mean runs on a CUDA device only (no CPU fallback)")
    if x.ndim != 3 or x.shape[1] == 0:
        raise ValueError(f"expected a non-empty [B, T, F] bag tensor, got {tuple(x.shape)}")
    if x.dtype not in (torch.float32, torch.float16):
        x = x.float()
    if x.stride(2) != 1 or x.stride(1) < x.shape[2]:
        x = x.contiguous()
    lib = _bind()
    B, T, F = x.shape
    half = int(x.dtype == torch.float16)
    splits = lib.stamp_bag_mean_splits(B, T, F, half)
    out = torch.empty((B, F), dtype=torch.float32, device=x.device)
    scratch = torch.empty((splits, B, F), dtype=torch.float32, device=x.device) if splits > 1 else None
    _lib.check(lib.stamp_bag_mean(x.data_ptr(), half, x.stride(1), x.stride(0) if B > 1 else T * x.stride(1), B, T, F, out.data_ptr(), F,
                                  scratch.data_ptr() if scratch is not None else None, splits,
                                  torch.cuda.current_stream().cuda_stream), "stamp_bag_mean")
    return out


def _linear(x: Tensor, lin: nn.Linear, relu: bool) -> Tensor:
    """fp32 ``x @ W^T + b`` (optionally ReLU) for a contiguous CUDA ``[B, F]`` matrix."""
    lib = _bind()
    w = lin.weight.detach().float().contiguous()
    b = lin.bias.detach().float().contiguous() if lin.bias is not None else None
    M, K = x.shape
    N = w.shape[0]
    out = torch.empty((M, N), dtype=torch.float32, device=x.device)
    _lib.check(lib.stamp_sgemm_batched_f32(x.data_ptr(), K, 0, w.data_ptr(), K, 0, out.data_ptr(), N, 0, M, N, K, 1, 1, 1.0, 0.0,
                                           b.data_ptr() if b is not None else None, 2 if relu else 0,
                                           torch.cuda.current_stream().cuda_stream), "stamp_sgemm_batched_f32")
    return out


def _pooled(x: Tensor, module: nn.Module) -> Tensor:
    if torch.is_grad_enabled() and any(p.requires_grad for p in module.parameters()):
        raise NotImplementedError("stamp_b200 MLP / Linear are inference-only: call them under torch.no_grad() / "
                                  "inference_mode() (training runs through the reference module)")
    if module.training:
        raise RuntimeError("call .eval() first: dropout of the training mode is not implemented")
    if x.ndim not in (2, 3):
        raise ValueError(f"Expected 2D or 3D input, got {x.shape}")
    if not x.is_cuda or not next(module.parameters()).is_cuda:
        raise RuntimeError("stamp_b200 MLP / Linear run on a CUDA device only (no CPU fallback)")
    if x.ndim == 3:
        return bag_mean(x)
    return x.detach().float().contiguous()


class MLP(nn.Module):
    def __init__(self, dim_input: int, dim_hidden: int, dim_output: int, num_layers: int, dropout: float) -> None:
        super().__init__()
        layers: list[nn.Module] = []
        in_dim = dim_input
        for _ in range(num_layers - 1):
            layers += [nn.Linear(in_dim, dim_hidden), nn.ReLU(), nn.Dropout(dropout)]
            in_dim = dim_hidden
        layers.append(nn.Linear(in_dim, dim_output))
        self.mlp = nn.Sequential(*layers)

    def forward(self, x: Tensor, **kwargs) -> Tensor:
        x = _pooled(x, self)
        linears = [m for m in self.mlp if isinstance(m, nn.Linear)]
        for i, lin in enumerate(linears):
            x = _linear(x, lin, relu=i + 1 < len(linears))
        return x


class Linear(nn.Module):
    def __init__(self, dim_input: int, dim_output: int) -> None:
        super().__init__()
        self.fc = nn.Linear(dim_input, dim_output)

    def forward(self, x: Tensor, **kwargs) -> Tensor:
        return _linear(_pooled(x, self), self.fc, relu=False)
