"""Host-side checks of the training / tiling C-ABI entry points that need no GPU: size queries and argument
validation (every entry point rejects bad arguments before it touches the device)."""

import ctypes as C

import pytest

from stamp_b200 import _lib
from stamp_b200.mil import StampMilConfig


def _cfg(**kw):
    base = dict(dim_input=1024, dim_model=512, n_layers=2, n_heads=8, dim_ff=512, dim_output=2, use_alibi=1)
    base.update(kw)
    return StampMilConfig(**base)


def test_train_ctx_bytes_and_envelope():
    from stamp_b200.train import _bind

    lib = _bind()
    small = lib.stamp_mil_train_ctx_bytes(C.byref(_cfg()), 1, 64)
    big = lib.stamp_mil_train_ctx_bytes(C.byref(_cfg()), 8, 4096)
    assert 0 < small < big < 4 * 2 ** 30                      # 8 bags of 4096 x 1024: ~1.6 GB of checkpoints
    assert lib.stamp_mil_train_ctx_bytes(C.byref(_cfg(use_alibi=0)), 2, 100) > 0      # nn.MultiheadAttention variant
    for bad in (dict(dim_model=520), dict(n_heads=3), dict(dim_input=1001), dict(dim_model=2048, n_heads=32),
                dict(n_layers=0), dict(dim_model=768, n_heads=8)):                   # head dim 96: unsupported
        assert lib.stamp_mil_train_ctx_bytes(C.byref(_cfg(**bad)), 2, 100) == 0, bad
    assert lib.stamp_mil_train_ctx_bytes(C.byref(_cfg()), 2, 0) == 0                  # empty bags cannot be trained on
    assert lib.stamp_pairwise_dist_mean_workspace_bytes(4, 1000) >= 4 * 1001 * 8
    assert lib.stamp_pairwise_dist_mean_workspace_bytes(0, 10) == 0


def test_entry_points_reject_null_arguments_without_a_device():
    from stamp_b200.tiling import _bind as bind_tiling
    from stamp_b200.train import _bind

    lib = _bind()
    bind_tiling()
    bad = -1
    assert lib.stamp_adamw_step(None, None, None, None, 10, 1e-3, 0.9, 0.999, 1e-8, 0.01, 1, 1.0, None) == bad
    assert lib.stamp_cross_entropy(None, None, None, 4, 3, 1.0, None, None, None) == bad
    assert lib.stamp_mil_train_dropout_mask(1, 0, 16, 0.5, None, None) == bad
    assert lib.stamp_mil_train_dropout_mask(1, 0, 16, 1.5, 8, None) == bad            # p outside [0, 1)
    assert lib.stamp_pairwise_dist_mean(None, 1, 10, None, None, 0, None) == bad
    assert lib.stamp_tile_texture_u8(None, 1, 224, 224, 40, 100, None, None, None) == bad
    assert lib.stamp_mil_train_forward(None, None, None, None, None, None, None, 1, 10, None, 0, None) == bad
    assert lib.stamp_mil_train_backward(None, None, None, None, None, None, None, None, 1, 10, None, 0, None) == bad
    assert b"argument" in _lib.load().stamp_b200_strerror(bad)


def test_training_api_refuses_cpu_tensors():
    import torch

    from stamp_b200 import train as T
    from stamp_b200.mil import VisionTransformer

    m = VisionTransformer(dim_output=2, dim_input=64, dim_model=128, n_layers=1, n_heads=2, dim_feedforward=128,
                          dropout=0.0, use_alibi=True)
    with pytest.raises(RuntimeError):
        m(torch.randn(1, 5, 64), coords=torch.rand(1, 5, 2), mask=None)       # autograd on, CPU: no fallback
    with pytest.raises(RuntimeError):
        T.FusedAdamW(m.parameters())
    with pytest.raises(RuntimeError):
        T.cross_entropy(torch.randn(2, 3, requires_grad=True), torch.rand(2, 3))
    with pytest.raises(RuntimeError):
        T.pairwise_dist_mean(torch.rand(1, 5, 2))
