"""Barspoon encoder-decoder aggregator (stamp_b200/barspoon.py) against the reference's own EncDecTransformer: the golden
(tests/golden/barspoon.npz, oracle/make_golden_barspoon.py) holds the reference module's state dict, seeded inputs and
its logits per target, with and without the positional code."""

from pathlib import Path

import numpy as np
import pytest
import torch

GOLD = Path(__file__).parent / "golden" / "barspoon.npz"


def _load():
    from stamp_b200.barspoon import EncDecTransformer

    z = np.load(GOLD)
    targets = {str(k): int(n) for k, n in zip(z["labels"], z["n_outs"])}
    sd = {k[4:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("sd::")}
    make = lambda pe: EncDecTransformer(d_features=64, target_n_outs=targets, d_model=128, num_encoder_heads=2,  # noqa: E731
                                        num_decoder_heads=2, num_encoder_layers=2, num_decoder_layers=2,
                                        dim_feedforward=256, positional_encoding=pe)
    return z, targets, sd, make


def test_barspoon_state_dict_is_the_reference_one():
    """Same sub-modules, same keys and shapes: load_state_dict(strict=True) of the reference module's state dict."""
    z, targets, sd, make = _load()
    model = make(True)
    assert set(model.state_dict()) == set(sd)
    model.load_state_dict(sd, strict=True)
    assert list(model.target_labels) == list(targets) and "grade__WHO_" in model.heads
    with pytest.raises((RuntimeError, NotImplementedError)):
        model.eval()(torch.zeros(1, 4, 64), torch.zeros(1, 4, 2))          # CPU tensors / grad mode: refused


@pytest.mark.gpu
@pytest.mark.parametrize("pe", [True, False])
def test_barspoon_matches_reference_golden(cuda_device, pe):
    z, targets, sd, make = _load()
    model = make(pe)
    model.load_state_dict(sd, strict=True)
    model = model.to(cuda_device).eval()
    tokens, positions = torch.from_numpy(z["tokens"]).to(cuda_device), torch.from_numpy(z["positions"]).to(cuda_device)
    with torch.inference_mode():
        out = model(tokens, positions)
    assert list(out) == list(targets)
    got = torch.cat([out[t].float().cpu() for t in targets], dim=1)
    want = torch.cat([torch.from_numpy(z[("logits::" if pe else "logits_nope::") + t]) for t in targets], dim=1)
    assert got.shape == want.shape == (2, 9)
    # fp16 tensor-core operands against the reference's fp32: same bound as the MIL aggregator, per bag
    err = ((got - want).norm(dim=1) / want.norm(dim=1)).max().item()
    print(f"barspoon (positional_encoding={pe}): max per-bag relative error {err:.2e}")
    assert err < 1e-3, err
    with pytest.raises(NotImplementedError):
        model(tokens, positions)                                            # grad mode: inference only
