"""TransMIL aggregator (stamp_b200/transmil.py) against the reference's own module: tests/golden/transmil.npz holds the
logits of the reference TransMIL (imported by path in oracle/make_golden_transmil.py) for seeded weights and bags of 300 and
1100 tiles; weights and inputs are regenerated here from the same seeds (inputs verified by checksum)."""

from pathlib import Path

import numpy as np
import pytest
import torch

GOLD = Path(__file__).parent / "golden" / "transmil.npz"


def _bags(n: int, g: torch.Generator) -> torch.Tensor:
    return torch.randn(2, n, 64, generator=g).half().float()


def test_transmil_state_dict_is_the_reference_one():
    from oracle.transmil_weights import transmil_state_dict
    from stamp_b200.transmil import TransMIL

    model = TransMIL(dim_output=3, dim_input=64, dim_hidden=512)
    sd = transmil_state_dict(3, 64, 512)
    assert set(model.state_dict()) == set(sd)
    model.load_state_dict(sd, strict=True)
    with pytest.raises((RuntimeError, NotImplementedError)):
        model.eval()(torch.zeros(1, 4, 64))


@pytest.mark.gpu
def test_transmil_matches_reference_golden(cuda_device):
    from oracle.transmil_weights import transmil_state_dict
    from stamp_b200.transmil import TransMIL

    z = np.load(GOLD)
    model = TransMIL(dim_output=3, dim_input=64, dim_hidden=512)
    model.load_state_dict(transmil_state_dict(3, 64, 512), strict=True)
    model = model.to(cuda_device).eval()
    g = torch.Generator().manual_seed(9)
    for n in (300, 1100):
        bags = _bags(n, g)
        assert abs(bags.double().sum().item() - float(z[f"bags_checksum_{n}"])) < 1e-9      # the golden's inputs
        with torch.inference_mode():
            got = model(bags.to(cuda_device)).float().cpu()
        want = torch.from_numpy(z[f"logits_{n}"])
        err = ((got - want).norm(dim=1) / want.norm(dim=1)).max().item()
        print(f"TransMIL, {n} tiles (_fc1 on the tcgen05 GEMM, fp16 operands): max per-bag relative error {err:.2e}")
        assert got.shape == (2, 3) and err < 1e-3, (n, err, got, want)
        model.fc1_fp32 = True                        # everything in fp32: the restatement itself is exact
        with torch.inference_mode():
            got32 = model(bags.to(cuda_device)).float().cpu()
        model.fc1_fp32 = False
        err32 = ((got32 - want).norm(dim=1) / want.norm(dim=1)).max().item()
        print(f"TransMIL, {n} tiles, all fp32: {err32:.2e}")
        assert err32 < 1e-4, (n, err32)
        # the reference's batch-wide scaling of the pseudo-inverse start: alone, bag 1 comes out differently
        with torch.inference_mode():
            alone = model(bags[1:].to(cuda_device)).float().cpu()
        assert (alone - got[1:]).abs().max() > 1e-5 or n == 300
