"""MLP / Linear aggregators (stamp_b200/mlp.py) against the reference's own modules: tests/golden/mlp.npz holds the outputs
of the reference MLP and Linear (imported by path in oracle/make_golden_mlp.py) for seeded weights; weights and inputs are
regenerated here from the same seeds (inputs verified by checksum).  The bag mean is additionally checked against fp64
at a full-size bag batch."""

from pathlib import Path

import numpy as np
import pytest
import torch

GOLD = Path(__file__).parent / "golden" / "mlp.npz"


def _models(F: int):
    from oracle.mlp_weights import linear_state_dict, mlp_state_dict
    from stamp_b200.mlp import MLP, Linear

    mlp = MLP(dim_input=F, dim_hidden=64, dim_output=4, num_layers=3, dropout=0.25)
    lin = Linear(dim_input=F, dim_output=4)
    assert set(mlp.state_dict()) == set(mlp_state_dict(F, 64, 4, 3))
    mlp.load_state_dict(mlp_state_dict(F, 64, 4, 3), strict=True)
    lin.load_state_dict(linear_state_dict(F, 4), strict=True)
    return mlp.eval(), lin.eval()


def test_mlp_state_dict_is_the_reference_one_and_cpu_raises():
    mlp, lin = _models(96)
    with torch.no_grad():
        for m in (mlp, lin):
            with pytest.raises(RuntimeError):
                m(torch.zeros(2, 5, 96))
            with pytest.raises(ValueError):             # mlp.py:42-43 / :59-60
                m(torch.zeros(2, 3, 5, 96))
    with pytest.raises(NotImplementedError):
        mlp(torch.zeros(2, 96))


@pytest.mark.gpu
def test_mlp_and_linear_match_reference_golden(cuda_device):
    from oracle.make_golden_mlp import CASES, inputs

    z = np.load(GOLD)
    g = torch.Generator().manual_seed(21)
    for name, shape in CASES.items():
        mlp, lin = (m.to(cuda_device) for m in _models(shape[-1]))
        x = inputs(name, g)
        assert abs(x.double().sum().item() - float(z[f"checksum_{name}"])) < 1e-9          # the golden's inputs
        for dtype in (torch.float32, torch.float16):                                       # the inputs are fp16-representable
            with torch.inference_mode():
                xd = x.to(cuda_device, dtype)
                for tag, model in (("mlp", mlp), ("linear", lin)):
                    got = model(xd).cpu()
                    want = torch.from_numpy(z[f"{tag}_{name}"])
                    err = ((got - want).norm() / want.norm()).item()
                    print(f"{tag} {name} {dtype}: relative error {err:.2e}")
                    assert got.shape == want.shape and got.dtype == torch.float32 and err < 2e-6, (tag, name, dtype, err)


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(8, 4096, 1024), (1, 4096, 1536), (64, 512, 768), (3, 1, 8), (2, 1000, 1000), (2, 33, 7)])
def test_bag_mean_matches_fp64(cuda_device, shape):
    from stamp_b200.mlp import bag_mean

    g = torch.Generator().manual_seed(shape[1])
    x = torch.randn(*shape, generator=g) + 0.5
    for dtype in (torch.float32, torch.float16):
        xd = x.to(cuda_device, dtype)
        got = bag_mean(xd)
        want = xd.double().mean(dim=1)
        assert got.shape == want.shape and got.dtype == torch.float32
        assert (got.double() - want).abs().max().item() < 4e-6
        assert torch.equal(got, bag_mean(xd))                                               # deterministic (no atomics)
    view = x.to(cuda_device)[:, ::2]                                                        # strided rows
    if view.shape[1]:
        assert (bag_mean(view).double() - view.double().mean(dim=1)).abs().max().item() < 4e-6
    if shape[2] > 8:                                                                        # a column window (unaligned rows)
        win = x.to(cuda_device)[:, :, 1:shape[2] - 2]
        assert (bag_mean(win).double() - win.double().mean(dim=1)).abs().max().item() < 4e-6
