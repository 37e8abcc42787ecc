"""Golden for the survival loss: the reference's own ``neg_partial_log_likelihood`` (src/stamp/modeling/models/cox.py,
imported by path; it only needs torch) on seeded scores / times / events -> tests/golden/cox_loss.npz (inputs, loss,
and the gradient autograd gives through the reference code).

    python oracle/make_golden_cox.py        # needs /root/reference
"""
import importlib.util
import warnings
from pathlib import Path

import numpy as np
import torch

SRC = Path("/root/reference/src/stamp/modeling/models/cox.py")


def cases() -> dict[str, tuple[torch.Tensor, torch.Tensor, torch.Tensor, str]]:
    g = torch.Generator().manual_seed(31)
    out = {}
    s = torch.randn(64, generator=g)
    out["distinct_times"] = (s, torch.rand(64, generator=g) * 100, torch.rand(64, generator=g) < 0.6, "efron")
    t = torch.randint(1, 12, (50,), generator=g).float()
    out["ties_efron"] = (torch.randn(50, generator=g) * 2, t, torch.rand(50, generator=g) < 0.7, "efron")
    out["ties_breslow"] = (torch.randn(50, generator=g) * 2, t, torch.rand(50, generator=g) < 0.7, "breslow")
    out["all_tied"] = (torch.randn(9, generator=g), torch.full((9,), 3.0), torch.tensor([1, 1, 0, 1, 0, 1, 1, 0, 1]).bool(), "efron")
    out["large_scores"] = (torch.randn(33, generator=g) * 30 + 50, torch.rand(33, generator=g), torch.rand(33, generator=g) < 0.5, "efron")
    out["one_event"] = (torch.randn(7, generator=g), torch.rand(7, generator=g), torch.tensor([0, 0, 0, 1, 0, 0, 0]).bool(), "efron")
    return out


def main() -> None:
    spec = importlib.util.spec_from_file_location("ref_cox", SRC)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    arrays = {}
    for name, (s, t, e, ties) in cases().items():
        s = s.clone().requires_grad_(True)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            loss = mod.neg_partial_log_likelihood(s, t, e, ties_method=ties)
        loss.backward()
        arrays.update({f"{name}/log_hz": s.detach().numpy(), f"{name}/time": t.numpy(), f"{name}/event": e.numpy(),
                       f"{name}/loss": loss.detach().numpy(), f"{name}/grad": s.grad.numpy(), f"{name}/ties": np.array(ties)})
        print(name, float(loss))
    dst = Path(__file__).resolve().parent.parent / "tests" / "golden" / "cox_loss.npz"
    np.savez_compressed(dst, **arrays)
    print(dst, dst.stat().st_size)


if __name__ == "__main__":
    main()
