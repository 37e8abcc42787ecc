"""Golden for the MLP / Linear aggregators: the reference's own modules (src/stamp/modeling/models/mlp.py, imported by
path: beartype and jaxtyping exist here) on seeded inputs -> tests/golden/mlp.npz (the reference's outputs for bags
[3, 777, 96], ragged widths [2, 301, 50] and plain feature vectors [5, 96]; weights and inputs are regenerated from
seeds, see oracle/mlp_weights.py).

    python oracle/make_golden_mlp.py        # needs /root/reference
"""
import importlib.util
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from oracle.mlp_weights import linear_state_dict, mlp_state_dict  # noqa: E402

SRC = Path("/root/reference/src/stamp/modeling/models/mlp.py")
CASES = {"bags": (3, 777, 96), "odd": (2, 301, 50), "vectors": (5, 96)}


def inputs(name: str, g: torch.Generator) -> torch.Tensor:
    return torch.randn(*CASES[name], generator=g).half().float()


def main() -> None:
    spec = importlib.util.spec_from_file_location("ref_mlp", SRC)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    g = torch.Generator().manual_seed(21)
    arrays = {}
    with torch.no_grad():
        for name, shape in CASES.items():
            F = shape[-1]
            mlp = mod.MLP(dim_input=F, dim_hidden=64, dim_output=4, num_layers=3, dropout=0.25).eval()
            mlp.load_state_dict(mlp_state_dict(F, 64, 4, 3), strict=True)
            lin = mod.Linear(dim_input=F, dim_output=4).eval()
            lin.load_state_dict(linear_state_dict(F, 4), strict=True)
            x = inputs(name, g)
            arrays[f"checksum_{name}"] = np.array(x.double().sum().item())
            arrays[f"mlp_{name}"] = mlp(x).numpy()
            arrays[f"linear_{name}"] = lin(x).numpy()
    dst = Path(__file__).resolve().parent.parent / "tests" / "golden" / "mlp.npz"
    np.savez_compressed(dst, **arrays)
    print(dst, dst.stat().st_size, {k: v.shape for k, v in arrays.items()})


if __name__ == "__main__":
    main()
