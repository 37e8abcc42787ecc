"""Golden for the TransMIL aggregator: the reference's own module (src/stamp/modeling/models/trans_mil.py, imported by
path: its imports -- beartype, jaxtyping, einops -- exist here) on seeded inputs -> tests/golden/transmil.npz
(the reference's logits for bags of 300 and 1100 tiles; weights and inputs are regenerated from seeds, see
oracle/transmil_weights.py).

    python oracle/make_golden_transmil.py        # needs /root/reference
"""
import importlib.util
from pathlib import Path

import sys

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from oracle.transmil_weights import transmil_state_dict  # noqa: E402

SRC = Path("/root/reference/src/stamp/modeling/models/trans_mil.py")


def main() -> None:
    spec = importlib.util.spec_from_file_location("ref_trans_mil", SRC)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    model = mod.TransMIL(dim_output=3, dim_input=64, dim_hidden=512).eval()
    model.load_state_dict(transmil_state_dict(3, 64, 512), strict=True)
    g = torch.Generator().manual_seed(9)
    arrays = {}
    with torch.no_grad():
        for n in (300, 1100):
            bags = torch.randn(2, n, 64, generator=g).half().float()
            arrays[f"bags_checksum_{n}"] = np.array(bags.double().sum().item())
            arrays[f"logits_{n}"] = model(bags).numpy()
    # the shape of the reference's own unit test (tests/test_model.py:135-166): 7 bags of 76 tiles, 457 input features, 4 classes
    model = mod.TransMIL(dim_output=4, dim_input=457, dim_hidden=512).eval()
    model.load_state_dict(transmil_state_dict(4, 457, 512), strict=True)
    with torch.no_grad():
        bags = torch.rand(7, 76, 457, generator=g).half().float()
        arrays["bags_checksum_odd"] = np.array(bags.double().sum().item())
        arrays["logits_odd"] = model(bags, coords=torch.rand(7, 76, 2), mask=torch.rand(7, 76) > 0.5).numpy()
    dst = Path(__file__).resolve().parent.parent / "tests" / "golden" / "transmil.npz"
    np.savez_compressed(dst, **arrays)
    print(dst, dst.stat().st_size, arrays["logits_300"], arrays["logits_1100"])


if __name__ == "__main__":
    main()
