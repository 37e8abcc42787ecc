"""Generates tests/golden/mil_*.npz by running the REFERENCE module itself.

Run in the build container only (``python oracle/make_golden.py``): it imports
``/root/reference/src/stamp/modeling/models/vision_tranformer.py`` by file path (the file needs only
torch / einops / beartype / jaxtyping), instantiates the reference ``VisionTransformer`` with a fixed
seed, runs it in eval mode on seeded synthetic bags and stores inputs, the reference state dict and
the reference outputs.  The fixtures pin ``oracle/mil_oracle.py`` (tests/test_oracle_cpu.py) and are
compared with the CUDA path on the GPU box, where /root/reference does not exist.
"""

from __future__ import annotations

import importlib.util
import sys
from pathlib import Path

import numpy as np
import torch

REF = Path("/root/reference/src/stamp/modeling/models/vision_tranformer.py")
OUT = Path(__file__).resolve().parent.parent / "tests" / "golden"


def load_reference():
    spec = importlib.util.spec_from_file_location("ref_vision_tranformer", REF)
    mod = importlib.util.module_from_spec(spec)
    sys.modules["ref_vision_tranformer"] = mod
    spec.loader.exec_module(mod)
    return mod


CASES = [
    # name, use_alibi, n_tiles, masked, batch, running_mean
    ("alibi_nomask", True, 65, False, 2, 1.0),
    ("alibi_mask", True, 70, True, 3, 1.0),
    ("alibi_trained_scale", True, 130, False, 1, 6000.0),
    ("mha_nomask", False, 65, False, 2, 1.0),
    ("mha_mask", False, 70, True, 3, 1.0),
]
DIMS = dict(dim_input=64, dim_model=128, n_layers=2, n_heads=2, dim_feedforward=128, dim_output=3)


def main() -> None:
    ref = load_reference()
    OUT.mkdir(parents=True, exist_ok=True)
    for i, (name, use_alibi, n, masked, batch, rm) in enumerate(CASES):
        torch.manual_seed(100 + i)
        model = ref.VisionTransformer(dropout=0.25, use_alibi=use_alibi, **DIMS).eval()
        if use_alibi and rm != 1.0:
            for layer in model.transformer.layers:
                for att in layer[0].mhsa.attentions:
                    att.scale_distance.running_mean.fill_(rm)
        g = torch.Generator().manual_seed(1000 + i)
        bags = torch.randn(batch, n, DIMS["dim_input"], generator=g).half().float()
        cells = torch.stack([torch.randperm(100 * 100, generator=g)[:n] for _ in range(batch)])
        coords = torch.stack([(cells % 100).float(), (cells // 100).float()], dim=-1) * 256.0
        mask = None
        if masked:
            # like a zero-padded batch: the tail of each bag is masked (bag_sizes differ)
            sizes = torch.randint(n // 2, n, (batch,), generator=g)
            mask = torch.arange(n)[None, :] >= sizes[:, None]
        with torch.no_grad():
            out = model(bags, coords=coords, mask=mask)
        arrays = {f"sd/{k}": v.numpy() for k, v in model.state_dict().items()}
        arrays.update(bags=bags.numpy(), coords=coords.numpy(), logits=out.numpy(),
                      n_heads=np.int64(DIMS["n_heads"]), use_alibi=np.bool_(use_alibi))
        if mask is not None:
            arrays["mask"] = mask.numpy()
        np.savez_compressed(OUT / f"mil_{name}.npz", **arrays)
        print(name, tuple(out.shape), out[0].tolist())


if __name__ == "__main__":
    main()
