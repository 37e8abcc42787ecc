"""Generates tests/golden/mil_long_*.npz: the REFERENCE module at the benched shape (S = 4097).

Run in the build container only (``python oracle/make_golden_long.py``).  The default-size model
(1024 -> 512, 8 heads, 2 layers, ff 512: 3.68 M parameters) and a 4096 x 1024 bag are too large to
commit, so the fixture stores what regenerates them -- the seeds of ``mil_oracle.init_state_dict`` and
``mil_oracle.synthetic_bag`` -- plus fp64 checksums of every regenerated tensor (a drifted generator
fails the test instead of silently comparing different inputs) and the logits the reference's
``VisionTransformer`` (src/stamp/modeling/models/vision_tranformer.py:298-384, imported by file path)
returns for them in eval mode.  These pin the tcgen05 long-bag attention kernel directly to
``_ALiBi.forward`` (:42-74) / ``nn.MultiheadAttention`` (:218-228) instead of transitively through
the oracle.
"""

from __future__ import annotations

import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

from oracle import mil_oracle  # noqa: E402
from oracle.make_golden import OUT, load_reference  # noqa: E402

N_TILES = 4096
CASES = [
    # name, use_alibi, sd seed, bag seed, batch, running_mean
    ("alibi_4096", True, 21, 4096, 1, 6000.0),
    ("mha_4096", False, 22, 4097, 1, 1.0),
]
DIMS = dict(dim_input=1024, dim_output=3, dim_model=512, n_layers=2, n_heads=8, dim_feedforward=512)


def checksums(sd: dict[str, torch.Tensor], bags: torch.Tensor, coords: torch.Tensor) -> np.ndarray:
    """[sum, sum of squares] in fp64 of every tensor in key order, then of bags and coords."""
    rows = [(v.double().sum().item(), (v.double() ** 2).sum().item()) for _, v in sorted(sd.items())]
    rows += [(t.double().sum().item(), (t.double() ** 2).sum().item()) for t in (bags, coords)]
    return np.asarray(rows, dtype=np.float64)


def regenerate(use_alibi: bool, sd_seed: int, bag_seed: int, batch: int, running_mean: float):
    sd = mil_oracle.init_state_dict(use_alibi=use_alibi, seed=sd_seed, running_mean=running_mean, **DIMS)
    bags, coords = mil_oracle.synthetic_bag(N_TILES, DIMS["dim_input"], seed=bag_seed, batch=batch)
    return sd, bags, coords


def main() -> None:
    ref = load_reference()
    torch.set_num_threads(max(1, torch.get_num_threads()))
    for name, use_alibi, sd_seed, bag_seed, batch, rm in CASES:
        sd, bags, coords = regenerate(use_alibi, sd_seed, bag_seed, batch, rm)
        model = ref.VisionTransformer(dropout=0.25, use_alibi=use_alibi, **DIMS).eval()
        model.load_state_dict(sd, strict=True)
        with torch.no_grad():
            out = model(bags, coords=coords, mask=None)
        np.savez_compressed(OUT / f"mil_long_{name}.npz", logits=out.numpy(), checksums=checksums(sd, bags, coords),
                            use_alibi=np.bool_(use_alibi), sd_seed=np.int64(sd_seed), bag_seed=np.int64(bag_seed),
                            batch=np.int64(batch), running_mean=np.float64(rm), n_tiles=np.int64(N_TILES))
        print(name, out.tolist())


if __name__ == "__main__":
    main()
