"""Generates tests/golden/mil_gradcam_trained_scale.npz with the REFERENCE's own heatmap code.

Run in the build container only.  The reference ``VisionTransformer`` (vision_tranformer.py, imported by path) is
loaded with the state dict of the existing fixture ``mil_alibi_trained_scale.npz`` and handed to the reference's
``_gradcam_per_category`` (src/stamp/heatmaps/__init__.py:36-56) -- the function's source segment is executed from
the reference file (the module itself needs openslide / matplotlib), so ``jacrev`` over ``model.forward`` runs
exactly as ``stamp heatmaps`` runs it.  Stored: the class-activation map [N, C] and the pre-softmax scores
``|mean_d(feats * d logit_c / d feats)|`` [C, N] (the softmax over tiles flattens differences).
"""

from __future__ import annotations

import ast
import sys
from pathlib import Path
from typing import cast

import numpy as np
import torch
from torch import Tensor
from torch.func import jacrev

sys.path.insert(0, str(Path(__file__).resolve().parent))
from make_golden import DIMS, OUT, load_reference  # noqa: E402

HEATMAPS = Path("/root/reference/src/stamp/heatmaps/__init__.py")


def reference_gradcam():
    src = HEATMAPS.read_text()
    ns = {"torch": torch, "Tensor": Tensor, "cast": cast, "jacrev": jacrev}
    for node in ast.parse(src).body:
        if isinstance(node, ast.FunctionDef) and node.name == "_gradcam_per_category":
            exec("from __future__ import annotations\n" + ast.get_source_segment(src, node), ns)
    return ns["_gradcam_per_category"]


def main() -> None:
    ref = load_reference()
    z = np.load(OUT / "mil_alibi_trained_scale.npz")
    sd = {k[3:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("sd/")}
    model = ref.VisionTransformer(dropout=0.25, use_alibi=True, **DIMS).eval()
    model.load_state_dict(sd)
    feats, coords = torch.from_numpy(z["bags"])[0], torch.from_numpy(z["coords"])[0]
    cam = reference_gradcam()(model, feats, coords)
    jac = jacrev(lambda b: model.forward(b.unsqueeze(0), coords=coords.unsqueeze(0), mask=None).squeeze(0))(feats)
    scores = (feats * jac).mean(-1).abs()
    assert torch.allclose(torch.softmax(scores, dim=-1).permute(-1, -2), cam)
    np.savez_compressed(OUT / "mil_gradcam_trained_scale.npz", cam=cam.detach().numpy(), scores=scores.detach().numpy())
    print("cam", tuple(cam.shape), "scores range", float(scores.min()), float(scores.max()))


if __name__ == "__main__":
    main()
