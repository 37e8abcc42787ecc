"""CPU checks of the training-step oracle against the reference's own training step.

tests/golden/mil_train_step.npz was written by oracle/make_golden_train.py from the REFERENCE module
(train mode, dropout p = 0): logits, loss, every parameter gradient, the updated running means and the
parameters after one torch.optim.AdamW step.  The fp64 oracle (oracle/mil_oracle.train_grads) must
reproduce them to fp32 round-off; the GPU tests then compare the CUDA path with both."""

from pathlib import Path

import numpy as np
import pytest
import torch

from oracle import mil_oracle

GOLD_DIR = Path(__file__).resolve().parent / "golden"
TRAIN_GOLDENS = ["mil_train_step", "mil_train_step_mha"]


def load_train_golden(name: str = "mil_train_step"):
    z = np.load(GOLD_DIR / f"{name}.npz")
    sd = {k[3:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("sd/")}
    grads = {k[5:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("grad/")}
    after = {k[6:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("after/")}
    t = lambda k: torch.from_numpy(z[k])
    return dict(sd=sd, grads=grads, after=after, bags=t("bags"), coords=t("coords"), targets=t("targets"),
                class_weights=t("class_weights"), logits=t("logits"), loss=t("loss"), n_heads=int(z["n_heads"]),
                use_alibi=bool(z["use_alibi"]) if "use_alibi" in z.files else True)


@pytest.mark.parametrize("name", TRAIN_GOLDENS)
def test_train_oracle_matches_reference_step(name):
    g = load_train_golden(name)
    logits, loss, grads, sd2 = mil_oracle.train_grads(g["sd"], g["bags"], g["coords"], g["targets"],
                                                      g["class_weights"], n_heads=g["n_heads"])
    assert torch.allclose(logits.float(), g["logits"], rtol=2e-4, atol=2e-5)
    assert abs(float(loss) - float(g["loss"])) < 1e-5
    assert set(grads) == set(g["grads"])
    for k, ref in g["grads"].items():
        # reference: fp32 with matmul-based cdist; oracle: fp64, exact distances.  The key-encoder bias
        # gradients are analytically zero (softmax is shift invariant): both sides hold round-off only.
        err = (grads[k].float() - ref).norm()
        assert err < 2e-3 * ref.norm() + 1e-6 * ref.numel() ** 0.5, (k, float(err), float(ref.norm()))
    for k, v in g["after"].items():
        if "scale_distance" in k:
            assert torch.allclose(sd2[k].float(), v, rtol=1e-5), k


def test_dropout_sites_change_the_oracle_forward():
    g = load_train_golden()
    B, n = g["bags"].shape[:2]
    d, ff = 128, 128
    gen = torch.Generator().manual_seed(0)
    masks = {0: torch.rand(B * n * d, generator=gen) > 0.25}
    for l in range(2):
        masks[1 + 2 * l] = torch.rand(B * (n + 1) * ff, generator=gen) > 0.5
        masks[2 + 2 * l] = torch.rand(B * (n + 1) * d, generator=gen) > 0.5
    a = mil_oracle.forward(g["sd"], g["bags"], g["coords"], None)
    b = mil_oracle.forward(g["sd"], g["bags"], g["coords"], None, drop_masks=masks, p_proj=0.25, p_ff=0.5)
    c = mil_oracle.forward(g["sd"], g["bags"], g["coords"], None, drop_masks=masks, p_proj=0.0, p_ff=0.0)
    assert not torch.allclose(a, b)
    assert torch.allclose(a, c)


def test_cross_entropy_restatement_matches_torch():
    g = load_train_golden()
    ref = torch.nn.functional.cross_entropy(g["logits"], g["targets"], weight=g["class_weights"])
    assert torch.allclose(mil_oracle.cross_entropy(g["logits"], g["targets"], g["class_weights"]), ref, rtol=1e-6)
