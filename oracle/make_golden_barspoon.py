"""Golden for the barspoon aggregator: the reference's own EncDecTransformer (src/stamp/modeling/models/barspoon.py:
24-170; the file imports lightning / torchmetrics at the top, so the class and `sanitize` are executed from their
source segments) on seeded inputs -> tests/golden/barspoon.npz (state dict, inputs, logits per target).

    python oracle/make_golden_barspoon.py        # needs /root/reference
"""
import ast
import re
from pathlib import Path

import numpy as np
import torch
from torch import nn

SRC = Path("/root/reference/src/stamp/modeling/models/barspoon.py")


def reference_class():
    src = SRC.read_text()
    ns = {"torch": torch, "nn": nn, "re": re, "F": torch.nn.functional}
    for node in ast.parse(src).body:
        if (isinstance(node, ast.ClassDef) and node.name == "EncDecTransformer") or \
                (isinstance(node, ast.FunctionDef) and node.name == "sanitize"):
            exec(ast.get_source_segment(src, node), ns)
    return ns["EncDecTransformer"]


def main() -> None:
    torch.manual_seed(11)
    targets = {"isMSIH": 2, "grade (WHO)": 3, "subtype": 4}
    model = reference_class()(d_features=64, target_n_outs=targets, d_model=128, num_encoder_heads=2, num_decoder_heads=2,
                              num_encoder_layers=2, num_decoder_layers=2, dim_feedforward=256).eval()
    g = torch.Generator().manual_seed(5)
    tokens = torch.randn(2, 333, 64, generator=g).half().float()
    positions = torch.rand(2, 333, 2, generator=g) * 30000.0
    with torch.no_grad():
        out = model(tokens, positions)
        nope = reference_class()(d_features=64, target_n_outs=targets, d_model=128, num_encoder_heads=2,
                                 num_decoder_heads=2, num_encoder_layers=2, num_decoder_layers=2, dim_feedforward=256,
                                 positional_encoding=False).eval()
        nope.load_state_dict(model.state_dict())
        out_nope = nope(tokens, positions)
    arrays = {f"sd::{k}": v.numpy() for k, v in model.state_dict().items()}
    arrays.update(tokens=tokens.numpy(), positions=positions.numpy(), labels=np.array(list(targets)),
                  n_outs=np.array(list(targets.values())))
    for k, v in out.items():
        arrays[f"logits::{k}"] = v.numpy()
    for k, v in out_nope.items():
        arrays[f"logits_nope::{k}"] = v.numpy()
    dst = Path(__file__).resolve().parent.parent / "tests" / "golden" / "barspoon.npz"
    np.savez_compressed(dst, **arrays)
    print(dst, dst.stat().st_size, {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
