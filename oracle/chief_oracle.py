"""ORACLE (test infrastructure, not product code) -- CPU restatement of CHIEF's gated-attention
pooling and EAGLE's top-k selection in plain torch fp32/fp64.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline leg may import this.

Follows src/stamp/encoding/encoder/chief.py: ``CHIEFModel.__init__`` :27-65 (size 'small' =
[768, 512, 256]; attention_net = Sequential(Linear, ReLU, Dropout(0.25), Attn_Net_Gated)),
``CHIEFModel.forward`` :74-89, ``Attn_Net_Gated`` :255-275, ``initialize_weights`` :211-219
(xavier_normal weights, zero biases); and src/stamp/encoding/encoder/eagle.py:104-120.
Parity pin: ``chief.py`` cannot be imported as a package member offline (top-level ``gdown`` / ``stamp.*``
imports), but its model classes are plain torch: ``oracle/make_golden_chief.py`` loads the file by path with
inert stand-ins for those imports, runs the reference ``CHIEFModel(size_arg="small")`` and EAGLE's selection
lines on seeded inputs and commits the outputs as ``tests/golden/chief_pool.npz``; ``tests/test_oracle_cpu.py``
checks this restatement against them.  Weights are synthetic (the pretrained ones live on Google Drive);
state-dict keys follow the reference module tree.
"""

from __future__ import annotations

import torch
import torch.nn.functional as F
from torch import Tensor


def init_state_dict(sizes=(768, 512, 256), seed: int = 0, zero_bias: bool = False) -> dict[str, Tensor]:
    g = torch.Generator().manual_seed(seed)

    def xavier(out_f, in_f):
        std = (2.0 / (in_f + out_f)) ** 0.5
        return torch.randn(out_f, in_f, generator=g) * std

    def bias(n):
        return torch.zeros(n) if zero_bias else 0.05 * torch.randn(n, generator=g)

    D, L, Dh = sizes
    return {
        "attention_net.0.weight": xavier(L, D), "attention_net.0.bias": bias(L),
        "attention_net.3.attention_a.0.weight": xavier(Dh, L), "attention_net.3.attention_a.0.bias": bias(Dh),
        "attention_net.3.attention_b.0.weight": xavier(Dh, L), "attention_net.3.attention_b.0.bias": bias(Dh),
        "attention_net.3.attention_c.weight": xavier(1, Dh), "attention_net.3.attention_c.bias": bias(1),
    }


def forward(sd: dict[str, Tensor], x: Tensor) -> dict[str, Tensor]:
    """CHIEFModel.forward in eval mode (dropout inactive)."""
    sd = {k: v.to(x.dtype) for k, v in sd.items()}
    h = F.relu(F.linear(x, sd["attention_net.0.weight"], sd["attention_net.0.bias"]))
    a = torch.tanh(F.linear(h, sd["attention_net.3.attention_a.0.weight"], sd["attention_net.3.attention_a.0.bias"]))
    b = torch.sigmoid(F.linear(h, sd["attention_net.3.attention_b.0.weight"], sd["attention_net.3.attention_b.0.bias"]))
    A = F.linear(a * b, sd["attention_net.3.attention_c.weight"], sd["attention_net.3.attention_c.bias"])
    A_raw = A.transpose(1, 0)
    A = torch.softmax(A_raw, dim=1)
    return {"attention_raw": A_raw, "WSI_feature": A @ x, "WSI_feature_transformed": A @ h}


def eagle_embedding(sd: dict[str, Tensor], feats: Tensor, agg_feats: Tensor) -> tuple[Tensor, Tensor]:
    attn = forward(sd, feats)["attention_raw"].squeeze(0)
    k = min(25, attn.shape[0])
    _, idx = torch.topk(attn, k)
    return agg_feats[idx].mean(0), idx
