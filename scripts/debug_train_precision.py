"""Development aid: where does the gradient error at long bags come from?  One bag, default model, N in argv;
compares class_token / bias-like gradients against the fp32 oracle with the tcgen05 and the mma.sync attention
kernels (stamp_b200_attention_tc_enable 1 / 0)."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch

from oracle import mil_oracle
from stamp_b200 import _lib, train as T
from stamp_b200.mil import VisionTransformer

dev = torch.device("cuda:0")
keys = ["class_token", "project_features.0.bias", "project_features.0.weight", "transformer.layers.0.0.norm.bias",
        "transformer.layers.1.0.norm.bias", "transformer.layers.1.0.mhsa.value_encoders.3.bias",
        "transformer.layers.0.0.mhsa.value_encoders.3.bias", "transformer.layers.0.0.mhsa.value_encoders.3.weight",
        "transformer.layers.0.0.mhsa.fc.weight", "transformer.layers.0.1.1.weight", "transformer.layers.1.0.mhsa.fc.bias"]
for N in [int(a) for a in sys.argv[1:]] or [1024, 2048, 4096]:
    sd = mil_oracle.init_state_dict(dim_input=1024, dim_output=2, seed=5)
    bags, coords = mil_oracle.synthetic_bag(N, 1024, seed=4242, batch=1, signal=True)
    targets = torch.tensor([[1.0, 0.0]])
    sd2 = mil_oracle.running_mean_update(sd, coords)
    params = {k: v.detach().clone().requires_grad_(True) for k, v in sd2.items() if "scale_distance" not in k}
    logits_ref = mil_oracle.forward({**sd2, **params}, bags, coords, None, exact_dist=True)
    mil_oracle.cross_entropy(logits_ref, targets, None).backward()
    for mode in (1, 0):
        _lib.load().stamp_b200_attention_tc_enable(mode)
        m = VisionTransformer(dim_output=2, dim_input=1024, dim_model=512, n_layers=2, n_heads=8, dim_feedforward=512,
                              dropout=0.0, use_alibi=True)
        m.load_state_dict(sd)
        for _, ff in m.transformer.layers:
            ff[3].p = 0.0; ff[5].p = 0.0
        m = m.to(dev).train()
        out = m(bags.to(dev), coords=coords.to(dev), mask=None)
        T.cross_entropy(out, targets.to(dev), None).backward()
        lerr = float((out.detach().cpu() - logits_ref.detach()).norm() / logits_ref.detach().norm())
        row = []
        for k in keys:
            g, r = dict(m.named_parameters())[k].grad.double().cpu().flatten(), params[k].grad.double().flatten()
            row.append(f"{float((g - r).norm() / r.norm()):.2e}")
        print(f"N={N} tc={mode} logits_err={lerr:.2e} " + " ".join(row), flush=True)
    _lib.load().stamp_b200_attention_tc_enable(1)
print("keys:", keys)
