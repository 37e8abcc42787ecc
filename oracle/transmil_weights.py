"""Seeded synthetic weights for the TransMIL golden (TEST INFRASTRUCTURE, see oracle/make_golden_transmil.py): the same
function builds the state dict the reference module was run with and the one the tests load, so the fixture only has
to hold the reference's logits."""

import torch


def transmil_state_dict(dim_output: int, dim_input: int, dim_hidden: int, seed: int = 3) -> dict[str, torch.Tensor]:
    g = torch.Generator().manual_seed(seed)
    d = dim_hidden

    def w(*shape, fan_in):
        return torch.randn(*shape, generator=g) * (1.0 / fan_in ** 0.5)

    sd = {
        "cls_token": torch.randn(1, 1, d, generator=g),
        "_fc1.0.weight": w(d, dim_input, fan_in=dim_input), "_fc1.0.bias": 0.1 * torch.randn(d, generator=g),
        "norm.weight": 1.0 + 0.1 * torch.randn(d, generator=g), "norm.bias": 0.1 * torch.randn(d, generator=g),
        "_fc2.weight": w(dim_output, d, fan_in=d), "_fc2.bias": 0.1 * torch.randn(dim_output, generator=g),
    }
    for name, ks in (("proj", 7), ("proj1", 5), ("proj2", 3)):
        sd[f"pos_layer.{name}.weight"] = w(d, 1, ks, ks, fan_in=ks * ks)
        sd[f"pos_layer.{name}.bias"] = 0.1 * torch.randn(d, generator=g)
    for layer in ("layer1", "layer2"):
        sd[f"{layer}.norm.weight"] = 1.0 + 0.1 * torch.randn(d, generator=g)
        sd[f"{layer}.norm.bias"] = 0.1 * torch.randn(d, generator=g)
        sd[f"{layer}.attn.to_qkv.weight"] = w(3 * d, d, fan_in=d)
        sd[f"{layer}.attn.to_out.0.weight"] = w(d, d, fan_in=d)
        sd[f"{layer}.attn.to_out.0.bias"] = 0.1 * torch.randn(d, generator=g)
        sd[f"{layer}.attn.res_conv.weight"] = w(8, 1, 33, 1, fan_in=33)
    return sd
