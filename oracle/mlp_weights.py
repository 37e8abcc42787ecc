"""Seeded synthetic weights for the MLP / Linear aggregator golden (TEST INFRASTRUCTURE, see oracle/make_golden_mlp.py):
the same function builds the state dict the reference modules were run with and the one the tests load."""

import torch


def mlp_state_dict(dim_input: int, dim_hidden: int, dim_output: int, num_layers: int, seed: int = 5) -> dict[str, torch.Tensor]:
    """Keys of the reference's ``MLP`` (src/stamp/modeling/models/mlp.py:24-31): Linear at mlp.0, mlp.3, mlp.6, ..."""
    g = torch.Generator().manual_seed(seed)
    sd, in_dim = {}, dim_input
    for i in range(num_layers):
        out_dim = dim_output if i == num_layers - 1 else dim_hidden
        sd[f"mlp.{3 * i}.weight"] = torch.randn(out_dim, in_dim, generator=g) / in_dim ** 0.5
        sd[f"mlp.{3 * i}.bias"] = 0.1 * torch.randn(out_dim, generator=g)
        in_dim = out_dim
    return sd


def linear_state_dict(dim_input: int, dim_output: int, seed: int = 6) -> dict[str, torch.Tensor]:
    g = torch.Generator().manual_seed(seed)
    return {"fc.weight": torch.randn(dim_output, dim_input, generator=g) / dim_input ** 0.5,
            "fc.bias": 0.1 * torch.randn(dim_output, generator=g)}
