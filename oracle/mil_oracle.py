"""ORACLE (test infrastructure, not product code) -- CPU restatement of STAMP's ALiBi
Transformer-MIL aggregator in plain torch fp32/fp64.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
legs may import this module, and only as the checker / the timed CPU baseline.

Follows, function by function (paths relative to the reference root):
  * ``_RunningMeanScaler.forward``  src/stamp/modeling/models/vision_tranformer.py:23-31
  * ``_ALiBi.forward``              :42-74    (bias subtracted AFTER the softmax; masks applied
                                               after the softmax with ``where``)
  * ``MultiHeadALiBi.forward``      :123-154  (H separate Linear(d, d/H) for q, k, v; head-major concat)
  * ``feed_forward``                :157-169
  * ``SelfAttention.forward``       :194-242  (LayerNorm first; nn.MultiheadAttention when use_alibi=False)
  * ``Transformer.forward``         :281-295
  * ``VisionTransformer.forward``   :332-384  (class token + (0,0) coordinate; mask construction)

Parity pin: ``tests/golden/mil_*.npz`` hold outputs of the reference module itself (imported by file
path from /root/reference by ``oracle/make_golden.py`` in the build container); ``tests/test_oracle_cpu.py``
checks this restatement against them.  The state-dict keys are the reference's.
"""

from __future__ import annotations

import math

import torch
import torch.nn.functional as F
from torch import Tensor


def init_state_dict(*, dim_input: int, dim_output: int, dim_model: int = 512, n_layers: int = 2,
                    n_heads: int = 8, dim_feedforward: int = 512, use_alibi: bool = True,
                    seed: int = 0, running_mean: float = 1.0) -> dict[str, Tensor]:
    """Seeded synthetic weights with the reference's state-dict layout."""
    g = torch.Generator().manual_seed(seed)

    def lin(out_f, in_f):
        bound = 1.0 / math.sqrt(in_f)
        w = (torch.rand(out_f, in_f, generator=g) * 2 - 1) * bound
        b = (torch.rand(out_f, generator=g) * 2 - 1) * bound
        return w, b

    def ln(d):
        return 1.0 + 0.1 * torch.randn(d, generator=g), 0.05 * torch.randn(d, generator=g)

    sd: dict[str, Tensor] = {}
    sd["class_token"] = torch.randn(dim_model, generator=g)
    sd["project_features.0.weight"], sd["project_features.0.bias"] = lin(dim_model, dim_input)
    hd = dim_model // n_heads
    for l in range(n_layers):
        p = f"transformer.layers.{l}."
        sd[p + "0.norm.weight"], sd[p + "0.norm.bias"] = ln(dim_model)
        if use_alibi:
            for name in ("query", "key", "value"):
                for h in range(n_heads):
                    w, b = lin(hd, dim_model)
                    sd[p + f"0.mhsa.{name}_encoders.{h}.weight"] = w
                    sd[p + f"0.mhsa.{name}_encoders.{h}.bias"] = b
            for h in range(n_heads):
                sd[p + f"0.mhsa.attentions.{h}.bias_scale"] = torch.rand(1, generator=g)
                sd[p + f"0.mhsa.attentions.{h}.scale_distance.running_mean"] = torch.full((1,), float(running_mean))
                sd[p + f"0.mhsa.attentions.{h}.scale_distance.items_so_far"] = torch.ones(1)
            sd[p + "0.mhsa.fc.weight"], sd[p + "0.mhsa.fc.bias"] = lin(dim_model, dim_model)
        else:
            w, b = lin(3 * dim_model, dim_model)
            sd[p + "0.mhsa.in_proj_weight"], sd[p + "0.mhsa.in_proj_bias"] = w, b
            sd[p + "0.mhsa.out_proj.weight"], sd[p + "0.mhsa.out_proj.bias"] = lin(dim_model, dim_model)
        sd[p + "1.0.weight"], sd[p + "1.0.bias"] = ln(dim_model)
        sd[p + "1.1.weight"], sd[p + "1.1.bias"] = lin(dim_feedforward, dim_model)
        sd[p + "1.4.weight"], sd[p + "1.4.bias"] = lin(dim_model, dim_feedforward)
    sd["transformer.norm.weight"], sd["transformer.norm.bias"] = ln(dim_model)
    sd["mlp_head.0.weight"], sd["mlp_head.0.bias"] = lin(dim_output, dim_model)
    return sd


def _n_layers(sd: dict[str, Tensor]) -> int:
    return 1 + max(int(k.split(".")[2]) for k in sd if k.startswith("transformer.layers."))


def _n_heads_alibi(sd: dict[str, Tensor], layer: int) -> int:
    p = f"transformer.layers.{layer}.0.mhsa.query_encoders."
    return 1 + max(int(k[len(p):].split(".")[0]) for k in sd if k.startswith(p))


def alibi_attention(q, k, v, coords_q, coords_k, bias_scale, running_mean, attn_mask, alibi_mask,
                    exact_dist: bool = False):
    """_ALiBi.forward (eval mode), vision_tranformer.py:58-74."""
    logits = torch.einsum("bqf,bkf->bqk", q, k) * (k.size(-1) ** -0.5)
    if exact_dist:
        dist = (coords_q[:, :, None, :] - coords_k[:, None, :, :]).norm(dim=-1)
    else:
        dist = torch.cdist(coords_q, coords_k)  # the reference's (matmul-based) distances
    sd = dist / running_mean * bias_scale
    if alibi_mask is not None:
        sd = sd.where(~alibi_mask, 0.0)
    w = torch.softmax(logits, dim=-1)
    if attn_mask is not None:
        w = (w - sd).where(~attn_mask, 0.0)
    else:
        w = w - sd
    return torch.einsum("bqk,bkf->bqf", w, v)


def _drop(x: Tensor, keep: Tensor | None, p: float) -> Tensor:
    """nn.Dropout in training mode with an explicit keep mask (flattened row-major like x)."""
    if keep is None or p <= 0.0:
        return x
    return x * keep.reshape(x.shape).to(x.dtype) / (1.0 - p)


def forward(sd: dict[str, Tensor], bags: Tensor, coords: Tensor, mask: Tensor | None, *,
            n_heads: int | None = None, exact_dist: bool = False,
            drop_masks: dict[int, Tensor] | None = None, p_proj: float = 0.0, p_ff: float = 0.0) -> Tensor:
    """VisionTransformer.forward. Returns logits [B, C].

    Default: eval mode (dropout inactive).  With ``drop_masks`` the three nn.Dropout sites run in
    training mode with the given keep masks (site 0: project_features :314-318; 1+2l / 2+2l: the two
    Dropouts of feed_forward :157-169 in layer l) -- differentiable, used as the gradient oracle.
    Running-mean handling of training mode is the caller's (see ``running_mean_update``)."""
    dt = bags.dtype
    sd = {k: v.to(dt) if v.is_floating_point() else v for k, v in sd.items()}
    dm = drop_masks or {}
    B = bags.shape[0]
    use_alibi = any(".query_encoders." in k for k in sd)
    x = F.gelu(F.linear(bags, sd["project_features.0.weight"], sd["project_features.0.bias"]))
    x = _drop(x, dm.get(0), p_proj)
    d = x.shape[-1]
    x = torch.cat([sd["class_token"].expand(B, 1, d), x], dim=1)
    coords = torch.cat([torch.zeros(B, 1, 2, dtype=coords.dtype, device=coords.device), coords], dim=1)

    attn_mask = alibi_mask = None
    if mask is not None:
        m = torch.cat([torch.zeros(B, 1, dtype=torch.bool, device=mask.device), mask], dim=1)
        attn_mask = m[:, :, None] & m[:, None, :]            # einsum("bq,bk->bqk") on bools
        attn_mask[:, 1:, 0] = True
        alibi_mask = torch.zeros_like(attn_mask)
        alibi_mask[:, 0, :] = True
        alibi_mask[:, :, 0] = True

    for l in range(_n_layers(sd)):
        p = f"transformer.layers.{l}."
        xn = F.layer_norm(x, (d,), sd[p + "0.norm.weight"], sd[p + "0.norm.bias"], 1e-5)
        if use_alibi:
            H = _n_heads_alibi(sd, l)
            heads = []
            for h in range(H):
                q = F.linear(xn, sd[p + f"0.mhsa.query_encoders.{h}.weight"], sd[p + f"0.mhsa.query_encoders.{h}.bias"])
                k = F.linear(xn, sd[p + f"0.mhsa.key_encoders.{h}.weight"], sd[p + f"0.mhsa.key_encoders.{h}.bias"])
                v = F.linear(xn, sd[p + f"0.mhsa.value_encoders.{h}.weight"], sd[p + f"0.mhsa.value_encoders.{h}.bias"])
                heads.append(alibi_attention(
                    q, k, v, coords, coords, sd[p + f"0.mhsa.attentions.{h}.bias_scale"],
                    sd[p + f"0.mhsa.attentions.{h}.scale_distance.running_mean"], attn_mask, alibi_mask,
                    exact_dist=exact_dist))
            att = torch.stack(heads).permute(1, 2, 0, 3).flatten(-2, -1)
            att = F.linear(att, sd[p + "0.mhsa.fc.weight"], sd[p + "0.mhsa.fc.bias"])
        else:
            if n_heads is None:
                raise ValueError("n_heads is required for the nn.MultiheadAttention variant")
            S = x.shape[1]
            hd = d // n_heads
            qkv = F.linear(xn, sd[p + "0.mhsa.in_proj_weight"], sd[p + "0.mhsa.in_proj_bias"])
            q, k, v = qkv.view(B, S, 3, n_heads, hd).permute(2, 0, 3, 1, 4)
            logits = q @ k.transpose(-1, -2) / math.sqrt(hd)
            if attn_mask is not None:
                # Reference quirk kept on purpose (vision_tranformer.py:222-226): the [B,S,S] mask is
                # expanded with .repeat(H,1,1) -> rows ordered (head, bag), but nn.MultiheadAttention
                # reads its [B*H,S,S] mask as (bag, head): bag b / head h gets the mask of bag (b*H+h) % B.
                idx = (torch.arange(B, device=x.device)[:, None] * n_heads + torch.arange(n_heads, device=x.device)[None, :]) % B
                logits = logits.masked_fill(attn_mask[idx], float("-inf"))
            o = (torch.softmax(logits, -1) @ v).permute(0, 2, 1, 3).reshape(B, S, d)
            att = F.linear(o, sd[p + "0.mhsa.out_proj.weight"], sd[p + "0.mhsa.out_proj.bias"])
        x = att + x
        h1 = F.layer_norm(x, (d,), sd[p + "1.0.weight"], sd[p + "1.0.bias"], 1e-5)
        h1 = _drop(F.gelu(F.linear(h1, sd[p + "1.1.weight"], sd[p + "1.1.bias"])), dm.get(1 + 2 * l), p_ff)
        x = _drop(F.linear(h1, sd[p + "1.4.weight"], sd[p + "1.4.bias"]), dm.get(2 + 2 * l), p_ff) + x
    x = F.layer_norm(x, (d,), sd["transformer.norm.weight"], sd["transformer.norm.bias"], 1e-5)
    return F.linear(x[:, 0], sd["mlp_head.0.weight"], sd["mlp_head.0.bias"])


def running_mean_update(sd: dict[str, Tensor], coords: Tensor) -> dict[str, Tensor]:
    """Training-mode side effect of _RunningMeanScaler.forward (vision_tranformer.py:23-31) on every
    head of every layer: rm <- mean(rm + (dist - rm) / n); n += 1, dist = cdist over tokens incl. the
    class token at (0,0).  Returns an updated copy of the state dict."""
    B = coords.shape[0]
    c = torch.cat([torch.zeros(B, 1, 2, dtype=coords.dtype, device=coords.device), coords], dim=1)
    dist = torch.cdist(c, c)
    out = dict(sd)
    for k in sd:
        if k.endswith("scale_distance.running_mean"):
            nk = k.replace("running_mean", "items_so_far")
            out[k] = (sd[k] + (dist - sd[k]) / sd[nk]).mean().reshape(1)
            out[nk] = sd[nk] + 1
    return out


def cross_entropy(logits: Tensor, targets: Tensor, class_weights: Tensor | None) -> Tensor:
    """LitTileClassifier._step loss (src/stamp/modeling/models/__init__.py:254-258): soft one-hot
    targets, class weights, mean over the batch = mean_b(-sum_c w_c y_bc log p_bc)."""
    logp = torch.log_softmax(logits, dim=1)
    w = class_weights if class_weights is not None else torch.ones(logits.shape[1], dtype=logits.dtype, device=logits.device)
    return -(w[None, :] * targets * logp).sum(dim=1).mean()


def train_grads(sd: dict[str, Tensor], bags: Tensor, coords: Tensor, targets: Tensor,
                class_weights: Tensor | None, *, drop_masks=None, p_proj: float = 0.0, p_ff: float = 0.0,
                dtype=torch.float64, n_heads: int | None = None):
    """One training-mode forward/backward: returns (logits, loss, grads by state-dict key, updated sd)."""
    sd2 = running_mean_update(sd, coords)
    params = {k: v.detach().to(dtype).requires_grad_(True) for k, v in sd2.items()
              if "scale_distance" not in k}
    full = {**{k: v.to(dtype) for k, v in sd2.items()}, **params}
    logits = forward(full, bags.to(dtype), coords.to(dtype), None, drop_masks=drop_masks, p_proj=p_proj,
                     p_ff=p_ff, exact_dist=True, n_heads=n_heads)
    loss = cross_entropy(logits, targets.to(dtype), None if class_weights is None else class_weights.to(dtype))
    loss.backward()
    grads = {k: v.grad.detach() for k, v in params.items()}
    return logits.detach(), loss.detach(), grads, sd2


def synthetic_bag(n_tiles: int, dim_input: int, seed: int, batch: int = 1, grid: int = 100,
                  tile_um: float = 256.0, signal: bool = False):
    """Seeded bag per SURVEY.md 8d: feats ~ N(0,1) rounded to fp16, coords = random cells of a
    grid x grid lattice times ``tile_um``."""
    g = torch.Generator().manual_seed(seed)
    feats = torch.randn(batch, n_tiles, dim_input, generator=g).half().float()
    if signal:
        k = max(1, n_tiles // 20)
        feats[:, :k] += 0.5
    cells = torch.stack([torch.randperm(grid * grid, generator=g)[:n_tiles] if n_tiles <= grid * grid
                         else torch.randint(0, grid * grid, (n_tiles,), generator=g) for _ in range(batch)])
    coords = torch.stack([(cells % grid).float(), (cells // grid).float()], dim=-1) * tile_um
    return feats, coords
