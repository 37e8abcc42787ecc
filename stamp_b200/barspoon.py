"""Barspoon encoder-decoder aggregator (multi-target MIL) on the B200 primitives -- inference.

Mirrors ``EncDecTransformer`` (src/stamp/modeling/models/barspoon.py:24-170; selected by
``ModelName.BARSPOON``, src/stamp/modeling/registry.py:60-66): same constructor, same sub-modules and therefore
the same state-dict keys (``projector.0``, ``transformer_encoder.layers.N.*``, ``class_tokens.<label>``,
``transformer_decoder.layers.N.*``, ``heads.<label>``), ``forward(tile_tokens [B, S, F], tile_positions [B, S, 2])
-> {label: logits [B, n_out]}``.  The arithmetic runs through the C-ABI primitives of this package:

  projector Linear + ReLU                       stamp_gemm_tn (fp16 operands, fp32 residual stream out)
  sinusoidal position code (:146-157)           element-wise on the device, added to the stream
  encoder layers (pre-norm, ReLU feed-forward)  stamp_layernorm -> stamp_gemm_tn (packed q|k|v) -> stamp_attention_fwd
                                                -> stamp_gemm_tn (+residual) -> stamp_layernorm -> stamp_gemm_tn (ReLU)
                                                -> stamp_gemm_tn (+residual)
  decoder layers over the class tokens          the same primitives; cross-attention = the class-token queries in the
                                                first rows of a packed q|k|v buffer whose k|v come from the encoder output
  one Linear head per target                    fp32 on the handful of class tokens

Inference only (``torch.no_grad``; the module refuses to record a graph); dropout is the identity in eval mode, as in
``nn.TransformerEncoderLayer``.  Head dimension 64 or 32 (``d_model / heads``), widths multiples of 8.
"""

from __future__ import annotations

import re

import torch
from torch import Tensor, nn

from . import ops


def sanitize(x: str) -> str:
    return re.sub(r"[^A-Za-z0-9_]", "_", x)


class EncDecTransformer(nn.Module):
    def __init__(self, d_features: int, target_n_outs: dict[str, int], *, d_model: int = 512,
                 num_encoder_heads: int = 8, num_decoder_heads: int = 8, num_encoder_layers: int = 2,
                 num_decoder_layers: int = 2, dim_feedforward: int = 2048, positional_encoding: bool = True) -> None:
        super().__init__()
        self.projector = nn.Sequential(nn.Linear(d_features, d_model), nn.ReLU())
        enc = nn.TransformerEncoderLayer(d_model=d_model, nhead=num_encoder_heads, dim_feedforward=dim_feedforward,
                                         batch_first=True, norm_first=True)
        self.transformer_encoder = nn.TransformerEncoder(enc, num_layers=num_encoder_layers, enable_nested_tensor=False)
        self.target_labels = target_n_outs.keys()
        self.class_tokens = nn.ParameterDict({sanitize(t): torch.rand(d_model) for t in target_n_outs})
        dec = nn.TransformerDecoderLayer(d_model=d_model, nhead=num_decoder_heads, dim_feedforward=dim_feedforward,
                                         batch_first=True, norm_first=True)
        self.transformer_decoder = nn.TransformerDecoder(dec, num_layers=num_decoder_layers)
        self.heads = nn.ModuleDict({sanitize(t): nn.Linear(d_model, n) for t, n in target_n_outs.items()})
        self.positional_encoding = positional_encoding
        self._cfg = dict(d_features=d_features, d_model=d_model, he=num_encoder_heads, hd=num_decoder_heads)
        for h in (num_encoder_heads, num_decoder_heads):
            if d_model % h or d_model // h not in (32, 64) or d_features % 8 or d_model % 8 or dim_feedforward % 8:
                raise ValueError("unsupported barspoon configuration for the sm_100a kernels (head dimension 32 / 64, "
                                 "widths in multiples of 8)")
        self._half: dict[int, tuple[tuple, Tensor]] = {}

    def _w16(self, p: Tensor) -> Tensor:
        """fp16 copy of a weight matrix, refreshed when the parameter changes."""
        try:
            key = (p._version, p.data_ptr())
        except RuntimeError:            # parameters created under inference_mode do not track versions
            key = (-1, p.data_ptr())
        hit = self._half.get(id(p))
        if hit is None or hit[0] != key or hit[1].device != p.device:
            hit = (key, p.detach().half().contiguous())
            self._half[id(p)] = hit
        return hit[1]

    @staticmethod
    def position_code(pos: Tensor, d_model: int) -> Tensor:
        """barspoon.py:146-157: [S, 2] positions -> [S, d_model] (sines of both axes, then cosines)."""
        x = pos.unsqueeze(-1) / 100_000 ** (torch.arange(d_model // 4, device=pos.device).type_as(pos) / d_model)
        return torch.cat([torch.sin(x).flatten(start_dim=-2), torch.cos(x).flatten(start_dim=-2)], dim=-1)

    def _self_attention_block(self, x: Tensor, norm: nn.LayerNorm, mha: nn.MultiheadAttention, heads: int) -> None:
        """x += out_proj(attention(in_proj(norm(x))))   (x fp32 [S, d], in place)."""
        S, d = x.shape
        y = ops.layernorm(x, norm.weight, norm.bias, norm.eps, torch.float16)
        qkv = torch.empty((S, 3 * d), dtype=torch.float16, device=x.device)
        ops.gemm_tn(y, self._w16(mha.in_proj_weight), out=qkv, bias=mha.in_proj_bias)
        att = ops.attention(qkv.view(1, S, 3 * d), heads).view(S, d)
        ops.gemm_tn(att, self._w16(mha.out_proj.weight), out=x, bias=mha.out_proj.bias, store=ops.ST_RESID32)

    def _feed_forward_block(self, x: Tensor, norm: nn.LayerNorm, lin1: nn.Linear, lin2: nn.Linear) -> None:
        y = ops.layernorm(x, norm.weight, norm.bias, norm.eps, torch.float16)
        h = torch.empty((x.shape[0], lin1.out_features), dtype=torch.float16, device=x.device)
        ops.gemm_tn(y, self._w16(lin1.weight), out=h, bias=lin1.bias, act=ops.ACT_RELU)
        ops.gemm_tn(h, self._w16(lin2.weight), out=x, bias=lin2.bias, store=ops.ST_RESID32)

    def _cross_attention_block(self, t: Tensor, memory16: Tensor, norm: nn.LayerNorm, mha: nn.MultiheadAttention,
                               heads: int) -> None:
        """t += out_proj(attention(q = Wq norm(t), k|v = Wk|Wv memory))   (t fp32 [T, d]; memory16 fp16 [S, d])."""
        T, d = t.shape
        S = memory16.shape[0]
        rows = max(S, T)
        buf = torch.zeros((rows, 3 * d), dtype=torch.float16, device=t.device)   # q | k | v; unused query rows stay 0
        y = ops.layernorm(t, norm.weight, norm.bias, norm.eps, torch.float16)
        w16 = self._w16(mha.in_proj_weight)
        ops.gemm_tn(y, w16[:d], out=buf[:T, :d], bias=mha.in_proj_bias[:d].contiguous())
        ops.gemm_tn(memory16, w16[d:], out=buf[:S, d:], bias=mha.in_proj_bias[d:].contiguous())
        if rows != S:
            raise ValueError("more class tokens than tiles: cross-attention over a shorter memory is not supported")
        att = ops.attention(buf.view(1, S, 3 * d), heads).view(S, d)[:T].contiguous()
        ops.gemm_tn(att, self._w16(mha.out_proj.weight), out=t, bias=mha.out_proj.bias, store=ops.ST_RESID32)

    def forward(self, tile_tokens: Tensor, tile_positions: Tensor) -> dict[str, Tensor]:
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            raise NotImplementedError("stamp_b200 EncDecTransformer is inference-only: call it under torch.no_grad() "
                                      "/ inference_mode() (training runs through the reference module)")
        if not tile_tokens.is_cuda or not self.projector[0].weight.is_cuda:
            raise RuntimeError("stamp_b200 EncDecTransformer runs on a CUDA device only (no CPU fallback)")
        if self.training:
            raise RuntimeError("call .eval() first: dropout of the training mode is not implemented")
        B, S, F = tile_tokens.shape
        d = self._cfg["d_model"]
        if F != self._cfg["d_features"] or tuple(tile_positions.shape) != (B, S, 2):
            raise TypeError(f"expected tile_tokens [B,S,{self._cfg['d_features']}] and tile_positions [B,S,2]")
        labels = list(self.target_labels)
        cls0 = torch.stack([self.class_tokens[sanitize(t)] for t in labels]).detach().float()
        out: dict[str, list[Tensor]] = {t: [] for t in labels}
        proj = self.projector[0]
        for b in range(B):
            feats16 = tile_tokens[b].detach().half().contiguous()
            x = torch.empty((S, d), dtype=torch.float32, device=feats16.device)
            ops.gemm_tn(feats16, self._w16(proj.weight), out=x, bias=proj.bias, act=ops.ACT_RELU, store=ops.ST_32)
            if self.positional_encoding:
                x += self.position_code(tile_positions[b].detach().float(), d)
            for layer in self.transformer_encoder.layers:
                self._self_attention_block(x, layer.norm1, layer.self_attn, self._cfg["he"])
                self._feed_forward_block(x, layer.norm2, layer.linear1, layer.linear2)
            memory16 = x.half()
            t = cls0.clone()
            for layer in self.transformer_decoder.layers:
                self._self_attention_block(t, layer.norm1, layer.self_attn, self._cfg["hd"])
                self._cross_attention_block(t, memory16, layer.norm2, layer.multihead_attn, self._cfg["hd"])
                self._feed_forward_block(t, layer.norm3, layer.linear1, layer.linear2)
            for i, lab in enumerate(labels):
                head = self.heads[sanitize(lab)]
                out[lab].append(torch.nn.functional.linear(t[i], head.weight.float(), head.bias.float()))
        return {lab: torch.stack(v) for lab, v in out.items()}
