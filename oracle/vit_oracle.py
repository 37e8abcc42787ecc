"""ORACLE (test infrastructure, not product code) -- CPU restatement of the tile-encoder ViT forward
that STAMP runs through ``timm`` in ``extract_`` (src/stamp/preprocessing/__init__.py:322-327),
in plain torch fp32/fp64.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
legs may import this module, and only as the checker / the timed CPU baseline.

The arithmetic lives in a third-party dependency that is NOT under /root/reference:
``timm==1.0.25`` (reference pins: pyproject.toml:46, uv.lock:3470-3471).  This file restates the
published algorithm of ``timm/models/vision_transformer.py`` @1.0.25 for the two configurations the
reference constructs:

  * UNI        src/stamp/preprocessing/extractor/uni.py:26-31     vit_large_patch16_224,
               init_values=1e-5 (LayerScale), dynamic_img_size=True, num_classes=0, class token pooling
  * Virchow2   src/stamp/preprocessing/extractor/virchow2.py:24-42 vit_huge_patch14_224, reg_tokens=4,
               mlp_ratio=5.3375, SwiGLUPacked + SiLU, wrapper returns token 0 of the output

timm pieces restated: ``PatchEmbed`` (Conv2d k=p, stride=p, bias), ``VisionTransformer._pos_embed``
(no_embed_class=False: concat [cls, reg, patches] THEN add pos_embed of length n_patches+prefix),
``Block`` (x += ls1(attn(norm1(x))); x += ls2(mlp(norm2(x)))), ``Attention`` (fused qkv with bias,
softmax(q k^T / sqrt(hd)) v, proj), ``LayerScale`` (elementwise gamma), ``Mlp`` (fc1, exact-erf GELU,
fc2), ``GluMlp``/``SwiGLUPacked`` (gate_last=False: silu(x1) * x2 with x1, x2 = fc1(x).chunk(2)),
final ``norm`` (LayerNorm eps 1e-6), ``global_pool='token'`` -> x[:, 0].
The transform is ``ToTensor`` + ``Normalize(IMAGENET_DEFAULT_MEAN, IMAGENET_DEFAULT_STD)`` on
224 x 224 tiles (resize / centre-crop are identities at that size).

PARITY UNPINNED: timm, the pretrained weights and the reference's only golden test at this boundary
(tests/test_feature_extractors.py:83-169, ctranspath only, needs network) are unavailable offline.
The restatement is cross-checked by weight mapping against two independent implementations:
torchvision's ViT (``tests/test_oracle_cpu.py::test_vit_oracle_matches_torchvision``) and Hugging Face's
``Dinov2WithRegistersModel`` for LayerScale / packed SwiGLU / register tokens
(``::test_vit_oracle_matches_hf_dinov2_with_registers``).  State-dict keys are timm's.
"""

from __future__ import annotations

import math
from dataclasses import dataclass

import torch
import torch.nn.functional as F
from torch import Tensor

IMAGENET_MEAN = (0.485, 0.456, 0.406)
IMAGENET_STD = (0.229, 0.224, 0.225)


@dataclass(frozen=True)
class VitConfig:
    name: str
    img: int = 224
    patch: int = 16
    dim: int = 1024
    depth: int = 24
    heads: int = 16
    mlp_hidden: int = 4096      # fc1 output width (packed x1|x2 for swiglu)
    mlp: str = "gelu"           # "gelu" | "swiglu"
    reg_tokens: int = 0
    ln_eps: float = 1e-6
    no_embed_class: bool = False  # timm: pos_embed covers the patch tokens only, prefix tokens get none
    mean: tuple = IMAGENET_MEAN
    std: tuple = IMAGENET_STD

    @property
    def n_patches(self) -> int:
        return (self.img // self.patch) ** 2

    @property
    def n_prefix(self) -> int:
        return 1 + self.reg_tokens

    @property
    def n_tokens(self) -> int:
        return self.n_patches + self.n_prefix

    @property
    def fc2_in(self) -> int:
        return self.mlp_hidden // 2 if self.mlp == "swiglu" else self.mlp_hidden

    def flops_per_tile(self) -> float:
        """Algorithmic FLOPs (2*MACs) of one tile forward, all tokens, all blocks."""
        T, D, hd = self.n_tokens, self.dim, self.dim // self.heads
        blk = 2 * T * D * 3 * D + 2 * 2 * self.heads * T * T * hd + 2 * T * D * D
        blk += 2 * T * D * self.mlp_hidden + 2 * T * self.fc2_in * D
        return self.depth * blk + 2 * self.n_patches * 3 * self.patch * self.patch * D


UNI = VitConfig("uni")  # ViT-L/16
VIRCHOW2 = VitConfig("virchow2", patch=14, dim=1280, depth=32, heads=16, mlp_hidden=6832,
                     mlp="swiglu", reg_tokens=4)  # ViT-H/14, int(1280 * 5.3375) = 6832
# SURVEY.md 8f row N4 -- same kernels, other configs (all [external]: timm / HF model cards):
# UNI2-h, src/stamp/preprocessing/extractor/uni2.py:18-32 (ViT-H/14, 8 register tokens, SwiGLUPacked,
# no_embed_class, int(1536 * 5.33334) = 8192)
UNI2 = VitConfig("uni2", patch=14, dim=1536, depth=24, heads=24, mlp_hidden=8192, mlp="swiglu", reg_tokens=8,
                 no_embed_class=True)
# H-optimus-0 / -1, h_optimus_0.py:14-28: timm vit_giant_patch14_reg4_dinov2 at 224 px, own mean / std
H_OPTIMUS = VitConfig("h_optimus_0", patch=14, dim=1536, depth=40, heads=24, mlp_hidden=8192, mlp="swiglu",
                      reg_tokens=4, no_embed_class=True, mean=(0.707223, 0.578729, 0.703617),
                      std=(0.211883, 0.230117, 0.177517))
# Prov-GigaPath tile encoder, src/stamp/preprocessing/extractor/gigapath.py:14-35: timm vit_giant_patch14_dinov2 with the
# hub config's patch_size 16 / img_size 224 (embed 1536, depth 40, 24 heads, SwiGLUPacked int(1536 * 5.33334) = 8192,
# class token with its own position row); its transform resamples the tile first (oracle/resize_oracle.py)
GIGAPATH = VitConfig("gigapath", patch=16, dim=1536, depth=40, heads=24, mlp_hidden=8192, mlp="swiglu")


def tiny_config(mlp: str = "gelu", reg_tokens: int = 0, patch: int = 16, depth: int = 2,
                no_embed_class: bool = False, mean=IMAGENET_MEAN, std=IMAGENET_STD) -> VitConfig:
    """A small architecture-complete ViT for CPU-speed tests (configs[0] plumbing)."""
    return VitConfig(f"tiny-{mlp}-r{reg_tokens}-p{patch}", patch=patch, dim=128, depth=depth, heads=2,
                     mlp_hidden=512 if mlp == "gelu" else 688, mlp=mlp, reg_tokens=reg_tokens,
                     no_embed_class=no_embed_class, mean=mean, std=std)


def make_weights(cfg: VitConfig, seed: int = 1234, dtype=torch.float32) -> dict[str, Tensor]:
    """Seeded synthetic weights: trunc_normal(0.02) linears, LayerScale gamma ~
    U[0.5, 1.5] so that every block contributes (the reference's init_values=1e-5 would hide a
    broken block under a 1e-3 tolerance), non-trivial LN affine and biases."""
    g = torch.Generator().manual_seed(seed)
    D = cfg.dim

    def tn(*shape, std=0.02):
        return torch.nn.init.trunc_normal_(torch.empty(*shape), std=std, a=-2 * std, b=2 * std, generator=g)

    def ln():
        return 1.0 + 0.1 * torch.randn(D, generator=g), 0.02 * torch.randn(D, generator=g)

    w: dict[str, Tensor] = {}
    w["cls_token"] = tn(1, 1, D)
    if cfg.reg_tokens:
        w["reg_token"] = tn(1, cfg.reg_tokens, D)
    w["pos_embed"] = 0.02 * torch.randn(1, cfg.n_patches if cfg.no_embed_class else cfg.n_tokens, D, generator=g)
    w["patch_embed.proj.weight"] = tn(D, 3, cfg.patch, cfg.patch)
    w["patch_embed.proj.bias"] = tn(D)
    for i in range(cfg.depth):
        p = f"blocks.{i}."
        w[p + "norm1.weight"], w[p + "norm1.bias"] = ln()
        w[p + "attn.qkv.weight"], w[p + "attn.qkv.bias"] = tn(3 * D, D), tn(3 * D)
        w[p + "attn.proj.weight"], w[p + "attn.proj.bias"] = tn(D, D), tn(D)
        w[p + "ls1.gamma"] = 0.5 + torch.rand(D, generator=g)
        w[p + "norm2.weight"], w[p + "norm2.bias"] = ln()
        w[p + "mlp.fc1.weight"], w[p + "mlp.fc1.bias"] = tn(cfg.mlp_hidden, D), tn(cfg.mlp_hidden)
        w[p + "mlp.fc2.weight"], w[p + "mlp.fc2.bias"] = tn(D, cfg.fc2_in), tn(D)
        w[p + "ls2.gamma"] = 0.5 + torch.rand(D, generator=g)
    w["norm.weight"], w["norm.bias"] = ln()
    return {k: v.to(dtype) for k, v in w.items()}


def transform_u8(tiles_u8: Tensor, dtype=torch.float32, mean=IMAGENET_MEAN, std=IMAGENET_STD) -> Tensor:
    """uint8 [B,H,W,3] -> normalised CHW float (ToTensor + Normalize; ImageNet constants for UNI / Virchow2,
    explicit ones for e.g. H-optimus, h_optimus_0.py:26-28)."""
    x = tiles_u8.permute(0, 3, 1, 2).to(dtype) / 255.0
    mean = torch.tensor(mean, dtype=dtype, device=tiles_u8.device).view(1, 3, 1, 1)
    std = torch.tensor(std, dtype=dtype, device=tiles_u8.device).view(1, 3, 1, 1)
    return (x - mean) / std


def block_forward(w: dict[str, Tensor], p: str, x: Tensor, cfg: VitConfig) -> Tensor:
    B, T, D = x.shape
    H, hd = cfg.heads, D // cfg.heads
    h = F.layer_norm(x, (D,), w[p + "norm1.weight"], w[p + "norm1.bias"], cfg.ln_eps)
    qkv = F.linear(h, w[p + "attn.qkv.weight"], w[p + "attn.qkv.bias"])
    q, k, v = qkv.reshape(B, T, 3, H, hd).permute(2, 0, 3, 1, 4)
    if x.is_cuda:   # the GPU baseline of bench.py: timm's fused_attn path (F.scaled_dot_product_attention)
        att = F.scaled_dot_product_attention(q, k, v)
    else:
        att = torch.softmax((q @ k.transpose(-2, -1)) * hd ** -0.5, dim=-1) @ v
    att = att.transpose(1, 2).reshape(B, T, D)
    att = F.linear(att, w[p + "attn.proj.weight"], w[p + "attn.proj.bias"])
    x = x + w[p + "ls1.gamma"] * att
    h = F.layer_norm(x, (D,), w[p + "norm2.weight"], w[p + "norm2.bias"], cfg.ln_eps)
    h = F.linear(h, w[p + "mlp.fc1.weight"], w[p + "mlp.fc1.bias"])
    if cfg.mlp == "swiglu":
        x1, x2 = h.chunk(2, dim=-1)
        h = F.silu(x1) * x2
    else:
        h = F.gelu(h)
    h = F.linear(h, w[p + "mlp.fc2.weight"], w[p + "mlp.fc2.bias"])
    return x + w[p + "ls2.gamma"] * h


def forward_tokens(w: dict[str, Tensor], cfg: VitConfig, x: Tensor) -> Tensor:
    """Normalised CHW input [B,3,H,W] -> all tokens after the final norm [B,T,D]."""
    B = x.shape[0]
    x = F.conv2d(x, w["patch_embed.proj.weight"], w["patch_embed.proj.bias"], stride=cfg.patch)
    x = x.flatten(2).transpose(1, 2)  # [B, n_patches, D], row-major over (py, px)
    prefix = [w["cls_token"].expand(B, -1, -1)]
    if cfg.reg_tokens:
        prefix.append(w["reg_token"].expand(B, -1, -1))
    if cfg.no_embed_class:
        # timm VisionTransformer._pos_embed, no_embed_class=True (UNI2-h, H-optimus: uni2.py:27,
        # timm vit_giant_patch14_reg4_dinov2): position table added to the patch tokens BEFORE the concat
        x = torch.cat(prefix + [x + w["pos_embed"]], dim=1)
    else:
        x = torch.cat(prefix + [x], dim=1) + w["pos_embed"]
    for i in range(cfg.depth):
        x = block_forward(w, f"blocks.{i}.", x, cfg)
    return F.layer_norm(x, (cfg.dim,), w["norm.weight"], w["norm.bias"], cfg.ln_eps)


def forward(w: dict[str, Tensor], cfg: VitConfig, tiles_u8: Tensor) -> Tensor:
    """uint8 HWC tiles -> class-token features [B, D] (what ``extract_`` stores, before .half())."""
    dt = w["norm.weight"].dtype
    return forward_tokens(w, cfg, transform_u8(tiles_u8, dt, cfg.mean, cfg.std))[:, 0]


def synthetic_tiles(n: int, seed: int, img: int = 224) -> Tensor:
    """H&E-like uint8 tiles: per-pixel stain concentrations c_H, c_E ~ Gamma(2, 0.5)
    smoothed with a 5x5 box filter, OD = HERef . c, I = clip(240 exp(-OD) + N(0, 2))."""
    g = torch.Generator().manual_seed(seed)
    he_ref = torch.tensor([[0.5626, 0.2159], [0.7201, 0.8012], [0.4062, 0.5581]])
    # Gamma(2, scale 0.5) = sum of two Exp(scale 0.5)
    c = -0.5 * (torch.log(torch.rand(n, 2, img, img, generator=g).clamp_min(1e-12))
                + torch.log(torch.rand(n, 2, img, img, generator=g).clamp_min(1e-12)))
    c = F.avg_pool2d(F.pad(c, (2, 2, 2, 2), mode="reflect"), 5, stride=1)
    od = torch.einsum("ck,nkhw->nchw", he_ref, c)
    img_f = 240.0 * torch.exp(-od) + 2.0 * torch.randn(n, 3, img, img, generator=g)
    return img_f.clamp(0, 255).round().to(torch.uint8).permute(0, 2, 3, 1).contiguous()
