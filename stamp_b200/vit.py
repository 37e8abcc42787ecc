"""Host side of the tile-encoder ViT: weight packing (timm state-dict layout -> the C ABI's
``StampVitWeights`` / ``StampVitBlock``) and an ``nn.Module`` whose ``forward`` is one call into
``stamp_vit_forward``.

reference: the module stands where ``extractor.model`` stands in
src/stamp/preprocessing/__init__.py:243,322-327 -- ``model.to(device).eval()`` then
``model(batch) -> Tensor[B, D]`` under ``inference_mode``; UNI / Virchow2 construction in
src/stamp/preprocessing/extractor/uni.py:26-31 and virchow2.py:24-42.
"""

from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import torch
from torch import Tensor, nn

from . import _lib

IMAGENET_MEAN = (0.485, 0.456, 0.406)
IMAGENET_STD = (0.229, 0.224, 0.225)


@dataclass(frozen=True)
class VitArch:
    """Architecture of a timm-style plain ViT tile encoder."""

    name: str
    img: int = 224
    patch: int = 16
    dim: int = 1024
    depth: int = 24
    heads: int = 16
    mlp_hidden: int = 4096
    mlp: str = "gelu"  # "gelu" | "swiglu" (timm SwiGLUPacked)
    reg_tokens: int = 0
    ln_eps: float = 1e-6
    mean: tuple = IMAGENET_MEAN
    std: tuple = IMAGENET_STD
    no_embed_class: bool = False  # timm: pos_embed has n_patches rows, prefix tokens get no position
    pool: str = "cls"    # "cls": class token; "cls_mean": class token | mean of the other tokens (virchow_full.py:24-35)
    pre_resize: int = 0  # > 0: transforms.Resize(pre_resize, BICUBIC) + CenterCrop(img) ahead of the model (GigaPath)

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
    def out_dim(self) -> int:
        return 2 * self.dim if self.pool == "cls_mean" else self.dim

    @property
    def kpad(self) -> int:
        return (3 * self.patch * self.patch + 7) // 8 * 8

    @property
    def fc2_in(self) -> int:
        return self.mlp_hidden // 2 if self.mlp == "swiglu" else self.mlp_hidden

    def flops_per_tile(self) -> float:
        """Algorithmic FLOPs (2 x MACs) of one tile forward (123.1 GFLOP for ViT-L/16)."""
        T, D, hd = self.n_tokens, self.dim, self.dim // self.heads
        blk = 2 * T * D * 3 * D + 2 * 2 * self.heads * T * T * hd + 2 * T * D * D
        blk += 2 * T * D * self.mlp_hidden + 2 * T * self.fc2_in * D
        return self.depth * blk + 2 * self.n_patches * 3 * self.patch * self.patch * D


# uni.py:26-31 -> timm vit_large_patch16_224; virchow2.py:37-42 -> vit_huge_patch14_224 with
# reg_tokens=4, mlp_ratio=5.3375 (int(1280*5.3375) = 6832), SwiGLUPacked, SiLU
UNI_ARCH = VitArch("uni")
VIRCHOW2_ARCH = VitArch("virchow2", patch=14, dim=1280, depth=32, heads=16, mlp_hidden=6832,
                        mlp="swiglu", reg_tokens=4)
# uni2.py:18-32 -> ViT-H/14, embed 1536, depth 24, 24 heads, 8 register tokens, SwiGLUPacked
# (int(1536 * 2.66667 * 2) = 8192), no_embed_class
UNI2_ARCH = VitArch("uni2", patch=14, dim=1536, depth=24, heads=24, mlp_hidden=8192, mlp="swiglu",
                    reg_tokens=8, no_embed_class=True)
# h_optimus_0.py:14-28 / h_optimus_1.py: timm vit_giant_patch14_reg4_dinov2 (embed 1536, depth 40, 24 heads,
# 4 register tokens, SwiGLUPacked, no_embed_class) at 224 px with the model card's mean / std
H_OPTIMUS_ARCH = VitArch("h_optimus", patch=14, dim=1536, depth=40, heads=24, mlp_hidden=8192, mlp="swiglu",
                         reg_tokens=4, no_embed_class=True, mean=(0.707223, 0.578729, 0.703617),
                         std=(0.211883, 0.230117, 0.177517))
# virchow.py:33-57 / virchow_full.py:38-62: Virchow (v1), ViT-H/14 without register tokens, SwiGLUPacked; the "full"
# variant returns class token | mean patch token (2560 values)
VIRCHOW_ARCH = VitArch("virchow", patch=14, dim=1280, depth=32, heads=16, mlp_hidden=6832, mlp="swiglu")
VIRCHOW_FULL_ARCH = VitArch("virchow_full", patch=14, dim=1280, depth=32, heads=16, mlp_hidden=6832, mlp="swiglu",
                            pool="cls_mean")
# dinobloom.py:30-78: DinoBloom-S = facebookresearch/dinov2 ViT-S/14 (embed 384, depth 12, 6 heads, LayerScale, GELU MLP
# 1536) with a 257-row position table (16 x 16 patches of a 224 px tile + class token: no resampling); returns the
# normalised class token.  The DINOv2 state-dict keys are timm's (plus an unused mask_token).
DINOBLOOM_ARCH = VitArch("dinobloom", patch=14, dim=384, depth=12, heads=6, mlp_hidden=1536)
# gigapath.py:14-35: timm vit_giant_patch14_dinov2 with patch 16 at 224 px (embed 1536, depth 40, 24 heads, SwiGLUPacked
# 8192, class token inside the position table); the transform's Resize(256, BICUBIC) + CenterCrop(224) runs on the GPU
GIGAPATH_ARCH = VitArch("gigapath", patch=16, dim=1536, depth=40, heads=24, mlp_hidden=8192, mlp="swiglu",
                        pre_resize=256)


class StampVitConfig(C.Structure):
    _fields_ = [("img", C.c_int), ("patch", C.c_int), ("dim", C.c_int), ("depth", C.c_int),
                ("heads", C.c_int), ("mlp_hidden", C.c_int), ("mlp_kind", C.c_int),
                ("reg_tokens", C.c_int), ("kpad", C.c_int), ("ln_eps", C.c_float),
                ("mean", C.c_float * 3), ("std", C.c_float * 3), ("pool", C.c_int)]


class StampVitWeights(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("patch_w", "patch_b", "prefix", "pos", "norm_w", "norm_b")]


class StampVitBlock(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("ln1_w", "ln1_b", "qkv_w", "qkv_b", "proj_w", "proj_b", "ls1",
                                          "ln2_w", "ln2_b", "fc1_w", "fc1_b", "fc2_w", "fc2_b", "ls2")]


def _bind() -> C.CDLL:
    lib = _lib.load()
    if not getattr(lib, "_vit_bound", False):
        lib.stamp_vit_workspace_bytes.restype = C.c_size_t
        lib.stamp_vit_workspace_bytes.argtypes = [C.POINTER(StampVitConfig), C.c_int]
        lib.stamp_vit_forward.restype = C.c_int
        lib.stamp_vit_forward.argtypes = [C.POINTER(StampVitConfig), C.POINTER(StampVitWeights),
                                          C.POINTER(StampVitBlock), C.c_void_p, C.c_void_p, C.c_int,
                                          C.c_void_p, C.c_size_t, C.c_void_p]
        lib._vit_bound = True
    return lib


def random_state_dict(arch: VitArch, seed: int = 1234) -> dict[str, Tensor]:
    """Random-init weights in timm's state-dict layout (for benchmarks: no checkpoints offline)."""
    g = torch.Generator().manual_seed(seed)
    D = arch.dim
    sd: dict[str, Tensor] = {
        "cls_token": 0.02 * torch.randn(1, 1, D, generator=g),
        "pos_embed": 0.02 * torch.randn(1, arch.n_patches if arch.no_embed_class else arch.n_tokens, D, generator=g),
        "patch_embed.proj.weight": 0.02 * torch.randn(D, 3, arch.patch, arch.patch, generator=g),
        "patch_embed.proj.bias": torch.zeros(D),
        "norm.weight": torch.ones(D), "norm.bias": torch.zeros(D),
    }
    if arch.reg_tokens:
        sd["reg_token"] = 0.02 * torch.randn(1, arch.reg_tokens, D, generator=g)
    for i in range(arch.depth):
        p = f"blocks.{i}."
        for n in ("norm1", "norm2"):
            sd[p + n + ".weight"], sd[p + n + ".bias"] = torch.ones(D), torch.zeros(D)
        sd[p + "attn.qkv.weight"] = 0.02 * torch.randn(3 * D, D, generator=g)
        sd[p + "attn.qkv.bias"] = torch.zeros(3 * D)
        sd[p + "attn.proj.weight"] = 0.02 * torch.randn(D, D, generator=g)
        sd[p + "attn.proj.bias"] = torch.zeros(D)
        sd[p + "mlp.fc1.weight"] = 0.02 * torch.randn(arch.mlp_hidden, D, generator=g)
        sd[p + "mlp.fc1.bias"] = torch.zeros(arch.mlp_hidden)
        sd[p + "mlp.fc2.weight"] = 0.02 * torch.randn(D, arch.fc2_in, generator=g)
        sd[p + "mlp.fc2.bias"] = torch.zeros(D)
        sd[p + "ls1.gamma"] = 0.5 + torch.rand(D, generator=g)
        sd[p + "ls2.gamma"] = 0.5 + torch.rand(D, generator=g)
    return sd


class TileEncoder(nn.Module):
    """ViT tile encoder running entirely in ``libstamp_b200.so``.

    ``forward(tiles)``: uint8 ``[B, H, W, 3]`` (or ``[B, 3, H, W]``) tiles on the module's CUDA
    device -> fp16 ``[B, dim]`` class-token features.  Weights are packed once from a timm
    state dict (fp16 GEMM operands, fp32 norms / biases / LayerScale / position table).
    """

    def __init__(self, arch: VitArch, state_dict: dict[str, Tensor], max_batch: int = 256) -> None:
        super().__init__()
        self.arch = arch
        self.max_batch = max_batch
        sd = {k.removeprefix("model."): v.detach().float() for k, v in state_dict.items()}
        D = arch.dim
        k = 3 * arch.patch * arch.patch
        pw = torch.zeros(D, arch.kpad)
        pw[:, :k] = sd["patch_embed.proj.weight"].reshape(D, k)
        prefix = [sd["cls_token"].reshape(1, D)]
        if arch.reg_tokens:
            prefix.append(sd["reg_token"].reshape(arch.reg_tokens, D))
        pos = sd["pos_embed"].reshape(-1, D)
        if pos.shape[0] == arch.n_patches and (arch.no_embed_class or arch.n_prefix == 0):
            # timm no_embed_class=True: the table covers the patch tokens only (UNI2-h, H-optimus)
            pos = torch.cat([torch.zeros(arch.n_prefix, D), pos], 0)
        elif pos.shape[0] != arch.n_tokens or arch.no_embed_class:
            raise ValueError(f"pos_embed has {pos.shape[0]} rows, expected "
                             f"{arch.n_patches if arch.no_embed_class else arch.n_tokens} "
                             "(dynamic resampling of the position table is not supported)")
        self._names: list[str] = []

        def reg(name: str, t: Tensor, half: bool = False) -> None:
            self.register_buffer(name, t.half().contiguous() if half else t.float().contiguous(),
                                 persistent=False)
            self._names.append(name)

        reg("patch_w", pw, half=True)
        reg("patch_b", sd["patch_embed.proj.bias"])
        reg("prefix", torch.cat(prefix, 0) + pos[: arch.n_prefix])
        reg("pos", pos[arch.n_prefix:])
        reg("norm_w", sd["norm.weight"])
        reg("norm_b", sd["norm.bias"])
        for i in range(arch.depth):
            p = f"blocks.{i}."
            reg(f"b{i}_ln1_w", sd[p + "norm1.weight"]); reg(f"b{i}_ln1_b", sd[p + "norm1.bias"])
            reg(f"b{i}_qkv_w", sd[p + "attn.qkv.weight"], half=True); reg(f"b{i}_qkv_b", sd[p + "attn.qkv.bias"])
            reg(f"b{i}_proj_w", sd[p + "attn.proj.weight"], half=True); reg(f"b{i}_proj_b", sd[p + "attn.proj.bias"])
            reg(f"b{i}_ls1", sd.get(p + "ls1.gamma", torch.ones(D)))
            reg(f"b{i}_ln2_w", sd[p + "norm2.weight"]); reg(f"b{i}_ln2_b", sd[p + "norm2.bias"])
            w1, b1 = sd[p + "mlp.fc1.weight"], sd[p + "mlp.fc1.bias"]
            if arch.mlp == "swiglu":
                # SwiGLUPacked: x1, x2 = fc1(x).chunk(2); interleave rows so (x1_j, x2_j) are
                # adjacent output columns of one GEMM tile
                h = arch.mlp_hidden // 2
                w1 = torch.stack([w1[:h], w1[h:]], dim=1).reshape(arch.mlp_hidden, D)
                b1 = torch.stack([b1[:h], b1[h:]], dim=1).reshape(arch.mlp_hidden)
            reg(f"b{i}_fc1_w", w1, half=True); reg(f"b{i}_fc1_b", b1)
            reg(f"b{i}_fc2_w", sd[p + "mlp.fc2.weight"], half=True); reg(f"b{i}_fc2_b", sd[p + "mlp.fc2.bias"])
            reg(f"b{i}_ls2", sd.get(p + "ls2.gamma", torch.ones(D)))
        self._structs = None
        self._workspace: Tensor | None = None

    # pointers are only valid for the device the buffers currently live on
    def _apply(self, fn, *a, **kw):
        self._structs = None
        self._workspace = None
        return super()._apply(fn, *a, **kw)

    def _build_structs(self):
        a = self.arch
        cfg = StampVitConfig(a.img, a.patch, a.dim, a.depth, a.heads, a.mlp_hidden,
                             1 if a.mlp == "swiglu" else 0, a.reg_tokens, a.kpad, a.ln_eps,
                             (C.c_float * 3)(*a.mean), (C.c_float * 3)(*a.std), {"cls": 0, "cls_mean": 1}[a.pool])
        g = lambda n: getattr(self, n).data_ptr()
        w = StampVitWeights(g("patch_w"), g("patch_b"), g("prefix"), g("pos"), g("norm_w"), g("norm_b"))
        blocks = (StampVitBlock * a.depth)()
        for i in range(a.depth):
            blocks[i] = StampVitBlock(*[g(f"b{i}_{n}") for n in (
                "ln1_w", "ln1_b", "qkv_w", "qkv_b", "proj_w", "proj_b", "ls1",
                "ln2_w", "ln2_b", "fc1_w", "fc1_b", "fc2_w", "fc2_b", "ls2")])
        self._structs = (cfg, w, blocks)

    def launches_per_batch(self) -> int:
        return 3 + 7 * self.arch.depth + 1 + (1 if self.arch.pre_resize else 0) + (1 if self.arch.pool == "cls_mean" else 0)

    @torch.no_grad()
    def forward(self, tiles: Tensor) -> Tensor:
        if not tiles.is_cuda or not self.patch_w.is_cuda:
            raise RuntimeError("TileEncoder runs on a CUDA device only (no CPU fallback); "
                               "call .to('cuda') and pass CUDA tiles")
        if tiles.dtype != torch.uint8:
            raise TypeError("TileEncoder expects uint8 tiles (use the extractor's transform)")
        if tiles.dim() == 4 and tiles.shape[1] == 3 and tiles.shape[-1] != 3:
            tiles = tiles.permute(0, 2, 3, 1)
        tiles = tiles.contiguous()
        a = self.arch
        if a.pre_resize:
            from .resize import resize_center_crop

            tiles = resize_center_crop(tiles, a.pre_resize, a.img)
        if tiles.dim() != 4 or tiles.shape[1:] != (a.img, a.img, 3):
            raise ValueError(f"expected tiles [B,{a.img},{a.img},3], got {tuple(tiles.shape)}")
        lib = _bind()
        if self._structs is None:
            self._build_structs()
        cfg, w, blocks = self._structs
        n = tiles.shape[0]
        out = torch.empty((n, a.out_dim), dtype=torch.float16, device=tiles.device)
        stream = torch.cuda.current_stream().cuda_stream
        for s in range(0, n, self.max_batch):
            b = min(self.max_batch, n - s)
            need = lib.stamp_vit_workspace_bytes(C.byref(cfg), b)
            if need == 0:
                raise ValueError("invalid ViT configuration for the sm_100a kernels")
            if self._workspace is None or self._workspace.numel() < need:
                self._workspace = torch.empty(need, dtype=torch.uint8, device=tiles.device)
            code = lib.stamp_vit_forward(C.byref(cfg), C.byref(w), blocks, tiles[s:s + b].data_ptr(),
                                         out[s:s + b].data_ptr(), b, self._workspace.data_ptr(),
                                         self._workspace.numel(), stream)
            _lib.check(code, "stamp_vit_forward")
        return out
