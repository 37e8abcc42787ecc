"""Development aid: which mode / rows of the short-sequence attention kernels differ from the fp32 reference."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch

from stamp_b200 import _lib, ops

dev = torch.device("cuda:0")


def ref_attn(qkv, H):
    B, S, _ = qkv.shape
    q, k, v = qkv.float().view(B, S, 3, H, 64).permute(2, 0, 3, 1, 4)
    a = torch.softmax(q @ k.transpose(-1, -2) / 8.0, -1) @ v
    return a.permute(0, 2, 1, 3).reshape(B, S, H * 64)


for B, S, H in [(2, 129, 2), (2, 128, 2), (5, 197, 16), (40, 197, 16), (2, 65, 1), (2, 193, 2), (2, 192, 2), (2, 200, 2), (192, 197, 16)]:
    g = torch.Generator(device="cpu").manual_seed(S * 3 + H)
    qkv = torch.randn(B, S, 3 * H * 64, generator=g).to(dev, torch.float16)
    ref = ref_attn(qkv, H)
    for mode in (1, 9, 65):
        _lib.load().stamp_b200_attention_tc_enable(mode)
        out = ops.attention(qkv, H).float()
        err = (out - ref).abs().amax(dim=(0, 2))       # per token row
        bad = (err > 5e-3).nonzero().flatten().tolist()
        perb = (out - ref).abs().amax(dim=(1, 2))
        print(f"B={B} S={S} H={H} mode={mode}: max err {float(err.max()):.3e} bad rows {bad[:12]}{'...' if len(bad) > 12 else ''} "
              f"n_bad={len(bad)} bad images {(perb > 5e-3).nonzero().flatten().tolist()[:10]} finite={bool(torch.isfinite(out).all())}")
    _lib.load().stamp_b200_attention_tc_enable(1)
