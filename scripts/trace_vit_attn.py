"""Development aid: per-phase cycle counts of the tcgen05 ViT attention kernel."""
import ctypes as C
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch

from stamp_b200 import _lib, ops

dev = torch.device("cuda:0")
B, S, H = int(sys.argv[1]) if len(sys.argv) > 1 else 192, 197, 16
qkv = torch.randn(B, S, 3 * H * 64, device=dev, dtype=torch.float16)
lib = _lib.load()
lib.stamp_b200_debug_attention_trace.argtypes = [C.c_void_p]
for _ in range(3):
    ops.attention(qkv, H)
n_cta = 2 * B * H
trace = torch.zeros(n_cta, 6, dtype=torch.int64, device=dev)
lib.stamp_b200_debug_attention_trace(trace.data_ptr())
ops.attention(qkv, H)
torch.cuda.synchronize()
lib.stamp_b200_debug_attention_trace(None)
t = trace.cpu().double()
names = ["start->S ready (TMA load + QK^T)", "softmax (max, exp, P store)", "wait O (P.V MMA)", "epilogue (TMEM->HBM)"]
for i, n in enumerate(names):
    print(f"{n:36s} mean {t[:, i].mean():9.0f}  p10 {t[:, i].quantile(0.1):9.0f}  p90 {t[:, i].quantile(0.9):9.0f} cycles")
span = (t[:, 4].max() - t[:, 4].min() + t[:, :4].sum(1).max())
print(f"  of which pass 1 (row max): mean {t[:, 5].mean():.0f} cycles")
print("per-CTA total mean", t[:, :4].sum(1).mean().item(), " kernel span ~", span.item(), "cycles; CTAs", n_cta)
