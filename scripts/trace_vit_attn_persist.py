"""Development aid: mean per-item phase cycles of the persistent ViT attention kernel."""
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
trace = torch.zeros(148, 6, dtype=torch.int64, device=dev)
lib.stamp_b200_debug_attention_trace(trace.data_ptr())
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
ops.attention(qkv, H)
e1.record()
torch.cuda.synchronize()
lib.stamp_b200_debug_attention_trace(None)
t = trace.cpu().double()
for i, n in enumerate(["wait S(i) in TMEM", "S->regs, max, exp (4 bar.sync)", "wait P free (PV(i-1))", "epilogue O(i-1)", "P store + fence"]):
    print(f"{n:34s} mean {t[:, i].mean():8.0f} cycles/item")
print("sum", t[:, :5].sum(1).mean().item(), "cycles/item; kernel", e0.elapsed_time(e1) * 1e3, "us; items/SM", 2 * B * H / 148)
