"""Quick MIL aggregator probe (development aid): slides/s for batch-1 whole-bag forwards."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch

from stamp_b200.mil import VisionTransformer

dev = torch.device("cuda:0")
n_tiles = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
alibi = "noalibi" not in sys.argv
mil = VisionTransformer(dim_output=2, dim_input=1024, dim_model=512, n_layers=2, n_heads=8, dim_feedforward=512,
                        dropout=0.25, use_alibi=alibi).to(dev).eval()
bags = torch.randn(8, n_tiles, 1024, device=dev).half().float()
coords = torch.randint(0, 100, (8, n_tiles, 2), device=dev).float() * 256.0
with torch.inference_mode():
    for i in range(8):
        mil(bags[i:i + 1], coords=coords[i:i + 1], mask=None)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for rep in range(4):
        for i in range(8):
            mil(bags[i:i + 1], coords=coords[i:i + 1], mask=None)
    e1.record()
    torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 32
print(f"MIL alibi={alibi} N={n_tiles}: {ms:.3f} ms/bag, {1e3 / ms:.0f} slides/s")
