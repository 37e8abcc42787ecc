"""Small run of the long-bag tcgen05 kernels, the tcgen05 weight-gradient kernel and the texture filter for
compute-sanitizer (development aid):  compute-sanitizer --tool memcheck python scripts/sanitize_probe.py"""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch

from stamp_b200 import train as T
from stamp_b200.mil import VisionTransformer
from stamp_b200.tiling import canny_edge_counts

dev = torch.device("cuda:0")
torch.manual_seed(0)
m = VisionTransformer(dim_output=2, dim_input=64, dim_model=128, n_layers=1, n_heads=2, dim_feedforward=128,
                      dropout=0.1, use_alibi=True).to(dev).train()
opt = T.FusedAdamW(m.parameters(), lr=1e-3)
bags = torch.randn(1, 333, 64, device=dev)
coords = torch.rand(1, 333, 2, device=dev) * 5000
y = torch.tensor([[0.0, 1.0]], device=dev)
loss = T.data_parallel_step(m, opt, (bags, coords, None, y), None)
m.eval()
with torch.inference_mode():
    out = m(bags, coords=coords, mask=None)
tiles = torch.randint(0, 255, (3, 224, 224, 3), dtype=torch.uint8, device=dev)
cnt = canny_edge_counts(tiles)
torch.cuda.synchronize()
print("probe ok", float(loss), out.tolist(), cnt.tolist())
