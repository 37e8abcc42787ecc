"""Development aid: predict_bags throughput with device-resident fp16 bags for 1..4 CUDA streams, and with host bags."""
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch

from stamp_b200.deploy import predict_bags
from stamp_b200.mil import VisionTransformer

dev = torch.device("cuda:0")
n_tiles, n_bags = 4096, 32
mil = VisionTransformer(dim_output=2, dim_input=1024, dim_model=512, n_layers=2, n_heads=8, dim_feedforward=512,
                        dropout=0.25, use_alibi=True).to(dev).eval()
bags = torch.randn(n_bags, n_tiles, 1024, device=dev).half()
coords = torch.randint(0, 100, (n_bags, n_tiles, 2), device=dev).float() * 256.0
hb, hc = bags.cpu().pin_memory(), coords.cpu().pin_memory()
for src, (B, C) in (("device", (bags, coords)), ("host", (hb, hc))):
    for ns in (1, 2, 3, 4):
        predict_bags(mil, ((B[i], C[i]) for i in range(n_bags)), dev, n_streams=ns)
        torch.cuda.synchronize()
        t = time.perf_counter()
        for _ in range(3):
            predict_bags(mil, ((B[i], C[i]) for i in range(n_bags)), dev, n_streams=ns)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t) / (3 * n_bags)
        print(f"{src:6s} bags, {ns} stream(s): {dt * 1e3:.3f} ms/bag, {1 / dt:.0f} slides/s", flush=True)
# host-side cost of one forward call (no sync inside): how far is the launch path from the GPU time?
with torch.inference_mode():
    t = time.perf_counter()
    for i in range(n_bags):
        mil(bags[i:i + 1], coords=coords[i:i + 1], mask=None)
    t_issue = (time.perf_counter() - t) / n_bags
    torch.cuda.synchronize()
print(f"host time to issue one forward: {t_issue * 1e6:.0f} us")
