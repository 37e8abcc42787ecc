"""Device time of the GigaPath transform's resampling (Resize(256, bicubic) + CenterCrop(224)) on 768 H&E-like tiles."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch

from bench_extra import synthetic_he_tiles
from stamp_b200.resize import resize_center_crop

dev = torch.device("cuda:0")
tiles = synthetic_he_tiles(768, 3, dev)
for _ in range(3):
    resize_center_crop(tiles, 256, 224)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    resize_center_crop(tiles, 256, 224)
e1.record()
torch.cuda.synchronize()
us = e0.elapsed_time(e1) / 20 * 1e3
print(f"resize 768 tiles: {us:.1f} us, {768 / us:.2f} M tiles/s, {2 * tiles.numel() / us / 1e3:.0f} GB/s")
