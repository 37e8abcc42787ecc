"""Cached-tile JPEG decode: Pillow on one core / a thread pool vs host Huffman + GPU kernels (768 H&E-like tiles)."""
import io
import sys
import time
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import torch
from PIL import Image

from oracle import vit_oracle as vo
from stamp_b200 import jpeg

dev = torch.device("cuda:0")
blobs = []
for t in vo.synthetic_tiles(96, seed=1).numpy():
    b = io.BytesIO()
    Image.fromarray(t).save(b, format="jpeg")
    blobs.append(b.getvalue())
blobs = blobs * 8
n = len(blobs)


def pil(b):
    return np.asarray(Image.open(io.BytesIO(b)).convert("RGB"))


t0 = time.perf_counter(); [pil(b) for b in blobs]; t_pil1 = time.perf_counter() - t0
with ThreadPoolExecutor(8) as ex:
    t0 = time.perf_counter(); list(ex.map(pil, blobs)); t_pil8 = time.perf_counter() - t0
info, coef, quant = jpeg.entropy_decode(blobs, max_workers=8, pin=True)      # allocates the pinned staging once
t0 = time.perf_counter(); jpeg.entropy_decode(blobs, max_workers=1, out=(coef, quant)); t_h1 = time.perf_counter() - t0
t0 = time.perf_counter(); jpeg.entropy_decode(blobs, max_workers=8, out=(coef, quant)); t_h8 = time.perf_counter() - t0
cd, qd = coef.to(dev), quant.to(dev)
out = torch.empty((n, info.height, info.width, 3), dtype=torch.uint8, device=dev)
for _ in range(3):
    jpeg.decode_coefficients(info, cd, qd, out=out)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    jpeg.decode_coefficients(info, cd, qd, out=out)
e1.record(); torch.cuda.synchronize()
us = e0.elapsed_time(e1) / 10 * 1e3
t0 = time.perf_counter()
jpeg.entropy_decode(blobs, max_workers=8, out=(coef, quant))
jpeg.decode_coefficients(info, coef.to(dev, non_blocking=True), quant.to(dev, non_blocking=True), out=out)
torch.cuda.synchronize(); t_e2e = time.perf_counter() - t0
byts = cd.numel() * 2 + out.numel()
print(f"{n} tiles, {sum(map(len, blobs)) / n:.0f} B each")
print(f"Pillow 1 thread   {n / t_pil1:9.0f} tiles/s      8 threads {n / t_pil8:9.0f} tiles/s")
print(f"host Huffman 1 th {n / t_h1:9.0f} tiles/s      8 threads {n / t_h8:9.0f} tiles/s")
print(f"GPU kernels       {n / us * 1e6:9.0f} tiles/s  ({us:.0f} us, {byts / us / 1e3:.0f} GB/s of coefficient-in + RGB-out bytes)")
print(f"end to end (8 threads + H2D + kernels) {n / t_e2e:9.0f} tiles/s")
