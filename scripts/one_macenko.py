"""Macenko over one batch of H&E-like tiles: timing, or one warm call for ncu launch lists (development aid)."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch

from bench_extra import synthetic_he_tiles
from stamp_b200.macenko import macenko_normalize

dev = torch.device("cuda:0")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 768
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 0
tiles = synthetic_he_tiles(n, 3, dev)
out = torch.empty_like(tiles)
for _ in range(2):
    macenko_normalize(tiles, out=out)
torch.cuda.synchronize()
if reps:
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        macenko_normalize(tiles, out=out)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / reps
    print(f"macenko {n} tiles: {us:.1f} us, {n / us * 1e6:.0f} tiles/s, {n * 301056 / us / 1e3:.1f} GB/s algorithmic")
