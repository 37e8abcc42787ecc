"""Quick probe of the HBM-bound kernels: Macenko (tiles/s, GB/s) and CHIEF pooling (GB/s)."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch

from stamp_b200 import _lib
from stamp_b200.encoder import GatedAttentionPool, topk
from stamp_b200.macenko import macenko_normalize

dev = torch.device("cuda:0")


def timeit(fn, n=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


for B in (96, 192, 768):
    tiles = torch.randint(20, 235, (B, 224, 224, 3), dtype=torch.uint8, device=dev)
    out = torch.empty_like(tiles)
    ms = timeit(lambda: macenko_normalize(tiles, out=out))
    print(f"macenko B={B}: {ms:.3f} ms, {B / ms * 1e3:.0f} tiles/s, algorithmic {2 * tiles.numel() / ms / 1e6:.0f} GB/s")

sd = {"attention_net.0.weight": torch.randn(512, 768) * 0.04, "attention_net.0.bias": torch.zeros(512),
      "attention_net.3.attention_a.0.weight": torch.randn(256, 512) * 0.05, "attention_net.3.attention_a.0.bias": torch.zeros(256),
      "attention_net.3.attention_b.0.weight": torch.randn(256, 512) * 0.05, "attention_net.3.attention_b.0.bias": torch.zeros(256),
      "attention_net.3.attention_c.weight": torch.randn(1, 256) * 0.06, "attention_net.3.attention_c.bias": torch.zeros(1)}
pool = GatedAttentionPool(sd).to(dev)
for N in (10000, 50000):
    x = torch.randn(N, 768, device=dev)
    _lib.profile_enable(True)
    pool(x)
    prof = _lib.profile_summary()
    _lib.profile_enable(False)
    ms = timeit(lambda: pool(x))
    print(f"chief pool N={N}: whole {ms:.3f} ms ({N * 768 * 4 / ms / 1e6:.0f} GB/s of x); pooling kernels alone {prof['pool']['ms']:.3f} ms "
          f"({prof['pool']['work'] / prof['pool']['ms'] / 1e6:.0f} GB/s), gemm {prof['gemm']['ms']:.3f} ms")
    s = torch.randn(N, device=dev)
    print(f"topk N={N} k=25: {timeit(lambda: topk(s, 25)):.3f} ms")
