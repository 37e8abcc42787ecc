"""Development aid: the bag-mean kernel (csrc/bagmean.cu) timed with CUDA events at bag shapes larger than L2, and the
other-aggregators bench block."""
import json
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch

from bench_extra import _timed, aggregators_block
from stamp_b200.mlp import bag_mean

dev = torch.device("cuda:0")
for shape, dt in (((64, 4096, 1024), torch.float16), ((32, 4096, 1024), torch.float32), ((1, 50000, 1536), torch.float16),
                  ((512, 512, 768), torch.float16), ((8, 4096, 1024), torch.float16)):
    x = torch.randn(*shape, device=dev).to(dt)
    t = _timed(lambda: bag_mean(x), reps=20, warm=3)
    print(shape, dt, f"{t * 1e6:.1f} us  {x.numel() * x.element_size() / t / 1e9:.0f} GB/s", flush=True)
    del x
print(json.dumps(aggregators_block(dev, 6543.0)))
