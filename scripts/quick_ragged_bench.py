"""MIL deploy over a cohort of bags: per-bag forwards on three streams (predict_bags) vs ragged batches
(predict_bags_ragged); equal 4096-tile bags (the bench shape) and bags of 2 000 .. 10 000 tiles."""
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch

from stamp_b200.deploy import predict_bags, predict_bags_ragged
from stamp_b200.mil import VisionTransformer

dev = torch.device("cuda:0")
torch.manual_seed(0)
model = VisionTransformer(dim_output=2, dim_input=1024, dim_model=512, n_layers=2, n_heads=8, dim_feedforward=512,
                          dropout=0.25, use_alibi=True).to(dev).eval()
g = torch.Generator().manual_seed(0)
for name, lengths in (("64 x 4096 tiles", [4096] * 64),
                      ("64 bags of 2000..10000 tiles", torch.randint(2000, 10001, (64,), generator=g).tolist())):
    host = [(torch.randn(n, 1024, generator=g).half().pin_memory(), (torch.rand(n, 2, generator=g) * 5e4).pin_memory())
            for n in lengths]
    devb = [(f.to(dev), c.to(dev)) for f, c in host]
    for label, fn, data in (("per-bag streams, device bags", predict_bags, devb), ("ragged, device bags", predict_bags_ragged, devb),
                            ("per-bag streams, host bags", predict_bags, host), ("ragged, host bags", predict_bags_ragged, host)):
        ref = fn(model, iter(data), dev)
        torch.cuda.synchronize()
        best = 1e9
        for _ in range(3):
            t0 = time.perf_counter()
            out = fn(model, iter(data), dev)
            torch.cuda.synchronize()
            best = min(best, time.perf_counter() - t0)
        print(f"{name:30s} {label:32s} {len(lengths) / best:8.0f} slides/s  {sum(lengths) / best / 1e6:6.2f} M tiles/s")
    a, b = predict_bags(model, iter(devb), dev, graphs=False), predict_bags_ragged(model, iter(devb), dev)
    print("   max |dp| between the two paths:", float((a - b).abs().max()))
