"""One warm MIL training step at the configs[3] shape, for ncu launch lists (development aid)."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch

from stamp_b200 import train as T
from stamp_b200.mil import VisionTransformer

dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
n_tiles = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
torch.manual_seed(0)
mil = VisionTransformer(dim_output=2, dim_input=1024, dim_model=512, n_layers=2, n_heads=8, dim_feedforward=512,
                        dropout=0.25, use_alibi=True).to(dev).train()
opt, sched = T.configure_optimizers(mil, total_steps=1000)
bags = torch.randn(B, n_tiles, 1024, device=dev).half().float()
coords = torch.randint(0, 100, (B, n_tiles, 2), device=dev).float() * 256.0
targets = torch.nn.functional.one_hot(torch.arange(B, device=dev) % 2, 2).float()
for _ in range(2):
    opt.zero_grad()
    T.training_step(mil, (bags, coords, None, targets), None).backward()
    opt.step()
    sched.step()
torch.cuda.synchronize()
