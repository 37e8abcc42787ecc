"""Quick MIL training-step probe (development aid): bags/s for forward + backward + AdamW at the
configs[3] shape (4096 x 1024 bags, default model), per-category kernel time split."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch

from stamp_b200 import _lib
from stamp_b200 import train as T
from stamp_b200.mil import VisionTransformer

dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
n_tiles = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
torch.manual_seed(0)
mil = VisionTransformer(dim_output=2, dim_input=1024, dim_model=512, n_layers=2, n_heads=8, dim_feedforward=512,
                        dropout=0.0, use_alibi=True).to(dev).train()
opt, sched = T.configure_optimizers(mil, total_steps=1000)
bags = torch.randn(B, n_tiles, 1024, device=dev).half().float()
coords = torch.randint(0, 100, (B, n_tiles, 2), device=dev).float() * 256.0
targets = torch.nn.functional.one_hot(torch.arange(B, device=dev) % 2, 2).float()


def step():
    opt.zero_grad()
    loss = T.training_step(mil, (bags, coords, None, targets), None)
    loss.backward()
    opt.step()
    sched.step()
    return loss


for _ in range(3):
    step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
reps = 5
e0.record()
for _ in range(reps):
    loss = step()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
print(f"MIL train B={B} N={n_tiles}: {ms:.2f} ms/step, {B * 1e3 / ms:.1f} bags/s, loss {float(loss):.4f}, "
      f"ctx {mil._train_ctx.numel() / 2**30:.2f} GiB")
_lib.profile_enable(True)
step()
prof = _lib.profile_summary()
_lib.profile_enable(False)
tot = sum(v["ms"] for v in prof.values())
for k, v in prof.items():
    if v["count"]:
        rate = v["work"] / (v["ms"] * 1e-3) / 1e12 if v["ms"] > 0 else 0.0
        print(f"  {k:10s} {v['ms']:8.2f} ms ({100 * v['ms'] / tot:4.1f} %)  n={v['count']:4d}  "
              f"{rate:8.2f} T{'FLOP' if k in ('gemm', 'attention') else 'B'}/s")
