"""Quick device-resident tiles/s probe for the ViT tile encoder (development aid, not bench.py)."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch

from stamp_b200.vit import TileEncoder, UNI_ARCH, VIRCHOW2_ARCH, random_state_dict

arch = VIRCHOW2_ARCH if "virchow2" in sys.argv else UNI_ARCH
dev = torch.device("cuda:0")
if "persist" in sys.argv:
    from stamp_b200 import _lib
    _lib.load().stamp_b200_attention_tc_enable(3)
    print("attention: persistent single-TMEM-pass kernel")
sd = random_state_dict(arch)
for B in [int(a) for a in sys.argv[1:] if a.isdigit()] or [64, 128, 256]:
    enc = TileEncoder(arch, sd, max_batch=B).to(dev).eval()
    tiles = torch.randint(0, 255, (B, 224, 224, 3), dtype=torch.uint8, device=dev)
    for _ in range(3):
        enc(tiles)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 10
    e0.record()
    for _ in range(n):
        enc(tiles)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    tps = B / ms * 1e3
    print(f"{arch.name} B={B}: {ms:.2f} ms/batch, {tps:.0f} tiles/s, {tps * arch.flops_per_tile() / 1e12:.1f} TFLOP/s")
