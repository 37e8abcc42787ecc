"""Small runs of every kernel family for compute-sanitizer (development aid / evidence under profiles/):

    compute-sanitizer --tool memcheck python scripts/sanitize_probe.py

MIL training step and inference forward at a ragged bag length (333 tiles: tcgen05 long-bag attention forward v3,
backward, weight gradients, row ops, fused AdamW), the ViT tile encoder for head dimensions 64 and 80 (GEMM epilogues
incl. GELU / SwiGLU / residual reduce-add, streaming attention, LayerNorm, patch kernel) on a ragged batch, Macenko,
the tissue-texture filter, the bicubic resampling, CHIEF pooling + top-k, the JPEG tile decode, a ragged MIL batch, the
TransMIL and barspoon aggregators."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch

from stamp_b200 import train as T
from stamp_b200.encoder import GatedAttentionPool, topk
from stamp_b200.macenko import macenko_normalize
from stamp_b200.mil import VisionTransformer
from stamp_b200.resize import resize_center_crop
from stamp_b200.tiling import canny_edge_counts
from stamp_b200.vit import TileEncoder, VitArch, random_state_dict

dev = torch.device("cuda:0")
torch.manual_seed(0)
m = VisionTransformer(dim_output=2, dim_input=64, dim_model=128, n_layers=1, n_heads=2, dim_feedforward=128,
                      dropout=0.1, use_alibi=True).to(dev).train()
opt = T.FusedAdamW(m.parameters(), lr=1e-3)
bags = torch.randn(1, 333, 64, device=dev)
coords = torch.rand(1, 333, 2, device=dev) * 5000
y = torch.tensor([[0.0, 1.0]], device=dev)
loss = T.data_parallel_step(m, opt, (bags, coords, None, y), None)
m.eval()
with torch.inference_mode():
    out = m(bags, coords=coords, mask=None)
    out16 = m(bags.half(), coords=coords, mask=None)
torch.cuda.synchronize()
print("mil ok", float(loss), out.tolist(), out16.tolist())

tiles = torch.randint(0, 255, (5, 224, 224, 3), dtype=torch.uint8, device=dev)
for arch in (VitArch("hd64", dim=128, depth=2, heads=2, mlp_hidden=512),
             VitArch("hd80", patch=14, dim=160, depth=2, heads=2, mlp_hidden=432, mlp="swiglu", reg_tokens=4),
             VitArch("resized", dim=128, depth=1, heads=2, mlp_hidden=512, pre_resize=256)):
    enc = TileEncoder(arch, random_state_dict(arch), max_batch=3).to(dev).eval()
    f = enc(tiles)
    torch.cuda.synchronize()
    print("vit ok", arch.name, f.shape, bool(torch.isfinite(f).all()))

norm = macenko_normalize(tiles)
cnt = canny_edge_counts(tiles)
rs = resize_center_crop(tiles, 112, 97)
g = torch.Generator().manual_seed(3)
sd = {"attention_net.0.weight": torch.randn(512, 768, generator=g) * 0.04, "attention_net.0.bias": torch.zeros(512),
      "attention_net.3.attention_a.0.weight": torch.randn(256, 512, generator=g) * 0.05,
      "attention_net.3.attention_a.0.bias": torch.zeros(256),
      "attention_net.3.attention_b.0.weight": torch.randn(256, 512, generator=g) * 0.05,
      "attention_net.3.attention_b.0.bias": torch.zeros(256),
      "attention_net.3.attention_c.weight": torch.randn(1, 256, generator=g) * 0.06,
      "attention_net.3.attention_c.bias": torch.zeros(1)}
pooled = GatedAttentionPool(sd).to(dev)(torch.randn(777, 768, device=dev))
vals, idx = topk(pooled["attention_raw"].squeeze(0).contiguous(), 25)
torch.cuda.synchronize()
print("side kernels ok", norm.shape, cnt.tolist(), rs.shape, idx[:3].tolist())

import io

from PIL import Image

from stamp_b200.barspoon import EncDecTransformer
from stamp_b200.jpeg import decode_jpeg_tiles
from stamp_b200.transmil import TransMIL

blobs = []
for t in tiles[:3].cpu().numpy():
    b = io.BytesIO()
    Image.fromarray(t[:100, :77]).save(b, format="jpeg")
    blobs.append(b.getvalue())
dec = decode_jpeg_tiles(blobs, dev, max_workers=1)
rag = [(torch.randn(n, 64).half(), torch.rand(n, 2) * 5000) for n in (300, 1, 77, 515)]
tok, crd, seq, smax = m.pack_ragged(rag, pin=False)
with torch.inference_mode():
    lr = m.forward_ragged(tok.to(dev), crd.to(dev), seq.to(dev), smax)
    tm = TransMIL(3, 64, 512).to(dev).eval()(torch.randn(2, 333, 64, device=dev))
    bs = EncDecTransformer(64, {"a": 2, "b": 3}, d_model=128, num_encoder_heads=2, num_decoder_heads=2,
                           dim_feedforward=256).to(dev).eval()(torch.randn(1, 200, 64, device=dev), torch.rand(1, 200, 2, device=dev) * 1e4)
torch.cuda.synchronize()
print("new paths ok", dec.shape, lr.shape, tm.shape, {k: v.shape for k, v in bs.items()})
