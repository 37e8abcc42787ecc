"""One warm unit of work between cudaProfilerStart/Stop, for ncu launch lists:

    ncu --profile-from-start off --metrics gpu__time_duration.sum,... --csv --log-file out.csv \\
        python scripts/profile_region.py {vit_l16|virchow2|gigapath|resize|jpeg|mil_deploy|mil_train|macenko|transmil}
"""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch

dev = torch.device("cuda:0")
what = sys.argv[1]
torch.manual_seed(0)

if what in ("vit_l16", "virchow2"):
    from stamp_b200.vit import UNI_ARCH, VIRCHOW2_ARCH, TileEncoder, random_state_dict

    arch, B = (UNI_ARCH, 384) if what == "vit_l16" else (VIRCHOW2_ARCH, 96)   # the batches bench.py runs
    enc = TileEncoder(arch, random_state_dict(arch), max_batch=B).to(dev).eval()
    tiles = torch.randint(0, 255, (B, 224, 224, 3), dtype=torch.uint8, device=dev)
    unit = lambda: enc(tiles)
elif what == "gigapath":
    from stamp_b200.vit import GIGAPATH_ARCH, TileEncoder, random_state_dict

    enc = TileEncoder(GIGAPATH_ARCH, random_state_dict(GIGAPATH_ARCH), max_batch=96).to(dev).eval()
    tiles = torch.randint(0, 255, (96, 224, 224, 3), dtype=torch.uint8, device=dev)
    unit = lambda: enc(tiles)
elif what == "resize":
    from bench_extra import synthetic_he_tiles
    from stamp_b200.resize import resize_center_crop

    tiles = synthetic_he_tiles(768, 3, dev)
    unit = lambda: resize_center_crop(tiles, 256, 224)
elif what == "jpeg":
    import io

    from PIL import Image

    from bench_extra import synthetic_he_tiles
    from stamp_b200 import jpeg

    blobs = []
    for t in synthetic_he_tiles(96, 3, dev).cpu().numpy():
        b = io.BytesIO()
        Image.fromarray(t).save(b, format="jpeg")
        blobs.append(b.getvalue())
    info, coef, quant = jpeg.entropy_decode(blobs * 8, max_workers=8)
    cd, qd = coef.to(dev), quant.to(dev)
    out = torch.empty((768, 224, 224, 3), dtype=torch.uint8, device=dev)
    unit = lambda: jpeg.decode_coefficients(info, cd, qd, out=out)
elif what == "mil_deploy":
    from stamp_b200.mil import VisionTransformer

    mil = VisionTransformer(dim_output=2, dim_input=1024, dim_model=512, n_layers=2, n_heads=8, dim_feedforward=512,
                            dropout=0.25, use_alibi=True).to(dev).eval()
    bag = torch.randn(1, 4096, 1024, device=dev).half()
    coords = torch.randint(0, 100, (1, 4096, 2), device=dev).float() * 256.0

    def unit():
        with torch.inference_mode():
            mil(bag, coords=coords, mask=None)
elif what == "mil_train":
    from stamp_b200 import train as T
    from stamp_b200.mil import VisionTransformer

    mil = VisionTransformer(dim_output=2, dim_input=1024, dim_model=512, n_layers=2, n_heads=8, dim_feedforward=512,
                            dropout=0.25, use_alibi=True).to(dev).train()
    opt, sched = T.configure_optimizers(mil, total_steps=1000)
    bags = torch.randn(8, 4096, 1024, device=dev).half().float()
    coords = torch.randint(0, 100, (8, 4096, 2), device=dev).float() * 256.0
    targets = torch.nn.functional.one_hot(torch.arange(8, device=dev) % 2, 2).float()

    def unit():
        opt.zero_grad()
        T.training_step(mil, (bags, coords, None, targets), None).backward()
        opt.step()
        sched.step()
elif what == "macenko":
    from bench_extra import synthetic_he_tiles
    from stamp_b200.macenko import macenko_normalize

    tiles = synthetic_he_tiles(768, 3, dev)
    out = torch.empty_like(tiles)
    unit = lambda: macenko_normalize(tiles, out=out)
elif what == "transmil":
    from stamp_b200.transmil import TransMIL

    tm = TransMIL(2, 1024, 512).to(dev).eval()
    bag = torch.randn(1, 4096, 1024, device=dev)

    def unit():
        with torch.inference_mode():
            tm(bag)
else:
    raise SystemExit(__doc__)

for _ in range(3):
    unit()
torch.cuda.synchronize()
torch.cuda.profiler.start()
unit()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
