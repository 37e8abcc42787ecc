"""Extra blocks of the bench line (imported by bench.py): the configurations of BASELINE.json beyond the headline,
run as specified, and the honest denominators.

* ``cohort_block``      configs[2]: the 256-slide Virchow2 + Macenko cohort, slides of 2 000 ... 10 000 H&E-like tiles
                        (SURVEY.md 8d generator), slide-sharded with the size-aware LPT assignment; reports the
                        measured makespan and, from the same per-slide times, what round-robin would have cost.
* ``crossval_block``    configs[3]: 5-fold cross-validation over 320 synthetic patients with 4096 x 1024 bags,
                        fold-per-GPU and data-parallel-in-fold (stamp_b200.crossval).
* ``torch_gpu_block``   the reference's own GPU path -- eager torch on the same B200 -- for the ViT-L/16 forward
                        (src/stamp/preprocessing/__init__.py:322-327 runs it in fp32, TF32 off) and the MIL aggregator
                        forward / training step (src/stamp/modeling/train.py:519: matmul precision "high").  timm and
                        lightning are absent offline, so the arithmetic is the oracle restatement of the reference's
                        modules (test infrastructure, used here as the timed baseline only).
"""

from __future__ import annotations

import time

import torch

HE_REF = [[0.5626, 0.2159], [0.7201, 0.8012], [0.4062, 0.5581]]


def synthetic_he_tiles(n: int, seed: int, device, img: int = 224) -> torch.Tensor:
    """SURVEY.md 8d tile generator on the device (same recipe as oracle/vit_oracle.synthetic_tiles): stain
    concentrations c_H, c_E ~ Gamma(2, 0.5) smoothed 5x5, OD = HERef . c, I = clip(240 exp(-OD) + N(0, 2))."""
    g = torch.Generator(device=device).manual_seed(seed)
    out = torch.empty((n, img, img, 3), dtype=torch.uint8, device=device)
    he = torch.tensor(HE_REF, device=device)
    for s in range(0, n, 512):
        m = min(512, n - s)
        u = torch.rand(2, m, 2, img, img, device=device, generator=g).clamp_min_(1e-12)
        c = -0.5 * (u[0].log() + u[1].log())
        c = torch.nn.functional.avg_pool2d(torch.nn.functional.pad(c, (2, 2, 2, 2), mode="reflect"), 5, stride=1)
        od = torch.einsum("ck,nkhw->nhwc", he, c)
        im = 240.0 * torch.exp(-od) + 2.0 * torch.randn(m, img, img, 3, device=device, generator=g)
        out[s:s + m] = im.clamp_(0, 255).round_().to(torch.uint8)
    return out


def cohort_sizes(n_slides: int = 256, seed: int = 7) -> list[int]:
    """N_tiles ~ U{2000 ... 10000}, seed 7 (SURVEY.md 8d configs[2])."""
    g = torch.Generator().manual_seed(seed)
    return torch.randint(2000, 10001, (n_slides,), generator=g).tolist()


def cohort_block(dev, rank: int, world: int, peak_tf: float, slides_per_gpu: int = 32, batch: int = 96,
                 tile_pool: int = 2048) -> dict | None:
    """configs[2].  The cohort is the first ``slides_per_gpu * world`` slides of the 256-slide cohort (all 256 at
    8 GPUs: weak scaling, ~190 k tiles per GPU).  Every slide is a different sequence of H&E-like tiles drawn from a
    device-resident pool (a slide of 10 000 distinct tiles is 1.5 GB; the pool is re-indexed per slide, the kernels
    see ordinary [batch, 224, 224, 3] uint8 tensors).  Per slide: Macenko fitted per batch of tiles + ViT-H/14."""
    import torch.distributed as dist

    from stamp_b200.extractor import virchow2
    from stamp_b200.macenko import macenko_normalize
    from stamp_b200.sharding import shard_lpt, shard_round_robin
    from stamp_b200.vit import VIRCHOW2_ARCH

    sizes = cohort_sizes()[: min(256, slides_per_gpu * world)]
    ids = list(range(len(sizes)))
    mine = shard_lpt(ids, sizes, rank, world)
    model = virchow2(weights="random", max_batch=batch).model.to(dev).eval()
    pool = synthetic_he_tiles(tile_pool, seed=1000 + rank, device=dev)
    norm = torch.empty((batch, 224, 224, 3), dtype=torch.uint8, device=dev)
    g = torch.Generator(device=dev).manual_seed(rank)

    def run_slide(n_tiles: int) -> float:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        idx = torch.randint(0, tile_pool, (n_tiles,), device=dev, generator=g)
        e0.record()
        for s in range(0, n_tiles, batch):
            tiles = pool[idx[s:s + batch]]                     # this slide's next batch of tiles (device gather)
            macenko_normalize(tiles, out=norm[: tiles.shape[0]])
            model(norm[: tiles.shape[0]])
        e1.record()
        e1.synchronize()
        return e0.elapsed_time(e1) * 1e-3

    run_slide(2 * batch)                                       # warm-up
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    times = torch.zeros(len(sizes), dtype=torch.float64, device=dev)
    for i in mine:
        times[i] = run_slide(sizes[i])
    torch.cuda.synchronize()
    mine_s = time.perf_counter() - t0
    span = torch.tensor([mine_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(times)                                 # host-side bookkeeping, not on the data path
        spans = [torch.zeros_like(span) for _ in range(world)]
        dist.all_gather(spans, span)
        spans = [float(s) for s in spans]
    else:
        spans = [mine_s]
    if rank != 0:
        return None
    t = times.cpu().tolist()
    total_tiles = sum(sizes)

    def makespan(assign) -> tuple[float, float]:
        loads = [sum(t[i] for i in assign(r)) for r in range(world)]
        return max(loads), max(loads) / (sum(loads) / world)

    lpt_ms, lpt_imb = makespan(lambda r: shard_lpt(ids, sizes, r, world))
    rr_ms, rr_imb = makespan(lambda r: shard_round_robin(ids, r, world))
    tps = total_tiles / max(spans)
    return {
        "metric": "tiles/sec, Macenko + Virchow2 ViT-H/14, slide-sharded cohort (BASELINE configs[2])",
        "slides": len(sizes), "tiles": total_tiles, "n_gpus": world, "value": tps, "unit": "tiles/s",
        "per_gpu_tiles_per_s": tps / world, "batch": batch,
        "roofline_frac": tps / world * VIRCHOW2_ARCH.flops_per_tile() / 1e12 / peak_tf,
        "sharding": "shard_lpt (size-aware, no collective)", "makespan_s": max(spans),
        "slowest_rank_over_mean": max(spans) / (sum(spans) / world),
        "lpt_from_slide_times": {"makespan_s": lpt_ms, "slowest_over_mean": lpt_imb},
        "round_robin_from_slide_times": {"makespan_s": rr_ms, "slowest_over_mean": rr_imb},
        "data": "H&E-like tiles (SURVEY 8d generator), 2000..10000 tiles per slide (seed 7), stain fit per batch of tiles",
    }


def crossval_block(dev, rank: int, world: int, n_patients: int = 320, n_tiles: int = 4096, dim: int = 1024,
                   epochs: int = 4) -> dict | None:
    """configs[3]: 5-fold cross-validation, ALiBi Transformer-MIL, bf16 training, bags of 4096 x 1024 resident in HBM.
    Patients per SURVEY.md 8d: feats ~ N(0,1) fp16, coords = random cells of a 100 x 100 grid x 256 um, binary label
    with a planted signal (mean shift 0.5 on 5 % of the tiles of class 1)."""
    import torch.distributed as dist

    from stamp_b200.crossval import Patient, crossval

    pats = []
    for i in range(n_patients):
        g = torch.Generator(device=dev).manual_seed(10_000 + i)
        f = torch.randn(n_tiles, dim, device=dev, generator=g)
        label = i % 2
        if label:
            f[: n_tiles // 20] += 0.5
        cells = torch.randperm(100 * 100, device=dev, generator=g)[:n_tiles]
        c = torch.stack([(cells % 100).float(), (cells // 100).float()], dim=-1) * 256.0
        pats.append(Patient(f"pat{i:03d}", f.half(), c, label))
    kw = dict(n_splits=5, n_classes=2, dim_input=dim, bag_size=n_tiles, batch_size=64, max_epochs=epochs,
              patience=epochs, max_lr=1e-4, seed=0,
              model_params=dict(dim_model=512, n_layers=2, n_heads=8, dim_feedforward=512, dropout=0.25, use_alibi=True))
    out = {"metric": "5-fold cross-validation wall time, ALiBi Transformer-MIL bf16, 320 patients x 4096 x 1024 "
                     "(BASELINE configs[3])", "n_gpus": world, "epochs_per_fold": epochs, "batch_size": 64,
           "unit": "s"}
    label = {p.pid: p.label for p in pats}

    def run(mode: str) -> dict:
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        res = crossval(pats, mode=mode, **kw)
        torch.cuda.synchronize()
        mine = time.perf_counter() - t0
        span = torch.tensor([mine], dtype=torch.float64, device=dev)
        hits = torch.tensor([sum(int(r.probs[i].argmax()) == label[pid] for r in res
                                 for i, pid in enumerate(r.test_patients)),
                             sum(len(r.test_patients) for r in res)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(span, op=dist.ReduceOp.MAX)
            if mode == "fold_per_gpu":
                dist.all_reduce(hits)
        steps = sum(r.train_steps for r in res)
        return {"wall_s": float(span), "folds_on_this_rank": [r.fold for r in res], "train_steps_rank0": steps,
                "held_out_accuracy": float(hits[0] / max(1.0, float(hits[1]))),
                "fold_seconds_rank0": [round(r.seconds, 3) for r in res]}

    out["fold_per_gpu"] = run("fold_per_gpu")
    out["fold_per_gpu"]["note"] = ("folds[rank::world], no collective; ranks beyond the fifth idle" if world > 5
                                   else "folds[rank::world], no collective")
    if world > 1:
        out["dp_in_fold"] = run("dp_in_fold")
        out["dp_in_fold"]["note"] = "every fold by all ranks: batch of 64 bags split over the ranks, one all-reduce per step"
    # run-to-run reproducibility of one fold on one GPU (fp32 atomic reductions in a few gradient kernels make the
    # training deterministic only up to summation order): max |delta p| between two trainings of fold 0
    if rank == 0:
        from stamp_b200.crossval import train_fold
        from stamp_b200.sharding import crossval_splits

        by = {p.pid: p for p in pats}
        tr, te = crossval_splits([p.pid for p in pats], [p.label for p in pats], 5)[0]
        fk = {k: v for k, v in kw.items() if k not in ("n_splits",)}
        fk.update(max_epochs=1, patience=1)
        a = train_fold(0, [by[i] for i in tr], [by[i] for i in te], data_parallel=False, **fk)
        b = train_fold(0, [by[i] for i in tr], [by[i] for i in te], data_parallel=False, **fk)
        out["repeat_fold0_max_abs_prob_diff"] = float((a.probs - b.probs).abs().max())
    if world > 1:
        dist.barrier()
    return out if rank == 0 else None


def _timed(fn, reps: int, warm: int = 1) -> float:
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e-3 / reps


def torch_gpu_block(dev, ours: dict) -> dict:
    """Eager torch on the same GPU: the reference's own GPU path for the three hot-path stages.  ``ours`` holds this
    framework's numbers for the ratios ({"vit_tiles_per_s", "mil_slides_per_s", "mil_train_bags_per_s"})."""
    from oracle import mil_oracle
    from oracle import vit_oracle as vo

    out: dict = {"note": "eager PyTorch on the same B200; arithmetic = oracle restatement of timm ViT-L/16 and of the "
                         "reference's VisionTransformer (timm / lightning are not installable offline)"}
    prev_tf32 = torch.backends.cuda.matmul.allow_tf32
    prev_prec = torch.get_float32_matmul_precision()
    try:
        # ---- ViT-L/16, batch 64 as extract_ runs it: fp32, TF32 off, model(tiles).half() ----
        w = {k: v.to(dev) for k, v in vo.make_weights(vo.UNI).items()}
        tiles = vo.synthetic_tiles(64, seed=0).to(dev)
        torch.backends.cuda.matmul.allow_tf32 = False
        torch.set_float32_matmul_precision("highest")
        with torch.inference_mode():
            t = _timed(lambda: vo.forward(w, vo.UNI, tiles).half(), reps=3)
            out["vit_l16_fp32"] = {"tiles_per_s": 64 / t, "batch": 64, "dtype": "fp32 (TF32 off, as extract_ runs it)"}
            with torch.autocast("cuda", dtype=torch.bfloat16):
                t = _timed(lambda: vo.forward(w, vo.UNI, tiles).half(), reps=5)
            out["vit_l16_bf16_autocast"] = {"tiles_per_s": 64 / t, "batch": 64, "dtype": "bf16 autocast (best case)"}
        del w, tiles
        # ---- MIL aggregator, one 4096 x 1024 bag, ALiBi: deploy.py:398 runs predictions at precision "medium" ----
        sd = {k: v.to(dev) for k, v in mil_oracle.init_state_dict(dim_input=1024, dim_output=2, seed=0,
                                                                  running_mean=6000.0).items()}
        bags, coords = mil_oracle.synthetic_bag(4096, 1024, seed=0)
        bags, coords = bags.to(dev), coords.to(dev)
        torch.set_float32_matmul_precision("medium")
        torch.backends.cuda.matmul.allow_tf32 = True
        with torch.inference_mode():
            t = _timed(lambda: mil_oracle.forward(sd, bags, coords, None), reps=5)
        out["mil_forward_4096"] = {"slides_per_s": 1 / t, "dtype": "fp32 storage, TF32 matmuls (precision 'medium')"}
        # ---- MIL training step, 8 bags of 4096 x 1024: train.py:519 precision "high" (TF32), AdamW ----
        torch.set_float32_matmul_precision("high")
        params = {k: v.clone().requires_grad_(True) for k, v in sd.items() if "scale_distance" not in k}
        opt = torch.optim.AdamW(list(params.values()), lr=1e-4)
        tb, tc = mil_oracle.synthetic_bag(4096, 1024, seed=1, batch=8)
        tb, tc = tb.to(dev), tc.to(dev)
        ty = torch.nn.functional.one_hot(torch.arange(8, device=dev) % 2, 2).float()

        def step():
            opt.zero_grad(set_to_none=True)
            loss = 0.0
            for b in range(8):       # the S x S intermediates of 8 x 8 heads do not fit next to each other: bag by bag
                logits = mil_oracle.forward({**sd, **params}, tb[b:b + 1], tc[b:b + 1], None)
                term = mil_oracle.cross_entropy(logits, ty[b:b + 1], None) / 8
                term.backward()
                loss += float(term.detach())
            opt.step()
            return loss

        t = _timed(step, reps=2)
        out["mil_train_step_8x4096"] = {"bags_per_s": 8 / t, "dtype": "fp32 storage, TF32 matmuls (precision 'high')",
                                        "note": "gradient accumulation over the 8 bags of the batch (S x S temporaries)"}
    finally:
        torch.backends.cuda.matmul.allow_tf32 = prev_tf32
        torch.set_float32_matmul_precision(prev_prec)
    r = {}
    if ours.get("vit_tiles_per_s"):
        r["vit_vs_fp32"] = ours["vit_tiles_per_s"] / out["vit_l16_fp32"]["tiles_per_s"]
        r["vit_vs_bf16_autocast"] = ours["vit_tiles_per_s"] / out["vit_l16_bf16_autocast"]["tiles_per_s"]
    if ours.get("mil_slides_per_s"):
        r["mil_forward"] = ours["mil_slides_per_s"] / out["mil_forward_4096"]["slides_per_s"]
    if ours.get("mil_train_bags_per_s"):
        r["mil_train"] = ours["mil_train_bags_per_s"] / out["mil_train_step_8x4096"]["bags_per_s"]
    out["speedup_of_this_framework"] = r
    torch.cuda.empty_cache()
    return out


def encoding_block(dev, rank: int, world: int, slides_per_gpu: int = 32) -> dict | None:
    """SURVEY.md 8e, slide-level encoding: the cohort's slides (2 000..10 000 tiles of 768-d features, seed 7) are
    sharded by tile count (``shard_lpt``, no collective); every rank runs the reference's ``encode_slides_`` walk
    (encoding/encoder/__init__.py:42-93) over ITS feature files with the CHIEF encoder: read ``.h5`` -> host->device
    -> gated-attention pooling kernels -> device->host -> slide-level ``.h5``.  The per-rank feature files are written
    before the timed region (they are the input of this stage)."""
    import shutil
    import tempfile
    from pathlib import Path

    import torch.distributed as dist

    from stamp_b200 import features
    from stamp_b200.encoder import ChiefB200
    from stamp_b200.sharding import shard_lpt

    sizes = cohort_sizes()[: min(256, slides_per_gpu * world)]
    mine = shard_lpt(list(range(len(sizes))), sizes, rank, world)
    gen = torch.Generator().manual_seed(3)
    sd = {"attention_net.0.weight": torch.randn(512, 768, generator=gen) * 0.04, "attention_net.0.bias": torch.zeros(512),
          "attention_net.3.attention_a.0.weight": torch.randn(256, 512, generator=gen) * 0.05,
          "attention_net.3.attention_a.0.bias": torch.zeros(256),
          "attention_net.3.attention_b.0.weight": torch.randn(256, 512, generator=gen) * 0.05,
          "attention_net.3.attention_b.0.bias": torch.zeros(256),
          "attention_net.3.attention_c.weight": torch.randn(1, 256, generator=gen) * 0.06,
          "attention_net.3.attention_c.bias": torch.zeros(1)}
    enc = ChiefB200(sd)
    root = Path(tempfile.mkdtemp(prefix=f"stamp_b200_enc_r{rank}_"))
    try:
        feat_dir = root / "chief-ctranspath"
        pool = torch.randn(12_000, 768, generator=gen).half()
        for i in mine:
            off = (i * 37) % 2000
            features.write_tile_features(feat_dir / f"slide_{i:03d}.h5", pool[off:off + sizes[i]],
                                         torch.rand(sizes[i], 2, generator=gen) * 5e4, extractor="chief-ctranspath",
                                         tile_size_um=256.0, tile_size_px=224)
        if hasattr(enc, "_encode_dir"):                         # stand-alone Encoder (no stamp package): warm-up
            enc.encode_slides_(root / "warm", feat_dir, dev, generate_hash=False)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        enc.encode_slides_(root / "out", feat_dir, dev, generate_hash=False)
        torch.cuda.synchronize()
        span = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        n_out = len(list((root / "out").rglob("*.h5")))
    finally:
        shutil.rmtree(root, ignore_errors=True)
    if world > 1:
        dist.all_reduce(span, op=dist.ReduceOp.MAX)
    if rank != 0:
        return None
    assert n_out == len(mine), (n_out, len(mine))
    secs = float(span)
    return {"metric": "slide embeddings/sec, CHIEF encode_slides_ over feature files (SURVEY 8e: slide-level encoding)",
            "slides": len(sizes), "tiles": sum(sizes), "n_gpus": world, "value": len(sizes) / secs, "unit": "slides/s",
            "tiles_per_s": sum(sizes) / secs, "wall_s": secs,
            "sharding": "shard_lpt on tile counts, no collective",
            "note": "end to end per slide: fp16 feature file (h5lite) -> fp32 -> H2D -> pooling kernels -> D2H -> "
                    "slide-level .h5; bound by the host-side file read / fp16->fp32 conversion, the pooling kernels "
                    "take < 1 ms per slide"}


def cache_to_features_block(dev, n_tiles: int = 6144, batch: int = 192) -> dict:
    """One cached slide end to end, the way `stamp preprocess` sees a slide whose tiles are already in the tile cache
    (tiling.py:380-406 -> preprocessing/__init__.py:306-366): JPEG tile cache (zip) -> tiles -> Canny tissue filter ->
    UNI ViT-L/16 -> fp16 features + coordinates -> the slide's `.h5`.  Timed twice: tiles decoded by Pillow on a thread
    pool into pinned host memory (the reference's decoder) and tiles Huffman-decoded on the host, finished on the GPU."""
    import io
    import json
    import shutil
    import tempfile
    import zipfile
    from pathlib import Path

    from PIL import Image

    from stamp_b200 import features
    from stamp_b200.extractor import extract_cache_features, extract_slide_features, uni
    from stamp_b200.tiling import has_enough_texture, tiles_from_cache_file, tiles_from_cache_file_gpu

    ext = uni(weights="random", max_batch=batch)
    ext.model.to(dev).eval()
    root = Path(tempfile.mkdtemp(prefix="stamp_b200_cache_"))
    try:
        pool = synthetic_he_tiles(128, seed=11, device=dev).cpu().numpy()
        blobs = []
        for t in pool:
            b = io.BytesIO()
            Image.fromarray(t).save(b, format="jpeg")
            blobs.append(b.getvalue())
        path = root / "slide.zip"
        with zipfile.ZipFile(path, "w", compression=zipfile.ZIP_STORED) as zf:
            zf.writestr("tiler_params.json", json.dumps({"tile_ext": "jpg", "tile_size_um": 256.0, "tile_size_px": 224}))
            for i in range(n_tiles):
                zf.writestr(f"tile_({256.0 * (i % 64)}, {256.0 * (i // 64)}).jpg", blobs[i % len(blobs)])

        def run(gpu_decode: bool, out_name: str) -> tuple[float, int]:
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            if gpu_decode:
                tiles, coords, _ = tiles_from_cache_file_gpu(path, dev, max_workers=8)
                keep = has_enough_texture(tiles, 0.02)
                tiles, coords = tiles[keep], coords[keep.cpu()]
            else:
                tiles, coords, _ = tiles_from_cache_file(path, max_workers=8)
                keep = torch.cat([has_enough_texture(tiles[s:s + 768].to(dev), 0.02) for s in range(0, len(tiles), 768)]).cpu()
                tiles, coords = tiles[keep], coords[keep]
            feats = extract_slide_features(ext, tiles.contiguous(), dev, batch_size=batch)
            features.write_tile_features(root / out_name, feats, coords.numpy(), extractor="uni", tile_size_um=256.0,
                                         tile_size_px=224)
            torch.cuda.synchronize()
            return time.perf_counter() - t0, int(keep.sum())

        def run_pipelined(out_name: str) -> tuple[float, int]:
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            feats, coords, _ = extract_cache_features(ext, path, dev, batch_size=batch, canny_cutoff=0.02, max_workers=8)
            features.write_tile_features(root / out_name, feats, coords.numpy(), extractor="uni", tile_size_um=256.0,
                                         tile_size_px=224)
            torch.cuda.synchronize()
            return time.perf_counter() - t0, feats.shape[0]

        run(True, "warm.h5")
        run_pipelined("warm2.h5")
        t_pipe, kept_pipe = run_pipelined("pipe.h5")
        t_gpu, kept = run(True, "gpu.h5")
        t_pil, kept_p = run(False, "pil.h5")
        ref_feats = torch.from_numpy(features.read_tile_features(str(root / "pil.h5"))[0])
        same = all(torch.equal(torch.from_numpy(features.read_tile_features(str(root / n))[0]), ref_feats)
                   for n in ("gpu.h5", "pipe.h5")) and kept_pipe == kept
    finally:
        shutil.rmtree(root, ignore_errors=True)
    return {"metric": "cached slide -> feature file, tiles/s (JPEG tile cache -> tissue filter -> UNI ViT-L/16 -> .h5)",
            "tiles": n_tiles, "tiles_kept": kept, "pipelined_tiles_per_s": n_tiles / t_pipe,
            "gpu_decode_tiles_per_s": n_tiles / t_gpu,
            "pillow_decode_tiles_per_s": n_tiles / t_pil, "identical_feature_files": bool(same and kept == kept_p),
            "note": "8 host threads in both arms; GPU arm: Huffman on the host, IDCT / up-sampling / colour on the GPU, "
                    "tiles never leave HBM (pipelined: extractor.extract_cache_features, decode of batch i+1 overlaps the encoder on "
                    "batch i); Pillow arm: decoded tiles staged in pinned memory and streamed in"}


def extractors_block(dev, peak_tf: float, batch: int = 96) -> dict:
    """the other tile encoders the reference ships (uni2.py:18-32, h_optimus_0.py:14-28) on the same
    kernels -- device-resident tiles/s at batch 96 with random-init weights of the architecture."""
    from stamp_b200.vit import GIGAPATH_ARCH, H_OPTIMUS_ARCH, UNI2_ARCH, TileEncoder, random_state_dict

    out = {}
    tiles = torch.randint(0, 255, (batch, 224, 224, 3), dtype=torch.uint8, device=dev)
    for name, arch in (("uni2_h", UNI2_ARCH), ("h_optimus_0", H_OPTIMUS_ARCH), ("gigapath", GIGAPATH_ARCH)):
        enc = TileEncoder(arch, random_state_dict(arch), max_batch=batch).to(dev).eval()
        for _ in range(3):
            enc(tiles)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 8
        e0.record()
        for _ in range(reps):
            enc(tiles)
        e1.record()
        torch.cuda.synchronize()
        tps = batch * reps / (e0.elapsed_time(e1) * 1e-3)
        out[name] = {"tiles_per_s": tps, "batch": batch, "gflop_per_tile": arch.flops_per_tile() / 1e9,
                     "roofline_frac": tps * arch.flops_per_tile() / 1e12 / peak_tf}
        del enc
        torch.cuda.empty_cache()
    out["note"] = "ViT-H/14 (UNI2-h, 24 blocks), ViT-g/14 (H-optimus-0, 40 blocks) and ViT-g/16 (Prov-GigaPath, 40 " \
                  "blocks, incl. its Resize(256, bicubic) + CenterCrop(224) on the GPU): SwiGLU, 8 / 4 / 0 register " \
                  "tokens, head dimension 64, same GEMM / attention / LayerNorm kernels as UNI and Virchow2"
    return out


def aggregators_block(dev, hbm_gbs: float) -> dict:
    """the reference's other aggregators (models/mlp.py, models/trans_mil.py) on device-resident bags.
    The bag mean of MLP / Linear is the HBM-bound kernel: 64 bags x 4096 tiles x 1024 fp16 features = 537 MB per call
    (larger than L2), every byte read once."""
    from stamp_b200.mlp import MLP, bag_mean
    from stamp_b200.transmil import TransMIL

    out = {}
    try:
        bags = torch.randn(64, 4096, 1024, device=dev).half()
        t = _timed(lambda: bag_mean(bags), reps=20, warm=3)
        gbs = bags.numel() * 2 / t / 1e9
        out["bag_mean"] = {"shape": list(bags.shape), "dtype": "f16", "us_per_call": t * 1e6, "achieved_GBps": gbs,
                           "peak_GBps": hbm_gbs, "frac": gbs / hbm_gbs, "algorithmic_bytes": bags.numel() * 2}
        mlp = MLP(1024, 512, 2, 3, 0.25).to(dev).eval()
        with torch.inference_mode():
            t = _timed(lambda: mlp(bags), reps=20, warm=3)
        out["mlp"] = {"slides_per_s": bags.shape[0] / t, "tiles_per_slide": 4096, "dim_input": 1024,
                      "note": "bag mean + three fp32 Linear layers, 64 bags per call"}
        one = bags[:1].float()
        tm = TransMIL(2, 1024, 512).to(dev).eval()
        with torch.inference_mode():
            t = _timed(lambda: tm(one), reps=5, warm=2)
        out["transmil"] = {"slides_per_s": 1.0 / t, "tiles_per_slide": 4096, "dim_input": 1024,
                           "note": "_fc1 on the tcgen05 GEMM, Nystrom attention / pseudo-inverse / PPEG in fp32 on the CUDA cores, one bag per call"}
    except Exception as e:  # noqa: BLE001 -- an auxiliary block must not take the bench line down
        out["error"] = f"{type(e).__name__}: {e}"
    return out
