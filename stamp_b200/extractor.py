"""Extractor registry entries backed by the B200 tile encoder.

Mirrors the reference's plugin interface (src/stamp/preprocessing/extractor/__init__.py:18-28):
``Extractor(model=, transform=, identifier=)`` is a frozen keyword-only dataclass; ``extract_``
accepts an instance directly (src/stamp/preprocessing/__init__.py:117,237-238), moves ``model`` to
the device, calls ``model(batch)`` under ``inference_mode`` and stores ``.half().cpu()``
(:243,322-327).  ``identifier`` is what lands in the output folder name and the ``extractor`` h5
attribute, so the factories below keep the reference's identifiers (the ``ExtractorName`` values
"uni", "virchow", "virchow-full", "virchow2", "uni2", "h-optimus-0", "h-optimus-1", "gigapath", "dino-bloom"; src/stamp/preprocessing/config.py:13-33).
When ``stamp`` is importable the factories return instances of the reference's own ``Extractor``.

The transform returns the tile as a uint8 HWC tensor (legal: the reference's ``empty`` extractor
does the same, src/stamp/preprocessing/extractor/empty.py:31-36); ToTensor + Normalize run on the
GPU inside the first kernel, which also cuts the host->device traffic 4x versus fp32 CHW tiles.
"""

from __future__ import annotations

from collections.abc import Callable, Mapping
from dataclasses import KW_ONLY, dataclass
from pathlib import Path
from typing import Generic, TypeVar

import numpy as np
import torch
from torch import Tensor, nn

from .vit import (DINOBLOOM_ARCH, GIGAPATH_ARCH, H_OPTIMUS_ARCH, UNI2_ARCH, UNI_ARCH, VIRCHOW2_ARCH, VIRCHOW_ARCH, VIRCHOW_FULL_ARCH,
                  TileEncoder, VitArch,
                  random_state_dict)

ExtractorModel = TypeVar("ExtractorModel", bound=nn.Module)

try:
    # The reference's own class whenever STAMP is importable: ``extract_`` dispatches on
    # ``case Extractor():`` (src/stamp/preprocessing/__init__.py:237-238), which only an instance of
    # *that* class satisfies -- a look-alike dataclass would fall through to ``assert_never``.
    from stamp.preprocessing.extractor import Extractor  # type: ignore[import-not-found]

    BOUND_TO_REFERENCE = True
except ImportError:
    BOUND_TO_REFERENCE = False

    @dataclass(frozen=True)
    class Extractor(Generic[ExtractorModel]):  # type: ignore[no-redef]
        """Stand-alone stand-in with the fields of stamp.preprocessing.extractor.Extractor (:18-28);
        only used when the ``stamp`` package is not installed."""

        _: KW_ONLY
        model: ExtractorModel
        transform: Callable[..., Tensor]
        identifier: str


def pil_to_u8_hwc(img) -> Tensor:
    """PIL RGB image (or HWC uint8 ndarray) -> uint8 [H, W, 3] tensor; no arithmetic."""
    arr = np.asarray(img.convert("RGB") if hasattr(img, "convert") else img, dtype=np.uint8)
    if arr.ndim != 3 or arr.shape[2] != 3:
        raise ValueError(f"expected an RGB image, got array of shape {arr.shape}")
    return torch.from_numpy(np.ascontiguousarray(arr))


def _resolve_weights(arch: VitArch, weights, hub_id: str | None, hub_kwargs: dict) -> Mapping[str, Tensor]:
    if isinstance(weights, Mapping):
        return weights
    if isinstance(weights, str) and weights == "random":
        return random_state_dict(arch)
    if isinstance(weights, (str, Path)):
        sd = torch.load(str(weights), map_location="cpu", weights_only=True)
        return sd.get("state_dict", sd)
    if weights is None and hub_id is not None:
        try:  # the reference's own source of weights (needs timm + network / HF cache)
            import timm

            return timm.create_model(hub_id, pretrained=True, **hub_kwargs).state_dict()
        except Exception as e:  # noqa: BLE001 - reported with instructions, never a silent fallback
            raise RuntimeError(
                f"could not obtain pretrained weights for {arch.name!r} via timm ({e}); pass "
                "weights=<state_dict | checkpoint path> or weights='random' for synthetic runs"
            ) from e
    raise ValueError("weights must be a state dict, a checkpoint path, 'random' or None")


def _make(arch: VitArch, identifier: str, weights, hub_id, hub_kwargs, max_batch: int) -> Extractor[TileEncoder]:
    sd = _resolve_weights(arch, weights, hub_id, hub_kwargs)
    return Extractor(model=TileEncoder(arch, sd, max_batch=max_batch), transform=pil_to_u8_hwc,
                     identifier=identifier)


def uni(weights=None, revision: str = "77ffbca1ee1cdcee6e87f6deebd2db8a5888c721",
        max_batch: int = 192) -> Extractor[TileEncoder]:
    """UNI ViT-L/16 (reference: src/stamp/preprocessing/extractor/uni.py:25-36)."""
    return _make(UNI_ARCH, "uni", weights, f"hf-hub:MahmoodLab/uni@{revision}",
                 dict(init_values=1e-5, dynamic_img_size=True), max_batch)


def virchow2(weights=None, max_batch: int = 96) -> Extractor[TileEncoder]:
    """Virchow2 ViT-H/14, class token only (reference: .../extractor/virchow2.py:24-54)."""
    hub_kwargs: dict = {}
    if weights is None:
        from timm.layers.mlp import SwiGLUPacked  # same constructor arguments as the reference

        hub_kwargs = dict(mlp_layer=SwiGLUPacked, act_layer=torch.nn.SiLU)
    return _make(VIRCHOW2_ARCH, "virchow2", weights, "hf-hub:paige-ai/Virchow2", hub_kwargs, max_batch)


def _virchow_hub_kwargs(weights) -> dict:
    if weights is not None:
        return {}
    from timm.layers.mlp import SwiGLUPacked  # same constructor arguments as the reference

    return dict(mlp_layer=SwiGLUPacked, act_layer=torch.nn.SiLU)


def virchow(weights=None, max_batch: int = 96) -> Extractor[TileEncoder]:
    """Virchow (v1) ViT-H/14, class token only, 1280 values (reference: .../extractor/virchow.py:24-57)."""
    return _make(VIRCHOW_ARCH, "virchow", weights, "hf-hub:paige-ai/Virchow", _virchow_hub_kwargs(weights), max_batch)


def virchow_full(weights=None, max_batch: int = 96) -> Extractor[TileEncoder]:
    """Virchow (v1) with the class token and the mean patch token concatenated, 2560 values (reference:
    .../extractor/virchow_full.py:24-62); the pooling runs on the GPU after the final norm over all tokens."""
    return _make(VIRCHOW_FULL_ARCH, "virchow-full", weights, "hf-hub:paige-ai/Virchow", _virchow_hub_kwargs(weights), max_batch)


def uni2(weights=None, max_batch: int = 96) -> Extractor[TileEncoder]:
    """UNI2-h ViT-H/14 with 8 register tokens (reference: .../extractor/uni2.py:16-46)."""
    hub_kwargs: dict = {}
    if weights is None:
        from timm.layers import SwiGLUPacked

        hub_kwargs = dict(img_size=224, patch_size=14, depth=24, num_heads=24, init_values=1e-5, embed_dim=1536,
                          mlp_ratio=2.66667 * 2, num_classes=0, no_embed_class=True, mlp_layer=SwiGLUPacked,
                          act_layer=torch.nn.SiLU, reg_tokens=8, dynamic_img_size=True)
    return _make(UNI2_ARCH, "uni2", weights, "hf-hub:MahmoodLab/UNI2-h", hub_kwargs, max_batch)


def h_optimus_0(weights=None, max_batch: int = 64) -> Extractor[TileEncoder]:
    """H-optimus-0 ViT-g/14 (reference: .../extractor/h_optimus_0.py:14-34; its Resize(224) is the identity
    on STAMP's 224 px tiles, mean / std are the model card's)."""
    return _make(H_OPTIMUS_ARCH, "h-optimus-0", weights, "hf-hub:bioptimus/H-optimus-0",
                 dict(init_values=1e-5, dynamic_img_size=False), max_batch)


def h_optimus_1(weights=None, max_batch: int = 64) -> Extractor[TileEncoder]:
    """H-optimus-1, same architecture and preprocessing (reference: .../extractor/h_optimus_1.py)."""
    return _make(H_OPTIMUS_ARCH, "h-optimus-1", weights, "hf-hub:bioptimus/H-optimus-1",
                 dict(init_values=1e-5, dynamic_img_size=False), max_batch)


def dino_bloom(weights, max_batch: int = 384) -> Extractor[TileEncoder]:
    """DinoBloom-S (reference: .../extractor/dinobloom.py:30-78).  ``weights``: the ``teacher`` state dict of
    DinoBloom-S.pth with the ``backbone.`` prefix removed (what the reference loads into the hub model), a path to a
    checkpoint in that layout, or "random"; the Zenodo download of the reference is the caller's business."""
    if isinstance(weights, Mapping) and any(k.startswith("backbone.") for k in weights):
        weights = {k.removeprefix("backbone."): v for k, v in weights.items() if "dino_head" not in k and "ibot_head" not in k}
    return _make(DINOBLOOM_ARCH, "dino-bloom", weights, None, {}, max_batch)


def gigapath(weights=None, max_batch: int = 64) -> Extractor[TileEncoder]:
    """Prov-GigaPath tile encoder, ViT-g/16 (reference: .../extractor/gigapath.py:14-35).  Its transform
    (Resize(256, BICUBIC) + CenterCrop(224) + ToTensor + Normalize) runs on the GPU: the resampling bit-exact with
    Pillow (stamp_b200/resize.py), the rest inside the patch kernel."""
    return _make(GIGAPATH_ARCH, "gigapath", weights, "hf_hub:prov-gigapath/prov-gigapath", {}, max_batch)


def extract_slide_features(extractor: Extractor, tiles_u8: Tensor, device: torch.device | str = "cuda",
                           batch_size: int = 192) -> Tensor:
    """The hot loop of ``extract_`` (src/stamp/preprocessing/__init__.py:322-327) for tiles already
    decoded to a host uint8 tensor [N, H, W, 3]: pinned double-buffered H2D copies on a side
    stream overlap the encoder; features come back as one fp16 host tensor [N, D].  CUDA tiles (decoded on the
    GPU) are consumed in place."""
    device = torch.device(device)
    model = extractor.model
    n = tiles_u8.shape[0]
    feats_host = torch.empty((n, model.arch.out_dim), dtype=torch.float16).pin_memory()
    if tiles_u8.is_cuda:
        # tiles already in HBM (decoded there: tiling.tiles_from_cache_file_gpu): nothing to stream in
        feats_dev = torch.empty((n, model.arch.out_dim), dtype=torch.float16, device=tiles_u8.device)
        for s in range(0, n, batch_size):
            feats_dev[s:s + batch_size] = model(tiles_u8[s:s + batch_size])
        feats_host.copy_(feats_dev, non_blocking=True)
        torch.cuda.current_stream(tiles_u8.device).synchronize()
        return feats_host
    if not tiles_u8.is_pinned():
        tiles_u8 = tiles_u8.pin_memory()
    copy_stream = torch.cuda.Stream(device=device)
    main = torch.cuda.current_stream(device)
    bufs = [torch.empty((batch_size, *tiles_u8.shape[1:]), dtype=torch.uint8, device=device) for _ in range(2)]
    ready = [torch.cuda.Event() for _ in range(2)]
    freed = [torch.cuda.Event() for _ in range(2)]
    starts = list(range(0, n, batch_size))

    def issue_copy(i: int) -> None:
        s = starts[i]
        b = min(batch_size, n - s)
        with torch.cuda.stream(copy_stream):
            if i >= 2:
                copy_stream.wait_event(freed[i % 2])
            bufs[i % 2][:b].copy_(tiles_u8[s:s + b], non_blocking=True)
            ready[i % 2].record(copy_stream)

    if starts:
        issue_copy(0)
    feats_dev = torch.empty((n, model.arch.out_dim), dtype=torch.float16, device=device)
    for i, s in enumerate(starts):
        if i + 1 < len(starts):
            issue_copy(i + 1)
        b = min(batch_size, n - s)
        main.wait_event(ready[i % 2])
        feats_dev[s:s + b] = model(bufs[i % 2][:b])
        freed[i % 2].record(main)
    feats_host.copy_(feats_dev, non_blocking=True)
    main.synchronize()
    return feats_host


def extract_to_feature_files(extractor: Extractor, slides, output_dir: str | Path, *, tile_size_um: float = 256.0,
                             tile_size_px: int = 224, device: torch.device | str = "cuda", batch_size: int = 192,
                             code_hash: str | None = None, rank: int = 0, world_size: int = 1) -> list[Path]:
    """The slide loop of ``extract_`` (src/stamp/preprocessing/__init__.py:276-366) downstream of tiling: ``slides``
    yields ``(relative_name, tiles_u8 [N, H, W, 3], coords_um [N, 2])``; every slide becomes
    ``output_dir/<identifier>[-<code_hash>]/<relative_name>.h5`` with the reference's datasets and attributes.
    Slides whose file already exists, and slides without tiles, are skipped like in the reference (:281-286,337-339).
    The file of slide i is written by a ``FeatureWriter`` thread while slide i+1 runs on the GPU; with
    ``world_size > 1`` rank r takes the slides r, r + world_size, ... (no data-path collective; pass an iterable
    already sharded by ``sharding.shard_lpt`` for size-aware assignment and leave ``world_size`` at 1)."""
    from .features import FeatureWriter

    out_dir = Path(output_dir) / (f"{extractor.identifier}-{code_hash}" if code_hash else str(extractor.identifier))
    written: list[Path] = []
    with FeatureWriter() as writer:
        for i, (name, tiles_u8, coords_um) in enumerate(slides):
            if i % world_size != rank:
                continue
            path = (out_dir / name).with_suffix(".h5")
            if path.exists() or len(tiles_u8) == 0:
                continue
            feats = extract_slide_features(extractor, tiles_u8, device=device, batch_size=batch_size)
            writer.submit(path, feats, coords_um, extractor=str(extractor.identifier), tile_size_um=tile_size_um,
                          tile_size_px=tile_size_px, code_hash=code_hash or "")
            written.append(path)
    return written


def _consume(ready, th, side, main, info, model, canny_cutoff, stain_normalize, feats_dev, kept_idx) -> None:
    """Consumer side of ``extract_cache_features``: GPU decode + tissue filter (+ stain normalisation) on the side
    stream, tile encoder on the main stream."""
    from . import jpeg
    from .tiling import has_enough_texture

    while True:
        item = ready.get()
        if item is None:
            return
        if isinstance(item, BaseException):
            raise item
        s, n, cd, qd = item
        with torch.cuda.stream(side):
            tiles = jpeg.decode_coefficients(info, cd, qd)
            cd.record_stream(side)
            qd.record_stream(side)
            if canny_cutoff is not None:
                keep = torch.nonzero(has_enough_texture(tiles, canny_cutoff)).squeeze(1)   # syncs the side stream only
                tiles = tiles[keep]
                kept_idx.append(keep.cpu() + s)
            else:
                kept_idx.append(torch.arange(s, s + n))
            if stain_normalize and tiles.shape[0]:
                from .macenko import macenko_normalize

                tiles = macenko_normalize(tiles.contiguous())
            done = torch.cuda.Event()
            done.record(side)
        if tiles.shape[0]:
            main.wait_event(done)
            tiles.record_stream(main)
            feats_dev.append(model(tiles))


def extract_cache_features(extractor: Extractor, cache_file_path: str | Path, device: torch.device | str = "cuda", *,
                           batch_size: int = 192, canny_cutoff: float | None = 0.02, max_workers: int = 8,
                           stain_normalize: bool = False) -> tuple[Tensor, Tensor, dict]:
    """A cached slide straight to features: what ``extract_`` does for a slide whose tiles are in the JPEG tile cache
    (``_tiles_from_cache_file`` src/stamp/preprocessing/tiling.py:380-406 feeding the loop at
    preprocessing/__init__.py:306-327), as a three-stage pipeline that keeps the tile encoder busy:

      host thread   zip -> Huffman decode of batch i+1 (thread pool, GIL released) -> pinned staging -> H2D
      side stream   inverse DCT / up-sampling / colour of batch i+1, Canny tissue filter (``canny_cutoff``; the
                    reference applies it before caching, ``None`` skips it), Macenko stain normalisation fitted on
                    the batch's kept tiles (``stain_normalize``; the new stage of BASELINE configs[2])
      main stream   tile encoder on the kept tiles of batch i

    Returns (features fp16 [n_kept, D] on the host, coordinates [n_kept, 2] in microns, tiler parameters); pixels,
    decisions and therefore features are those of the Pillow + OpenCV path."""
    import json
    import queue
    import re
    import threading
    from zipfile import ZipFile

    from . import jpeg
    from .tiling import has_enough_texture

    device = torch.device(device)
    model = extractor.model
    zf = ZipFile(Path(cache_file_path), "r")
    params = json.loads(zf.read("tiler_params.json").decode())
    ext = params.get("tile_ext", "jpg")
    if ext.lower() not in ("jpg", "jpeg"):
        zf.close()
        raise ValueError(f"tile cache holds {ext!r} tiles; the GPU decoder reads JPEG caches only")
    pat = re.compile(rf"tile_\((\d+\.\d+), (\d+\.\d+)\)\.{re.escape(ext)}")
    names, coords = [], []
    for name in zf.namelist():
        m = pat.fullmatch(name)
        if m is not None:
            names.append(name)
            coords.append((float(m.group(1)), float(m.group(2))))
    coords_t = torch.tensor(coords, dtype=torch.float32).reshape(-1, 2)
    if not names:
        zf.close()
        return torch.empty((0, model.arch.out_dim), dtype=torch.float16), coords_t, params
    info = jpeg.read_header(zf.read(names[0]))
    n_coef = jpeg.coef_count(info)
    starts = list(range(0, len(names), batch_size))
    side = torch.cuda.Stream(device=device)
    main = torch.cuda.current_stream(device)
    stage = [(torch.empty((batch_size, n_coef), dtype=torch.int16).pin_memory(),
              torch.empty((batch_size, 3, 64), dtype=torch.int16).pin_memory()) for _ in range(2)]
    copied: list[torch.cuda.Event | None] = [None, None]
    ready: queue.Queue = queue.Queue(maxsize=2)

    stop = threading.Event()

    def hand_over(item) -> bool:
        while not stop.is_set():
            try:
                ready.put(item, timeout=0.2)
                return True
            except queue.Full:
                continue
        return False

    def producer() -> None:
        try:
            for i, s in enumerate(starts):
                if stop.is_set():
                    return
                part = [zf.read(n) for n in names[s:s + batch_size]]   # the zip is read here, behind the encoder
                if copied[i % 2] is not None:
                    copied[i % 2].synchronize()              # the staging set's previous H2D copy has finished
                _, coef, quant = jpeg.entropy_decode(part, max_workers=max_workers, out=stage[i % 2])
                with torch.cuda.stream(side):
                    cd, qd = coef.to(device, non_blocking=True), quant.to(device, non_blocking=True)
                    ev = torch.cuda.Event()
                    ev.record(side)
                copied[i % 2] = ev
                if not hand_over((s, len(part), cd, qd)):
                    return
            hand_over(None)
        except BaseException as e:  # noqa: BLE001 - handed to the consumer
            hand_over(e)
        finally:
            zf.close()

    th = threading.Thread(target=producer, name="stamp-b200-jpeg-feed", daemon=True)
    th.start()
    feats_dev, kept_idx = [], []
    try:
        _consume(ready, th, side, main, info, model, canny_cutoff, stain_normalize, feats_dev, kept_idx)
    finally:
        stop.set()        # a consumer error must not leave the feed thread blocked on a full queue
        th.join()
    idx = torch.cat(kept_idx) if kept_idx else torch.empty(0, dtype=torch.long)
    if feats_dev:
        feats = torch.cat(feats_dev)
        host = torch.empty(feats.shape, dtype=torch.float16).pin_memory()
        host.copy_(feats, non_blocking=True)
        main.synchronize()
    else:
        host = torch.empty((0, model.arch.out_dim), dtype=torch.float16)
    return host, coords_t[idx], params
