"""Tile-level tissue filter of the tiling stage on the GPU.

Mirrors ``_has_enough_texture`` (src/stamp/preprocessing/tiling.py:279-291), which the reference runs per
tile on the CPU inside its tiling workers (``tile.convert("L")`` -> ``cv2.Canny(gray, 40, 100)`` ->
``edges.mean() / 255 >= cutoff``): here a whole batch of decoded uint8 tiles is scored by one kernel launch
(``stamp_tile_texture_u8``), bit-exact with Pillow + OpenCV, so exactly the same tiles survive the filter.
No CPU fallback.
"""

from __future__ import annotations

import ctypes as C
import json
import re
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path
from zipfile import ZipFile

import numpy as np
import torch
from torch import Tensor

from . import _lib

CANNY_LOW, CANNY_HIGH = 40, 100   # "hardcoded thresholds", tiling.py:284-285


def _bind() -> C.CDLL:
    lib = _lib.load()
    if not getattr(lib, "_texture_bound", False):
        lib.stamp_tile_texture_u8.restype = C.c_int
        lib.stamp_tile_texture_u8.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                              C.c_void_p, C.c_void_p]
        lib._texture_bound = True
    return lib


def canny_edge_counts(tiles: Tensor, *, return_edges: bool = False):
    """uint8 ``[B, H, W, 3]`` CUDA tiles -> int32 ``[B]`` Canny edge-pixel counts (and the uint8 0/255 edge maps)."""
    if not tiles.is_cuda:
        raise RuntimeError("canny_edge_counts runs on a CUDA device only (no CPU fallback)")
    if tiles.dtype != torch.uint8 or tiles.dim() != 4 or tiles.shape[-1] != 3 or not tiles.is_contiguous():
        raise TypeError("tiles must be a contiguous uint8 [B,H,W,3] tensor")
    B, H, W, _ = tiles.shape
    counts = torch.empty(B, dtype=torch.int32, device=tiles.device)
    edges = torch.empty((B, H, W), dtype=torch.uint8, device=tiles.device) if return_edges else None
    if B > 0:
        code = _bind().stamp_tile_texture_u8(tiles.data_ptr(), B, H, W, CANNY_LOW, CANNY_HIGH, counts.data_ptr(),
                                             None if edges is None else edges.data_ptr(),
                                             torch.cuda.current_stream().cuda_stream)
        _lib.check(code, "stamp_tile_texture_u8")
    return (counts, edges) if return_edges else counts


def edge_scores(tiles: Tensor) -> Tensor:
    """``np.array(edges).mean() / 255`` per tile, in the same double arithmetic as NumPy (fp64 ``[B]``)."""
    counts = canny_edge_counts(tiles)
    # tensor / tensor is an IEEE division; tensor / python-scalar may be turned into a multiplication by the
    # reciprocal, which differs from NumPy in the last bit (and the result is compared with a cutoff)
    hw = torch.tensor(float(tiles.shape[1] * tiles.shape[2]), dtype=torch.float64, device=tiles.device)
    c255 = torch.tensor(255.0, dtype=torch.float64, device=tiles.device)
    return (counts.double() * c255 / hw) / c255


def has_enough_texture(tiles: Tensor, cutoff: float) -> Tensor:
    """``_has_enough_texture`` for a batch: bool ``[B]``, True where the tile is likely to contain tissue."""
    return edge_scores(tiles) >= cutoff


def foreground_coords(slide_dimensions: tuple[int, int], tile_size_slide_px: int, thumbnail,
                      brightness_cutoff: int | None) -> list[tuple[int, int]]:
    """Brightness rejection of ``_foreground_coords`` (src/stamp/preprocessing/tiling.py:250-277): level-0 pixel
    coordinates (x, y) of the tiles whose thumbnail pixel is darker than ``brightness_cutoff`` (all tiles for
    ``None``), in the reference's row-major order.  ``thumbnail`` is what ``slide.get_thumbnail(2 * grid)`` returns
    (a PIL image): it is resized to one pixel per tile and converted to 32-bit grayscale exactly as there.  A few
    hundred pixels per slide -- host work, nothing for the GPU."""
    w, h = int(slide_dimensions[0]), int(slide_dimensions[1])
    grid = (-(-w // tile_size_slide_px), -(-h // tile_size_slide_px))
    gray = np.array(thumbnail.resize(grid).convert("I"))
    fg = gray < brightness_cutoff if brightness_cutoff is not None else np.ones_like(gray, dtype=bool)
    ys, xs = np.nonzero(fg[: len(range(0, h, tile_size_slide_px)), : len(range(0, w, tile_size_slide_px))])
    return [(int(x) * tile_size_slide_px, int(y) * tile_size_slide_px) for y, x in zip(ys, xs)]


def tiles_from_cache_file(cache_file_path: str | Path, *, max_workers: int = 8,
                          pin_memory: bool = True) -> tuple[Tensor, Tensor, dict]:
    """Reads a STAMP tile cache (``_tiles_from_cache_file``, src/stamp/preprocessing/tiling.py:380-406: a zip with
    ``tiler_params.json`` and one ``tile_(x_um, y_um).<ext>`` image per tile) into ONE uint8 ``[N, H, W, 3]`` host
    tensor (pinned, so ``extract_slide_features`` can stream it to the GPU) plus the ``[N, 2]`` coordinates in
    microns and the tiler parameters.  Decoding is Pillow's, like the reference's (bit-identical pixels), on a
    thread pool; tiles come back in zip order, which is the order the reference iterates them in."""
    from PIL import Image

    path = Path(cache_file_path)
    with ZipFile(path, "r") as zf:
        params = json.loads(zf.read("tiler_params.json").decode())
        ext = params.get("tile_ext", "jpg")          # "jpg as default for backwards compatibility"
        pat = re.compile(rf"tile_\((\d+\.\d+), (\d+\.\d+)\)\.{re.escape(ext)}")
        names, coords = [], []
        for name in zf.namelist():
            m = pat.fullmatch(name)
            if m is not None:
                names.append(name)
                coords.append((float(m.group(1)), float(m.group(2))))
        blobs = [zf.read(n) for n in names]
    if not names:
        return torch.empty((0, 0, 0, 3), dtype=torch.uint8), torch.empty((0, 2)), params

    def decode(blob: bytes) -> np.ndarray:
        import io

        with Image.open(io.BytesIO(blob)) as im:
            return np.asarray(im.convert("RGB"), dtype=np.uint8)

    first = decode(blobs[0])
    out = torch.empty((len(blobs), *first.shape), dtype=torch.uint8)
    if pin_memory and torch.cuda.is_available():
        out = out.pin_memory()
    dst = out.numpy()
    dst[0] = first

    def work(i: int) -> None:
        a = decode(blobs[i])
        if a.shape != first.shape:
            raise ValueError(f"tile {names[i]!r} has shape {a.shape}, expected {first.shape}")
        dst[i] = a

    with ThreadPoolExecutor(max_workers=max_workers) as ex:
        list(ex.map(work, range(1, len(blobs))))
    return out, torch.tensor(coords, dtype=torch.float32), params


def tiles_from_cache_file_gpu(cache_file_path: str | Path, device: torch.device | str = "cuda", *, max_workers: int = 8,
                              batch: int = 512) -> tuple[Tensor, Tensor, dict]:
    """``tiles_from_cache_file`` with the tiles born on the GPU: for a JPEG cache (the reference's default,
    ``tile_ext = "jpg"``) the host only Huffman-decodes (``stamp_b200.jpeg.entropy_decode``, a thread pool over the
    tiles), the inverse DCT / chroma up-sampling / colour conversion run on ``device`` in batches of ``batch`` tiles.
    Returns the same pixels as the Pillow path (bit-exact), as a uint8 ``[N, H, W, 3]`` CUDA tensor.  Caches in another
    image format, or JPEG variants Pillow does not write by default (progressive, CMYK), raise -- use
    ``tiles_from_cache_file`` for those."""
    from . import jpeg

    device = torch.device(device)
    path = Path(cache_file_path)
    with ZipFile(path, "r") as zf:
        params = json.loads(zf.read("tiler_params.json").decode())
        ext = params.get("tile_ext", "jpg")
        if ext.lower() not in ("jpg", "jpeg"):
            raise ValueError(f"tile cache holds {ext!r} tiles; the GPU decoder reads JPEG caches only")
        pat = re.compile(rf"tile_\((\d+\.\d+), (\d+\.\d+)\)\.{re.escape(ext)}")
        names, coords = [], []
        for name in zf.namelist():
            m = pat.fullmatch(name)
            if m is not None:
                names.append(name)
                coords.append((float(m.group(1)), float(m.group(2))))
        blobs = [zf.read(n) for n in names]
    if not names:
        return torch.empty((0, 0, 0, 3), dtype=torch.uint8, device=device), torch.empty((0, 2)), params
    info = jpeg.read_header(blobs[0])
    out = torch.empty((len(blobs), info.height, info.width, 3), dtype=torch.uint8, device=device)
    batch = min(batch, len(blobs))
    n = jpeg.coef_count(info)
    # two pinned staging sets: the pool Huffman-decodes batch i+1 while batch i crosses PCIe and runs its kernels
    stage = [(torch.empty((batch, n), dtype=torch.int16).pin_memory(), torch.empty((batch, 3, 64), dtype=torch.int16).pin_memory())
             for _ in range(2)]
    copied: list[torch.cuda.Event | None] = [None, None]
    for i, s in enumerate(range(0, len(blobs), batch)):
        part = blobs[s:s + batch]
        if copied[i % 2] is not None:
            copied[i % 2].synchronize()
        _, coef, quant = jpeg.entropy_decode(part, max_workers=max_workers, out=stage[i % 2])
        cd, qd = coef.to(device, non_blocking=True), quant.to(device, non_blocking=True)
        copied[i % 2] = torch.cuda.Event()
        copied[i % 2].record()
        jpeg.decode_coefficients(info, cd, qd, out=out[s:s + len(part)])
    return out, torch.tensor(coords, dtype=torch.float32), params
