"""Feature files: what ``extract_`` writes per slide and what the bag loaders / encoders read back.

Mirrors the reference's behaviour at its own call sites, on ``h5lite`` (h5py is not in this image):

* ``write_tile_features``  <- src/stamp/preprocessing/__init__.py:342-366 (datasets ``coords`` [N,2] in
  microns and ``feats`` [N,D] fp16; attributes ``stamp_version``, ``extractor``, ``unit``,
  ``tile_size_um``, ``tile_size_px``, ``code_hash``, ``feat_type``; written under a temporary name in
  the output directory, then renamed, so that a crash never leaves a half-written ``.h5``);
* ``write_encoded_features`` <- src/stamp/encoding/encoder/__init__.py:203-229 (``feats`` + ``version``,
  ``encoder``, ``precision``, ``stamp_version``, ``code_hash``, ``feat_type``);
* ``get_coords`` <- src/stamp/modeling/data.py:741-808 (the three coordinate conventions: v2 files
  with ``tile_size`` + ``unit == "um"``, newer files with ``tile_size_um``, historic files whose
  coordinates are 224 px strides of 256 um tiles; multiplex files without ``coords``);
* ``read_bag`` <- ``BagDataset.__getitem__`` (src/stamp/modeling/data.py:603-655: ``feats`` or
  ``patch_embeddings``, coordinates through ``get_coords``);
* ``FeatureWriter``: the write runs on a background thread so that encoding the next slide does not wait
  for the file system;
* ``load_cohort_to_device``: every feature file of a cohort is read once, straight into one pinned staging
  buffer and copied to one HBM-resident fp16 tensor; bags are views into it, so the per-step host
  up-cast / collate / H2D copy of the reference's DataLoader disappears.
"""

from __future__ import annotations

import os
import queue
import tempfile
import threading
from collections.abc import Mapping, Sequence
from dataclasses import dataclass
from pathlib import Path
from typing import Any

import numpy as np
import torch
from torch import Tensor

from . import h5lite

try:
    from stamp import __version__ as STAMP_VERSION  # type: ignore[import-not-found]
except Exception:  # noqa: BLE001 - the package is optional here
    STAMP_VERSION = "2.4.0"


@dataclass(frozen=True)
class CoordsInfo:
    """src/stamp/modeling/data.py:722-738."""

    coords_um: np.ndarray
    tile_size_um: float
    tile_size_px: int | None = None

    @property
    def mpp(self) -> float:
        if not self.tile_size_px:
            raise RuntimeError("tile size in pixels is not available. Please reextract them using `stamp preprocess`.")
        return self.tile_size_um / self.tile_size_px


def _atomic_h5(path: Path, fill) -> None:  # noqa: ANN001
    path.parent.mkdir(parents=True, exist_ok=True)
    fd, tmp = tempfile.mkstemp(dir=path.parent)
    try:
        with os.fdopen(fd, "wb") as fp, h5lite.File(fp, "w") as h5:
            fill(h5)
        os.replace(tmp, path)
    except BaseException:
        Path(tmp).unlink(missing_ok=True)
        raise


def _as_numpy(x: Any) -> np.ndarray:
    return x.detach().cpu().numpy() if isinstance(x, Tensor) else np.asarray(x)


def write_tile_features(path: str | os.PathLike, feats: Any, coords_um: Any, *, extractor: str, tile_size_um: float,
                        tile_size_px: int, code_hash: str = "", stamp_version: str = STAMP_VERSION) -> None:
    feats, coords_um = _as_numpy(feats), _as_numpy(coords_um)
    if feats.ndim != 2 or coords_um.shape != (feats.shape[0], 2):
        raise ValueError(f"feats {feats.shape} / coords {coords_um.shape}: expected [N, D] and [N, 2]")

    def fill(h5: h5lite.File) -> None:
        h5["coords"] = coords_um
        h5["feats"] = feats
        h5.attrs["stamp_version"] = stamp_version
        h5.attrs["extractor"] = str(extractor)
        h5.attrs["unit"] = "um"
        h5.attrs["tile_size_um"] = float(tile_size_um)
        h5.attrs["tile_size_px"] = int(tile_size_px)
        h5.attrs["code_hash"] = code_hash
        h5.attrs["feat_type"] = "tile"

    _atomic_h5(Path(path), fill)


def write_encoded_features(path: str | os.PathLike, feats: Any, *, encoder: str, precision: torch.dtype,
                           feat_type: str, code_hash: str = "", stamp_version: str = STAMP_VERSION) -> None:
    feats = _as_numpy(feats)

    def fill(h5: h5lite.File) -> None:
        h5["feats"] = feats
        h5.attrs["version"] = stamp_version
        h5.attrs["encoder"] = str(encoder)
        h5.attrs["precision"] = str(precision)
        h5.attrs["stamp_version"] = stamp_version
        h5.attrs["code_hash"] = code_hash
        h5.attrs["feat_type"] = feat_type

    _atomic_h5(Path(path), fill)


def get_stride(coords: np.ndarray) -> float:
    """Minimum step between any two distinct x or y coordinates (src/stamp/modeling/data.py:1150-1161)."""
    xs, ys = np.unique(coords[:, 0]), np.unique(coords[:, 1])
    return float(min(np.diff(xs).min(), np.diff(ys).min()))


def _version_tuple(v: str) -> tuple[int, ...]:
    out = []
    for part in str(v).split("+")[0].split("."):
        digits = "".join(ch for ch in part if ch.isdigit())
        out.append(int(digits) if digits else 0)
    return tuple(out)


def get_coords(h5: h5lite.File, stamp_version: str = STAMP_VERSION) -> CoordsInfo:
    attrs: Mapping[str, Any] = h5.attrs
    if "coords" not in h5:
        n = h5["patch_embeddings"].shape[0]
        return CoordsInfo(np.stack([np.arange(n), np.zeros(n)], axis=1).astype(np.float32), 0.0, 0)
    coords = np.asarray(h5["coords"][:])
    tile_size_um: float | None = None
    tile_size_px: int | None = None
    coords_um: np.ndarray | None = None
    if (tile_size := attrs.get("tile_size", None)) and attrs.get("unit", None) == "um":
        tile_size_um, coords_um = float(tile_size), coords
    elif tile_size := attrs.get("tile_size_um", None):
        tile_size_um, coords_um = float(tile_size), coords
    elif round(float(attrs["tile_size"]) if "tile_size" in attrs else get_stride(coords.astype(np.float32))) == 224:
        tile_size_um, tile_size_px, coords_um = 256.0, 224, coords / 224 * 256
    if (v := attrs.get("stamp_version")) and _version_tuple(v) > _version_tuple(stamp_version):
        raise RuntimeError("features were extracted with a newer version of stamp, please update your stamp to at "
                           f"least version {v}.")
    if not tile_size_px and "tile_size_px" in attrs:
        tile_size_px = int(attrs["tile_size_px"])
    if not tile_size_um or coords_um is None:
        raise RuntimeError("unable to infer coordinates from feature file. Please reextract them using "
                           "`stamp preprocess`.")
    return CoordsInfo(coords_um, tile_size_um, tile_size_px)


def _feats_dataset(h5: h5lite.File) -> h5lite.Dataset:
    ds = h5["feats"] if "feats" in h5 else h5["patch_embeddings"]
    if not isinstance(ds, h5lite.Dataset):
        raise RuntimeError(f"expected 'feats' to be an HDF5 dataset but got {type(ds)}")
    return ds


def read_bag(paths: Sequence[str | os.PathLike]) -> tuple[Tensor, Tensor]:
    """All feature files of one patient -> (feats [N, D] fp32, coords_um [N, 2] fp32), concatenated in order."""
    feats, coords = [], []
    for p in paths:
        with h5lite.File(p, "r") as h5:
            feats.append(torch.from_numpy(np.asarray(_feats_dataset(h5)[()])))
            coords.append(torch.from_numpy(np.asarray(get_coords(h5).coords_um)))
    return torch.concat(feats).float(), torch.concat(coords).float()


def read_tile_features(path: str | os.PathLike) -> tuple[np.ndarray, CoordsInfo, str]:
    """What ``Encoder._read_h5`` needs (src/stamp/encoding/encoder/__init__.py:180-201): features as stored, the
    coordinates, the extractor's name with a trailing code hash stripped."""
    path = os.fspath(path)
    if not os.path.exists(path):
        raise FileNotFoundError(f"File does not exist: {path}")
    if not path.endswith(".h5"):
        raise ValueError(f"File is not of type .h5: {os.path.basename(path)}")
    with h5lite.File(path, "r") as h5:
        feats = np.asarray(h5["feats"][()])
        coords = get_coords(h5)
        extractor = h5.attrs.get("extractor", "")
    if extractor == "":
        raise ValueError(f"Feature file does not have extractor's name in the metadata: {os.path.basename(path)}")
    return feats, coords, resolve_extractor_name(extractor)


def resolve_extractor_name(name: str) -> str:
    """``uni-1a2b3c4d`` -> ``uni`` when the suffix is a hash (src/stamp/encoding/encoder/__init__.py:235-251)."""
    if not name:
        raise ValueError("Empty extractor name")
    name = str(name).strip()
    if "-" not in name:
        return name
    base, suffix = name.rsplit("-", 1)
    if len(suffix) >= 6 and all(c in "0123456789abcdefABCDEF" for c in suffix):
        return base
    return name


class FeatureWriter:
    """Background writer: ``submit`` returns as soon as the fp16 features sit in host memory; the file is laid out,
    written under a temporary name and renamed by one worker thread.  ``close`` (or leaving the ``with`` block) waits
    for the queue to drain and re-raises the first error."""

    def __init__(self, max_pending: int = 4) -> None:
        self._q: queue.Queue = queue.Queue(maxsize=max_pending)
        self._error: BaseException | None = None
        self.written: list[Path] = []
        self._thread = threading.Thread(target=self._run, name="stamp-b200-h5-writer", daemon=True)
        self._thread.start()

    def _run(self) -> None:
        while True:
            job = self._q.get()
            try:
                if job is None:
                    return
                if self._error is None:
                    path, kwargs = job
                    write_tile_features(path, **kwargs)
                    self.written.append(Path(path))
            except BaseException as e:  # noqa: BLE001 - reported by close()
                self._error = e
            finally:
                self._q.task_done()

    def submit(self, path: str | os.PathLike, feats: Tensor | np.ndarray, coords_um: Tensor | np.ndarray,
               **attrs: Any) -> None:
        if self._error is not None:
            self.close()
        self._q.put((Path(path), dict(feats=feats, coords_um=coords_um, **attrs)))

    def close(self) -> None:
        if self._thread.is_alive():
            self._q.put(None)
            self._thread.join()
        if self._error is not None:
            err, self._error = self._error, None
            raise err

    def __enter__(self) -> "FeatureWriter":
        return self

    def __exit__(self, exc_type, *_exc) -> None:  # noqa: ANN001
        if exc_type is None:
            self.close()
        elif self._thread.is_alive():
            self._q.put(None)
            self._thread.join()


@dataclass
class DeviceCohort:
    """All tiles of a cohort in one device tensor; ``bag(i)`` returns views, nothing is copied."""

    feats: Tensor            # [total, D] fp16 on the device
    coords: Tensor           # [total, 2] fp32 on the device
    offsets: list[int]       # len = n_bags + 1
    ids: list[str]

    def __len__(self) -> int:
        return len(self.ids)

    def bag(self, i: int) -> tuple[Tensor, Tensor]:
        return self.feats[self.offsets[i]:self.offsets[i + 1]], self.coords[self.offsets[i]:self.offsets[i + 1]]


def load_cohort_to_device(bags: Mapping[str, Sequence[str | os.PathLike]], device: torch.device | str,
                          chunk_bytes: int = 256 << 20) -> DeviceCohort:
    """``bags``: patient id -> that patient's feature files (src/stamp/modeling/data.py:536-561 builds the same
    mapping from the slide table).  Two passes: headers only (shapes), then each ``feats`` dataset is read with
    ``read_direct`` into a slice of a pinned staging buffer (two buffers of ``chunk_bytes``, so disk reads overlap
    the asynchronous copies) and lands in its slot of the device tensor; fp32 files are narrowed to fp16 on the
    device, coordinates are converted to microns by ``get_coords`` on the host (2 floats per tile)."""
    device = torch.device(device)
    ids = list(bags)
    shapes: list[list[tuple[int, int]]] = []
    dim = None
    for pid in ids:
        per = []
        for p in bags[pid]:
            with h5lite.File(p, "r") as h5:
                ds = _feats_dataset(h5)
                if len(ds.shape) != 2:
                    raise RuntimeError(f"{p}: expected a [tiles, dim] feature matrix, got shape {ds.shape}")
                if dim is None:
                    dim = ds.shape[1]
                elif ds.shape[1] != dim:
                    raise RuntimeError(f"{p}: feature dimension {ds.shape[1]} differs from {dim}")
                per.append((ds.shape[0], ds.dtype.itemsize))
        shapes.append(per)
    if dim is None:
        raise ValueError("empty cohort")
    counts = [sum(n for n, _ in per) for per in shapes]
    offsets = [0]
    for c in counts:
        offsets.append(offsets[-1] + c)
    total = offsets[-1]
    feats_dev = torch.empty((total, dim), dtype=torch.float16, device=device)
    coords_host = torch.empty((total, 2), dtype=torch.float32)
    pinned = device.type == "cuda"
    staging = [torch.empty(chunk_bytes, dtype=torch.uint8, pin_memory=pinned) for _ in range(2)]
    events: list[Any] = [None, None]
    which, used = 0, 0
    row = 0
    for pid in ids:
        for p in bags[pid]:
            with h5lite.File(p, "r") as h5:
                ds = _feats_dataset(h5)
                n, item = ds.shape[0], ds.dtype.itemsize
                coords_host[row:row + n] = torch.from_numpy(np.asarray(get_coords(h5).coords_um, dtype=np.float32))
                nbytes = n * dim * item
                if nbytes > chunk_bytes:
                    raise RuntimeError(f"{p}: {nbytes} bytes exceed the staging buffer; raise chunk_bytes")
                used = (used + 15) & ~15
                if used + nbytes > chunk_bytes:
                    which, used = which ^ 1, 0
                    if events[which] is not None:
                        events[which].synchronize()
                buf = staging[which][used:used + nbytes]
                host = buf.numpy().view(ds.dtype).reshape(n, dim)
                ds.read_direct(host)
                src = torch.from_numpy(host)
                feats_dev[row:row + n].copy_(src, non_blocking=pinned)
                if pinned:
                    events[which] = torch.cuda.Event()
                    events[which].record()
                used += nbytes
                row += n
    coords_dev = coords_host.to(device)
    if pinned:
        torch.cuda.current_stream(device).synchronize()
    return DeviceCohort(feats_dev, coords_dev, offsets, ids)
