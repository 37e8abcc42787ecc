"""Slide-level encoders backed by the B200 pooling kernels: CHIEF (gated-attention pooling) and
EAGLE (CHIEF attention -> top-25 tiles -> mean of the aggregation features).

Mirrors the reference's ``Encoder`` contract (src/stamp/encoding/encoder/__init__.py:29-171):
``_generate_slide_embedding(feats [N, D], device, **kw) -> np.ndarray [D']``.  Model arithmetic:
src/stamp/encoding/encoder/chief.py:27-89,255-275 and eagle.py:92-120.
"""

from __future__ import annotations

import ctypes as C
from abc import ABC, abstractmethod
from collections.abc import Mapping
from enum import StrEnum

import numpy as np
import torch
from torch import Tensor

from . import _lib

try:
    # The reference's own base class and enums whenever STAMP is importable: ``init_slide_encoder_`` /
    # ``init_patient_encoder_`` dispatch on ``case Encoder():`` (src/stamp/encoding/__init__.py:72-73,158-159),
    # which only a subclass of *that* class satisfies, and the subclass inherits ``encode_slides_`` /
    # ``encode_patients_`` (the h5 walk + atomic save, encoder/__init__.py:42-171,203-229) unchanged.
    from stamp.encoding.config import EncoderName  # type: ignore[import-not-found]
    from stamp.encoding.encoder import Encoder  # type: ignore[import-not-found]
    from stamp.preprocessing.config import ExtractorName  # type: ignore[import-not-found]

    BOUND_TO_REFERENCE = True
except ImportError:
    BOUND_TO_REFERENCE = False

    class EncoderName(StrEnum):  # type: ignore[no-redef]
        """The values of stamp.encoding.config.EncoderName this package provides."""

        EAGLE = "eagle"
        CHIEF_CTRANSPATH = "chief"

    class ExtractorName(StrEnum):  # type: ignore[no-redef]
        """The values of stamp.preprocessing.config.ExtractorName the encoders below name."""

        CTRANSPATH = "ctranspath"
        CHIEF_CTRANSPATH = "chief-ctranspath"
        VIRCHOW2 = "virchow2"

    class Encoder(ABC):  # type: ignore[no-redef]
        """Stand-alone stand-in for stamp.encoding.encoder.Encoder (same constructor, abstract methods and
        feature-file walk, encoder/__init__.py:29-229); only used when the ``stamp`` package is not installed.
        Feature files go through ``stamp_b200.features`` (h5lite) instead of h5py; the output folder carries no code
        hash of the reference's sources (``generate_hash`` appends this module's own)."""

        def __init__(self, model, identifier, precision: torch.dtype, required_extractors: list) -> None:
            self.model, self.identifier = model, identifier
            self.precision, self.required_extractors = precision, required_extractors

        @abstractmethod
        def _generate_slide_embedding(self, feats: Tensor, device, **kwargs) -> np.ndarray: ...

        @abstractmethod
        def _generate_patient_embedding(self, feats_list: list[Tensor], device, **kwargs) -> np.ndarray: ...

        def _encode_dir(self, output_dir, level: str, generate_hash: bool):
            from pathlib import Path

            name = f"{self.identifier}-{level}"
            if generate_hash:
                name += f"-{_code_hash()}"
            out = Path(output_dir) / name
            out.mkdir(parents=True, exist_ok=True)
            return out

        def _validate_and_read_features(self, h5_path: str):
            from . import features

            feats, coords, extractor = features.read_tile_features(h5_path)
            if extractor not in self.required_extractors:
                raise ValueError(f"Features must be extracted with one of {self.required_extractors}. "
                                 f"Features located in {h5_path} are extracted with {extractor}")
            # features stay in their storage dtype (fp16) on the host: the subclasses move them to the device first
            # and widen there (a host fp16 -> fp32 pass over a 10 k-tile slide costs more than the pooling kernels)
            return torch.from_numpy(feats), coords

        def _save_features_(self, output_path, feats: np.ndarray, feat_type: str) -> None:
            from . import features

            features.write_encoded_features(output_path, feats, encoder=str(self.identifier), precision=self.precision,
                                            feat_type=feat_type, code_hash=_code_hash())

        def encode_slides_(self, output_dir, feat_dir, device, generate_hash: bool = False, **kwargs) -> None:
            """encoder/__init__.py:42-93: every ``*.h5`` under ``feat_dir`` -> one slide embedding file under
            ``output_dir/<identifier>-slide``, keeping the relative folder structure; existing outputs are skipped,
            files from the wrong extractor are reported and skipped."""
            import logging
            from pathlib import Path

            encode_dir = self._encode_dir(output_dir, "slide", generate_hash)
            self.model.to(device).eval()
            for h5_path in sorted(Path(feat_dir).rglob("*.h5")):
                output_path = (encode_dir / h5_path.relative_to(feat_dir)).with_suffix(".h5")
                if output_path.exists():
                    continue
                try:
                    feats, coords = self._validate_and_read_features(str(h5_path))
                except ValueError as e:
                    logging.getLogger("stamp").warning(str(e))
                    continue
                emb = self._generate_slide_embedding(feats, device, coords=coords)
                self._save_features_(output_path, emb, "slide")

        def encode_patients_(self, output_dir, feat_dir, slide_table_path, patient_label: str, filename_label: str,
                             device, generate_hash: bool = False, **kwargs) -> None:
            """encoder/__init__.py:95-156: the slide table groups feature files by patient; one embedding file per
            patient under ``output_dir/<identifier>-pat``."""
            import os
            from pathlib import Path

            import pandas as pd

            encode_dir = self._encode_dir(output_dir, "pat", generate_hash)
            self.model.to(device).eval()
            table_path = Path(slide_table_path)
            table = pd.read_excel(table_path) if table_path.suffix == ".xlsx" else pd.read_csv(table_path)
            for patient_id, group in table.groupby(patient_label):
                output_path = (encode_dir / str(patient_id)).with_suffix(".h5")
                if output_path.exists():
                    continue
                feats_list = [self._validate_and_read_features(os.path.join(feat_dir, row[filename_label]))[0]
                              for _, row in group.iterrows()]
                if not feats_list:
                    continue
                emb = self._generate_patient_embedding(feats_list, device, **kwargs)
                self._save_features_(output_path, emb, "patient")

    def _code_hash() -> str:
        import hashlib
        from pathlib import Path

        return hashlib.sha256(Path(__file__).read_bytes()).hexdigest()[:8]

_EagleBase = Encoder
if BOUND_TO_REFERENCE:
    try:  # EAGLE's feature-file pairing (ctranspath + virchow2 h5 per slide, eagle.py:40-89,136-300) is host I/O
        from stamp.encoding.encoder.eagle import Eagle as _EagleBase  # type: ignore[import-not-found,no-redef]
    except ImportError:  # its module imports gdown etc. at the top
        pass


class StampGatedAttnWeights(C.Structure):
    _fields_ = [("fc_w_hi", C.c_void_p), ("fc_w_lo", C.c_void_p), ("fc_b", C.c_void_p),
                ("ab_w_hi", C.c_void_p), ("ab_w_lo", C.c_void_p), ("ab_b", C.c_void_p),
                ("c_w", C.c_void_p), ("c_b", C.c_float)]


def _bind() -> C.CDLL:
    lib = _lib.load()
    if not getattr(lib, "_pool_bound", False):
        lib.stamp_gated_attn_pool_workspace_bytes.restype = C.c_size_t
        lib.stamp_gated_attn_pool_workspace_bytes.argtypes = [C.c_int] * 4
        lib.stamp_gated_attn_pool.restype = C.c_int
        lib.stamp_gated_attn_pool.argtypes = [C.POINTER(StampGatedAttnWeights), C.c_void_p, C.c_int, C.c_int,
                                              C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                              C.c_size_t, C.c_void_p]
        lib.stamp_topk_f32.restype = C.c_int
        lib.stamp_topk_f32.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        lib.stamp_gather_mean_f32.restype = C.c_int
        lib.stamp_gather_mean_f32.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        lib._pool_bound = True
    return lib


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def topk(scores: Tensor, k: int, largest: bool = True) -> tuple[Tensor, Tensor]:
    """Exact top-k of a 1-D fp32 CUDA tensor -> (values, int64 indices); ties by lower index."""
    if not scores.is_cuda or scores.dtype != torch.float32 or scores.dim() != 1 or not scores.is_contiguous():
        raise TypeError("scores must be a contiguous 1-D float32 CUDA tensor")
    idx = torch.empty(k, dtype=torch.int64, device=scores.device)
    val = torch.empty(k, dtype=torch.float32, device=scores.device)
    code = _bind().stamp_topk_f32(scores.data_ptr(), scores.numel(), k, int(largest), idx.data_ptr(),
                                  val.data_ptr(), _stream())
    _lib.check(code, "stamp_topk_f32")
    return val, idx


class GatedAttentionPool(torch.nn.Module):
    """CHIEF's ``attention_net`` (fc + Attn_Net_Gated) + softmax pooling, weights from a CHIEF
    state dict (keys ``attention_net.0.*``, ``attention_net.3.attention_{a,b}.0.*``,
    ``attention_net.3.attention_c.*`` -- chief.py:45-58,255-268)."""

    def __init__(self, state_dict: Mapping[str, Tensor]) -> None:
        super().__init__()
        sd = {k: v.detach().float() for k, v in state_dict.items()}
        fc_w, fc_b = sd["attention_net.0.weight"], sd["attention_net.0.bias"]
        gate = next(k for k in sd if k.endswith("attention_a.0.weight")).rsplit("attention_a", 1)[0]
        a_w, a_b = sd[gate + "attention_a.0.weight"], sd[gate + "attention_a.0.bias"]
        b_w, b_b = sd[gate + "attention_b.0.weight"], sd[gate + "attention_b.0.bias"]
        c_w, c_b = sd[gate + "attention_c.weight"], sd[gate + "attention_c.bias"]
        self.D, self.L, self.Dh = fc_w.shape[1], fc_w.shape[0], a_w.shape[0]
        ab_w = torch.stack([a_w, b_w], 1).reshape(2 * self.Dh, self.L)   # rows a_0, b_0, a_1, b_1, ...
        ab_b = torch.stack([a_b, b_b], 1).reshape(2 * self.Dh)

        def split(name: str, w: Tensor) -> None:
            hi = w.half()
            self.register_buffer(name + "_hi", hi.contiguous(), persistent=False)
            self.register_buffer(name + "_lo", (w - hi.float()).half().contiguous(), persistent=False)

        split("fc_w", fc_w)
        split("ab_w", ab_w)
        self.register_buffer("fc_b", fc_b.contiguous(), persistent=False)
        self.register_buffer("ab_b", ab_b.contiguous(), persistent=False)
        self.register_buffer("c_w", c_w.reshape(-1).contiguous(), persistent=False)
        self.c_b = float(c_b.reshape(-1)[0])

    @torch.no_grad()
    def forward(self, feats: Tensor) -> dict[str, Tensor]:
        """feats fp32 [N, D] (CUDA) -> {"attention_raw": [1, N], "WSI_feature": [1, D]}."""
        if not feats.is_cuda or not self.c_w.is_cuda:
            raise RuntimeError("GatedAttentionPool runs on a CUDA device only (no CPU fallback)")
        x = feats.detach().float().contiguous()
        N, D = x.shape
        if D != self.D:
            raise ValueError(f"expected {self.D}-dim features, got {D}")
        lib = _bind()
        w = StampGatedAttnWeights(self.fc_w_hi.data_ptr(), self.fc_w_lo.data_ptr(), self.fc_b.data_ptr(),
                                  self.ab_w_hi.data_ptr(), self.ab_w_lo.data_ptr(), self.ab_b.data_ptr(),
                                  self.c_w.data_ptr(), self.c_b)
        attn = torch.empty(N, dtype=torch.float32, device=x.device)
        pooled = torch.empty(D, dtype=torch.float32, device=x.device)
        ws = torch.empty(lib.stamp_gated_attn_pool_workspace_bytes(N, D, self.L, self.Dh), dtype=torch.uint8,
                         device=x.device)
        code = lib.stamp_gated_attn_pool(C.byref(w), x.data_ptr(), N, D, self.L, self.Dh, attn.data_ptr(),
                                         pooled.data_ptr(), ws.data_ptr(), ws.numel(), _stream())
        _lib.check(code, "stamp_gated_attn_pool")
        return {"attention_raw": attn[None], "WSI_feature": pooled[None]}


class ChiefB200(Encoder):
    """CHIEF slide encoder (``EncoderName.CHIEF_CTRANSPATH``; needs chief-ctranspath features): the constructor
    arguments of the reference's ``CHIEF`` (chief.py:114-121) with the model replaced by the B200 pooling module."""

    def __init__(self, state_dict: Mapping[str, Tensor]) -> None:
        super().__init__(model=GatedAttentionPool(state_dict), identifier=EncoderName.CHIEF_CTRANSPATH,
                         precision=torch.float32, required_extractors=[ExtractorName.CHIEF_CTRANSPATH])

    def _generate_slide_embedding(self, feats: Tensor, device, **kwargs) -> np.ndarray:
        self.model.to(device)
        out = self.model(feats.to(device).to(self.precision))
        return out["WSI_feature"].float().squeeze().cpu().numpy()

    def _generate_patient_embedding(self, feats_list: list[Tensor], device, **kwargs) -> np.ndarray:
        return self._generate_slide_embedding(torch.cat(feats_list, 0), device)


class EagleB200(_EagleBase):
    """EAGLE (``EncoderName.EAGLE``): CHIEF attention over ctranspath features -> top-25 tiles -> mean of the
    matching Virchow2 features (eagle.py:28-39,92-134).  With STAMP installed this subclasses the reference's
    ``Eagle`` and inherits its feature-file pairing; only the arithmetic is replaced."""

    def __init__(self, state_dict: Mapping[str, Tensor]) -> None:
        self.required_agg_extractor = ExtractorName.VIRCHOW2
        Encoder.__init__(self, model=GatedAttentionPool(state_dict), identifier=EncoderName.EAGLE,
                         precision=torch.float32,
                         required_extractors=[ExtractorName.CTRANSPATH, ExtractorName.CHIEF_CTRANSPATH])

    def _generate_slide_embedding(self, feats: Tensor, device, agg_feats: Tensor | None = None, **kwargs) -> np.ndarray:
        if agg_feats is None:
            raise ValueError("agg_feats is required for slide embedding")
        self.model.to(device)
        attn = self.model(feats.to(device).to(self.precision))["attention_raw"].squeeze(0).contiguous()
        k = min(25, attn.shape[0])
        _, idx = topk(attn, k)
        agg = agg_feats.to(device).float().contiguous()
        out = torch.empty(agg.shape[1], dtype=torch.float32, device=agg.device)
        code = _bind().stamp_gather_mean_f32(agg.data_ptr(), idx.data_ptr(), k, agg.shape[1], out.data_ptr(), _stream())
        _lib.check(code, "stamp_gather_mean_f32")
        return out.cpu().numpy()

    def _generate_patient_embedding(self, feats_list: list[Tensor], device, agg_feats_list: list[Tensor] | None = None,
                                    **kwargs) -> np.ndarray:
        """eagle.py:122-134: all slides of the patient concatenated, then the slide-level selection."""
        if agg_feats_list is None:
            raise ValueError("agg_feats_list is required for patient embedding")
        return self._generate_slide_embedding(torch.cat(feats_list, dim=0), device,
                                              agg_feats=torch.cat(agg_feats_list, dim=0))


def align_by_coords(ref_coords_um, other_coords_um, other_feats: Tensor, decimals: int = 5) -> tuple[Tensor, np.ndarray]:
    """eagle.py:303-330: permute ``other_feats`` so that its tiles come in the order of ``ref_coords_um`` (coordinates
    matched after rounding to ``decimals``; duplicates are consumed first come, first served); raises ``ValueError``
    when a coordinate is missing on either side."""
    from collections import defaultdict, deque

    ref = np.round(np.asarray(ref_coords_um, dtype=np.float64), decimals)
    oth = np.round(np.asarray(other_coords_um, dtype=np.float64), decimals)
    buckets: dict = defaultdict(deque)
    for j, key in enumerate(map(tuple, oth)):
        buckets[key].append(j)
    perm = np.empty(ref.shape[0], dtype=np.int64)
    for i, key in enumerate(map(tuple, ref)):
        if not buckets[key]:
            raise ValueError(f"Missing coord in other set: {key}")
        perm[i] = buckets[key].popleft()
    unused = sum(len(q) for q in buckets.values())
    if unused != 0:
        raise ValueError(f"virchow2 features contain {unused} extra coords not in ref.")
    return other_feats[torch.from_numpy(perm)], np.asarray(other_coords_um)[perm]


if not BOUND_TO_REFERENCE:
    # Stand-alone EAGLE: the feature-file pairing of the reference's ``Eagle`` (eagle.py:40-89, 136-300) -- every slide
    # has a ctranspath file in ``feat_dir`` and a Virchow2 file of the same name in ``agg_feat_dir``; the two are
    # checked against the required extractors and brought into the same tile order by their coordinates.
    def _eagle_read_pair(self, h5_ctp: str, h5_vir2: str, slide_name: str) -> tuple[Tensor, Tensor]:
        from . import features

        feats, coords, extractor = features.read_tile_features(h5_ctp)
        if extractor not in self.required_extractors:
            raise ValueError(f"Features must be extracted with one of {self.required_extractors}. "
                             f"Features located in {h5_ctp} are extracted with {extractor}")
        agg, agg_coords, extractor = features.read_tile_features(h5_vir2)
        if extractor != self.required_agg_extractor:
            raise ValueError(f"Aggregated features must be extracted with {self.required_agg_extractor} "
                             f"Features located in {h5_vir2} are extracted with {extractor}")
        feats_t, agg_t = torch.from_numpy(feats), torch.from_numpy(agg)
        same = coords.coords_um.shape == agg_coords.coords_um.shape and \
            np.allclose(coords.coords_um, agg_coords.coords_um, atol=1e-5, rtol=0)
        if not same:
            try:
                agg_t, _ = align_by_coords(coords.coords_um, agg_coords.coords_um, agg_t)
            except ValueError as e:
                raise ValueError(f"Coordinates mismatch between ctranspath and virchow2 features for slide "
                                 f"{slide_name}. Alignment attempt failed: {e}") from e
        return feats_t, agg_t

    def _eagle_encode_slides(self, output_dir, feat_dir, device, generate_hash: bool = False, **kwargs) -> None:
        import logging
        import os
        from pathlib import Path

        agg_feat_dir = kwargs.get("agg_feat_dir")
        if not agg_feat_dir:
            raise ValueError("agg_feat_dir that contains virchow2 features is required for Eagle's encode_slides")
        encode_dir = self._encode_dir(output_dir, "slide", generate_hash)
        self.model.to(device).eval()
        for name in sorted(os.listdir(feat_dir)):
            if not name.endswith(".h5"):
                continue
            output_path = (encode_dir / Path(name).name).with_suffix(".h5")
            if output_path.exists():
                continue
            try:
                feats, agg = _eagle_read_pair(self, os.path.join(feat_dir, name), os.path.join(agg_feat_dir, name), name)
            except ValueError as e:
                logging.getLogger("stamp").warning(str(e))
                continue
            self._save_features_(output_path, self._generate_slide_embedding(feats, device, agg), "slide")

    def _eagle_encode_patients(self, output_dir, feat_dir, slide_table_path, patient_label: str, filename_label: str,
                               device, generate_hash: bool = False, **kwargs) -> None:
        import logging
        import os
        from pathlib import Path

        import pandas as pd

        agg_feat_dir = kwargs.get("agg_feat_dir")
        if not agg_feat_dir:
            raise ValueError("agg_feat_dir that contains virchow2 features is required for Eagle's encode_patients")
        encode_dir = self._encode_dir(output_dir, "pat", generate_hash)
        self.model.to(device).eval()
        for patient_id, group in pd.read_csv(slide_table_path).groupby(patient_label):
            output_path = (encode_dir / str(patient_id)).with_suffix(".h5")
            if output_path.exists():
                continue
            feats_list, agg_list = [], []
            for _, row in group.iterrows():
                fn = row[filename_label]
                try:
                    feats, agg = _eagle_read_pair(self, os.path.join(feat_dir, fn), os.path.join(agg_feat_dir, fn), Path(fn).stem)
                except FileNotFoundError as e:
                    logging.getLogger("stamp").warning(f"[{patient_id}] skip slide (FileNotFoundError): {fn} -> {e}")
                    continue
                feats_list.append(feats)
                agg_list.append(agg)
            if not feats_list:
                continue
            self._save_features_(output_path, self._generate_patient_embedding(feats_list, device, agg_list), "patient")

    EagleB200.encode_slides_ = _eagle_encode_slides          # type: ignore[method-assign]
    EagleB200.encode_patients_ = _eagle_encode_patients      # type: ignore[method-assign]


# ---- TITAN: NOT built.  The slide transformer is un-vendored Hugging Face remote code with
# gated weights; nothing here restates or accelerates it.  What remains are the two pieces of input preparation of the
# reference's wrapper that are in-tree arithmetic, for callers that feed their own TITAN model. -----------------------
def titan_coords_px(coords_um, mpp: float, device=None) -> Tensor:
    """titan.py:47-53: micron coordinates -> level-0 pixel coordinates, truncated to int64."""
    c = torch.as_tensor(coords_um, dtype=torch.float32)
    c = (c / mpp).to(torch.int64)
    return c if device is None else c.to(device)


def titan_virtual_slide(feats_list: list[Tensor], coords_um_list: list, tile_size_um: float):
    """titan.py:131-168 + :64-83: all slides of a patient laid side by side along x (each slide shifted by the
    right edge of the previous one), features concatenated -> (feats [1, N, D], coords_um [N, 2])."""
    shifted, offset = [], 0.0
    for c in coords_um_list:
        c = np.array(c, dtype=np.float64, copy=True)
        c[:, 0] += offset
        offset = float(c[:, 0].max()) + float(tile_size_um)
        shifted.append(c)
    return torch.cat(feats_list, dim=0).unsqueeze(0), np.concatenate(shifted, axis=0)
