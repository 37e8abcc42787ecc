"""Generates tests/golden/chief_pool.npz by running the REFERENCE's CHIEFModel itself.

Run in the build container only (``python oracle/make_golden_chief.py``).  ``chief.py`` imports ``gdown`` and
``stamp.*`` package modules at its top (neither importable offline), but its model classes are plain torch:
the file is loaded by path with inert stand-ins for exactly those imports, ``CHIEFModel(size_arg="small")`` is
instantiated as ``CHIEF.__init__`` does (src/stamp/encoding/encoder/chief.py:91-93), given the seeded synthetic
weights of ``oracle/chief_oracle.init_state_dict`` (the pretrained ones live on Google Drive) and run in eval
mode on seeded features.  Stored: inputs, the outputs ``attention_raw`` / ``WSI_feature``, and EAGLE's top-25
selection + mean (eagle.py:104-118) computed with the reference's own lines.  The weights are regenerated from
the seed by the tests (same torch, same generator), so only inputs and outputs are stored.
"""

from __future__ import annotations

import importlib.util
import sys
import types
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from oracle import chief_oracle as co  # noqa: E402

REF = Path("/root/reference/src/stamp/encoding/encoder/chief.py")
OUT = Path(__file__).resolve().parent.parent / "tests" / "golden"


def load_reference_chief():
    def stub(name: str, **attrs):
        m = types.ModuleType(name)
        for k, v in attrs.items():
            setattr(m, k, v)
        sys.modules[name] = m
        return m

    class _Any:                      # inert stand-in for enums / base classes the model code never touches
        def __getattr__(self, item):
            return item

        def __init_subclass__(cls, **kw):
            pass

    stub("gdown")
    for pkg in ("stamp", "stamp.encoding", "stamp.preprocessing", "stamp.utils"):
        stub(pkg).__path__ = []      # mark as packages
    stub("stamp.encoding.config", EncoderName=_Any())
    stub("stamp.encoding.encoder", Encoder=_Any)
    stub("stamp.preprocessing.config", ExtractorName=_Any())
    stub("stamp.types", DeviceLikeType=object, PandasLabel=str)
    stub("stamp.utils.cache", STAMP_CACHE_DIR=Path("/nonexistent"), file_digest=None, get_processing_code_hash=None)
    spec = importlib.util.spec_from_file_location("ref_chief", REF)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def main() -> None:
    ref = load_reference_chief()
    torch.manual_seed(0)
    model = ref.CHIEFModel(size_arg="small", dropout=True, n_classes=2).eval()
    sd = co.init_state_dict(seed=3)
    missing, unexpected = model.load_state_dict(sd, strict=False)
    assert not unexpected and all(not k.startswith("attention_net.") for k in missing), (missing, unexpected)
    g = torch.Generator().manual_seed(11)
    arrays = {}
    for name, n in (("small", 37), ("slide", 600)):
        # fp16-representable values (what the feature files hold), so the fixture stores them losslessly in fp16
        x = torch.randn(n, 768, generator=g)
        x[: n // 10] += 0.3
        x = x.half().float()
        agg = torch.randn(n, 256, generator=g).half().float()
        with torch.no_grad():
            res = model(x)
            attention_raw = res["attention_raw"].squeeze(0).cpu()
            k = min(25, attention_raw.shape[0])                       # eagle.py:107-118
            _, topk_indices = torch.topk(attention_raw, k)
            top_agg_feats = torch.stack([agg[i] for i in topk_indices.numpy()])
            eagle = torch.mean(top_agg_feats, dim=0)
        arrays.update({f"{name}/x": x.numpy().astype(np.float16), f"{name}/agg": agg.numpy().astype(np.float16),
                       f"{name}/attention_raw": res["attention_raw"].numpy(), f"{name}/wsi": res["WSI_feature"].numpy(),
                       f"{name}/topk": topk_indices.numpy(), f"{name}/eagle": eagle.numpy()})
        srt = attention_raw.sort(descending=True).values
        print(name, "top-25 margin", float(srt[k - 1] - srt[k]) if n > k else None)
    arrays["weights_seed"] = np.int64(3)
    np.savez_compressed(OUT / "chief_pool.npz", **arrays)


if __name__ == "__main__":
    main()
