"""The drop-in boundary has to BIND: ``extract_`` and ``init_slide_encoder_`` dispatch with ``match`` statements on
the reference's own classes (``case Extractor():`` src/stamp/preprocessing/__init__.py:237-238, ``case Encoder():``
src/stamp/encoding/__init__.py:72-73) and end in ``assert_never``.  A stand-in package reproducing those two
dispatch points (tests/ref_stub) is put on ``sys.path`` in a subprocess; the factories of ``stamp_b200`` must then
hand out objects that go through them.  Without a ``stamp`` package the local stand-in classes are used."""

import subprocess
import sys
from pathlib import Path

import pytest
import torch

ROOT = Path(__file__).resolve().parent.parent
STUB = ROOT / "tests" / "ref_stub"

_DRIVER = r"""
import sys
sys.path.insert(0, {stub!r}); sys.path.insert(0, {root!r})
import torch
import stamp.preprocessing as ref_pre, stamp.encoding as ref_enc
from stamp.preprocessing.extractor import Extractor as RefExtractor
from stamp.encoding.encoder import Encoder as RefEncoder
from stamp.encoding.config import EncoderName
from stamp_b200 import extractor as X, encoder as E
from stamp_b200.vit import VitArch, random_state_dict

assert X.BOUND_TO_REFERENCE and E.BOUND_TO_REFERENCE
assert X.Extractor is RefExtractor and E.Encoder is RefEncoder
arch = VitArch("tiny", dim=64, depth=1, heads=1, mlp_hidden=128)
ext = X._make(arch, "uni", random_state_dict(arch), None, {{}}, 4)
assert isinstance(ext, RefExtractor)
ident, model, batch = ref_pre.extract_(extractor=ext, tiles=[torch.zeros(224, 224, 3, dtype=torch.uint8).numpy()] * 2,
                                       device="cpu")
assert ident == "uni" and batch.dtype == torch.uint8 and tuple(batch.shape) == (2, 224, 224, 3)
assert not model.training
try:
    ref_pre.extract_(extractor=object(), tiles=[], device="cpu")       # a look-alike must NOT get through
except AssertionError:
    pass
else:
    raise SystemExit("assert_never did not fire for a foreign object")

g = torch.Generator().manual_seed(0)
sd = {{"attention_net.0.weight": torch.randn(512, 768, generator=g) * 0.03, "attention_net.0.bias": torch.zeros(512),
      "attention_net.3.attention_a.0.weight": torch.randn(256, 512, generator=g) * 0.03,
      "attention_net.3.attention_a.0.bias": torch.zeros(256),
      "attention_net.3.attention_b.0.weight": torch.randn(256, 512, generator=g) * 0.03,
      "attention_net.3.attention_b.0.bias": torch.zeros(256),
      "attention_net.3.attention_c.weight": torch.randn(1, 256, generator=g) * 0.03,
      "attention_net.3.attention_c.bias": torch.zeros(1)}}
for cls, name in ((E.ChiefB200, EncoderName.CHIEF_CTRANSPATH), (E.EagleB200, EncoderName.EAGLE)):
    enc = cls(sd)
    assert isinstance(enc, RefEncoder) and enc.identifier == name and enc.precision == torch.float32
    assert len(enc.required_extractors) >= 1
    assert ref_enc.init_slide_encoder_(enc, output_dir="o")[:2] == ("walked", name)    # the inherited h5 walk runs
assert E.EagleB200(sd).required_agg_extractor == "virchow2"
print("BOUND")
"""


def test_factories_bind_to_the_reference_classes_when_stamp_is_importable():
    code = _DRIVER.format(stub=str(STUB), root=str(ROOT))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "BOUND" in r.stdout, r.stdout + r.stderr


def test_stand_alone_classes_have_the_reference_shape():
    """No ``stamp`` package in this interpreter: the stand-ins keep constructor, fields and enum values."""
    from stamp_b200 import encoder as E
    from stamp_b200 import extractor as X

    if X.BOUND_TO_REFERENCE:
        pytest.skip("a stamp package is importable here")
    ext = X.Extractor(model=torch.nn.Identity(), transform=X.pil_to_u8_hwc, identifier="uni")
    with pytest.raises(Exception):
        ext.identifier = "other"                                   # frozen, like the reference's
    with pytest.raises(TypeError):
        X.Extractor(torch.nn.Identity(), X.pil_to_u8_hwc, "uni")   # keyword-only
    assert E.EncoderName.CHIEF_CTRANSPATH == "chief" and E.EncoderName.EAGLE == "eagle"
    assert E.ExtractorName.CHIEF_CTRANSPATH == "chief-ctranspath"
    assert issubclass(E.ChiefB200, E.Encoder) and issubclass(E.EagleB200, E.Encoder)
    with pytest.raises(TypeError):
        E.Encoder(None, "x", torch.float32, [])                    # abstract


def test_extractor_identifiers_are_the_reference_enum_values():
    import inspect

    from stamp_b200 import extractor as X

    src = {name: inspect.getsource(getattr(X, name)) for name in ("uni", "virchow2", "uni2", "h_optimus_0", "h_optimus_1")}
    for name, ident in (("uni", '"uni"'), ("virchow2", '"virchow2"'), ("uni2", '"uni2"'),
                        ("h_optimus_0", '"h-optimus-0"'), ("h_optimus_1", '"h-optimus-1"')):
        assert ident in src[name], name
