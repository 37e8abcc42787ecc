"""The dispatch of ``extract_`` and the three lines of its hot loop that touch the extractor."""
from typing import assert_never

import torch

from stamp.preprocessing.config import ExtractorName
from stamp.preprocessing.extractor import Extractor


def extract_(*, extractor: ExtractorName | Extractor, tiles, device):
    match extractor:
        case ExtractorName.UNI | ExtractorName.VIRCHOW2 | ExtractorName.EMPTY:
            raise RuntimeError("the stub has no built-in extractors")
        case Extractor():
            extractor = extractor
        case _ as unreachable:
            assert_never(unreachable)
    model = extractor.model.to(device).eval()
    batch = torch.stack([extractor.transform(t) for t in tiles])
    return extractor.identifier, model, batch
