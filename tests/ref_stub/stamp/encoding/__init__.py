"""The dispatch of ``init_slide_encoder_``."""
from typing import assert_never

from stamp.encoding.config import EncoderName
from stamp.encoding.encoder import Encoder


def init_slide_encoder_(encoder: EncoderName | Encoder, **kwargs):
    match encoder:
        case EncoderName.TITAN | EncoderName.EAGLE | EncoderName.CHIEF_CTRANSPATH:
            raise RuntimeError("the stub has no built-in encoders")
        case Encoder():
            selected = encoder
        case _ as unreachable:
            assert_never(unreachable)
    return selected.encode_slides_(**kwargs)
