from enum import StrEnum


class EncoderName(StrEnum):
    EAGLE = "eagle"
    CHIEF_CTRANSPATH = "chief"
    TITAN = "titan"
