from enum import StrEnum


class ExtractorName(StrEnum):
    CTRANSPATH = "ctranspath"
    CHIEF_CTRANSPATH = "chief-ctranspath"
    UNI = "uni"
    UNI2 = "uni2"
    H_OPTIMUS_0 = "h-optimus-0"
    H_OPTIMUS_1 = "h-optimus-1"
    VIRCHOW2 = "virchow2"
    EMPTY = "empty"
