from collections.abc import Callable
from dataclasses import KW_ONLY, dataclass
from typing import Any, Generic, TypeVar

M = TypeVar("M")


@dataclass(frozen=True)
class Extractor(Generic[M]):
    _: KW_ONLY
    model: M
    transform: Callable[[Any], Any]
    identifier: str
