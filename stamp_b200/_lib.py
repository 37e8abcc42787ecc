"""ctypes binding of ``libstamp_b200.so`` (the C ABI declared in ``include/stamp_b200.h``).

There is deliberately no fallback: if the library is missing or a call fails the caller gets an
exception.  PyTorch is only used by callers for device memory and streams; this module passes raw
device pointers (``tensor.data_ptr()``) and the current CUDA stream handle.
"""

from __future__ import annotations

import ctypes as C
from pathlib import Path

_LIB_PATH = Path(__file__).resolve().parent / "libstamp_b200.so"
_lib: C.CDLL | None = None

c_void_p, c_int, c_ll, c_float = C.c_void_p, C.c_int, C.c_longlong, C.c_float
c_fp = C.POINTER(C.c_float)

# name -> (restype, argtypes); must list every symbol of include/stamp_b200.h
SIGNATURES: dict[str, tuple] = {
    "stamp_b200_abi_version": (c_int, []),
    "stamp_b200_strerror": (C.c_char_p, [c_int]),
    "stamp_b200_launch_count": (c_ll, []),
    "stamp_b200_reset_launch_count": (None, []),
    "stamp_b200_profile_enable": (None, [c_int]),
    "stamp_b200_profile_summary": (c_int, [C.POINTER(C.c_double), C.POINTER(C.c_double),
                                           C.POINTER(C.c_longlong), c_int]),
    "stamp_b200_gemm_force_mode": (None, [c_int]),
    "stamp_b200_attention_tc_enable": (None, [c_int]),
    "stamp_gemm_tn": (c_int, [c_void_p, c_ll, c_void_p, c_ll, c_void_p, c_ll, c_int, c_int, c_int,
                              c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_ll, c_int, c_int,
                              c_int, c_void_p]),
    "stamp_layernorm": (c_int, [c_void_p, c_ll, c_void_p, c_void_p, c_void_p, c_void_p, c_ll, c_int, c_int,
                                c_float, c_int, c_void_p]),
    "stamp_fill_rows": (c_int, [c_void_p, c_ll, c_int, c_int, c_int, c_void_p, c_ll, c_void_p, c_ll,
                                c_int, c_int, c_void_p]),
    "stamp_tiles_to_patches": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_fp, c_fp,
                                       c_int, c_void_p]),
    "stamp_attention_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_ll, c_ll, c_void_p, c_ll, c_ll,
                                    c_int, c_int, c_int, c_int, c_int, c_float, c_void_p, c_void_p,
                                    c_void_p, c_void_p, c_int, c_void_p]),
    "stamp_alibi_dist_scale": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p]),
}


class StampB200Error(RuntimeError):
    """A C-ABI call returned a negative status."""


def lib_path() -> Path:
    return _LIB_PATH


def load() -> C.CDLL:
    """Load the shared library (once) and attach signatures. Fails loudly if it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not _LIB_PATH.exists():
        raise ImportError(
            f"{_LIB_PATH} is missing: build it with `python -m stamp_b200.build` "
            "(or __graft_entry__.build()). There is no CPU fallback for the B200 hot path."
        )
    lib = C.CDLL(str(_LIB_PATH))
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(code: int, what: str) -> None:
    if code != 0:
        msg = load().stamp_b200_strerror(code).decode()
        raise StampB200Error(f"{what} failed with status {code}: {msg}")


def launch_count() -> int:
    return int(load().stamp_b200_launch_count())


def reset_launch_count() -> None:
    load().stamp_b200_reset_launch_count()


PROFILE_CATEGORIES = ("gemm", "attention", "rowops", "macenko", "pool")


def profile_enable(on: bool) -> None:
    load().stamp_b200_profile_enable(int(on))


def profile_summary() -> dict[str, dict[str, float]]:
    """Per-category {ms, work, count} of the launches recorded since the last summary."""
    n = len(PROFILE_CATEGORIES)
    ms, work, cnt = (C.c_double * n)(), (C.c_double * n)(), (C.c_longlong * n)()
    load().stamp_b200_profile_summary(ms, work, cnt, n)
    return {name: {"ms": ms[i], "work": work[i], "count": int(cnt[i])}
            for i, name in enumerate(PROFILE_CATEGORIES)}
