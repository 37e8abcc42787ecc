"""In-tree build of the C-ABI library ``libstamp_b200.so`` (nvcc, sm_100a only).

The library is plain CUDA C++ behind ``extern "C"`` entry points (see
``include/stamp_b200.h``); it links the CUDA runtime statically and resolves the one
driver symbol it needs (``cuTensorMapEncodeTiled``) at run time, so it has no torch or
libcuda link-time dependency and cross-compiles on a machine without a GPU.
"""

from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
OBJ = PKG / "_build"
LIB = PKG / "libstamp_b200.so"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
    "-I", str(PKG.parent / "include"),
]


def _nvcc() -> str:
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not Path(nvcc).exists():
        raise RuntimeError("nvcc not found; cannot build libstamp_b200.so")
    return nvcc


def _newer(a: Path, deps: list[Path]) -> bool:
    if not a.exists():
        return False
    t = a.stat().st_mtime
    return all(d.stat().st_mtime <= t for d in deps)


def build(force: bool = False, verbose: bool = False) -> Path:
    """Compile every ``csrc/*.cu`` and link the shared library. Returns its path."""
    nvcc = _nvcc()
    OBJ.mkdir(exist_ok=True)
    headers = sorted(CSRC.glob("*.cuh")) + sorted((PKG.parent / "include").glob("*.h"))
    sources = sorted(CSRC.glob("*.cu"))
    jobs = []
    for src in sources:
        obj = OBJ / (src.stem + ".o")
        if force or not _newer(obj, [src, *headers]):
            jobs.append((src, obj))

    def compile_one(job):
        src, obj = job
        cmd = [nvcc, *NVCC_FLAGS, "-c", str(src), "-o", str(obj)]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src.name}:\n{r.stdout}\n{r.stderr}")
        if verbose:
            sys.stderr.write(r.stderr)
        return obj

    if jobs:
        with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
            list(ex.map(compile_one, jobs))
    objs = [OBJ / (s.stem + ".o") for s in sources]
    if force or jobs or not _newer(LIB, objs):
        cmd = [nvcc, "-shared", "-o", str(LIB), *map(str, objs),
               "-gencode", "arch=compute_100a,code=sm_100a", "-cudart", "static"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return LIB


if __name__ == "__main__":
    p = build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(p)
