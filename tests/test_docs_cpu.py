"""Keeps the documents honest: every C-ABI symbol, profile file, golden fixture and source file that DESIGN.md /
INTEGRATION.md / profiles/README.md name must exist in the tree."""

import re
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
DOCS = [ROOT / "DESIGN.md", ROOT / "INTEGRATION.md", ROOT / "profiles" / "README.md", ROOT / "README.md"]


def _text() -> str:
    return "\n".join(p.read_text() for p in DOCS)


def test_c_abi_symbols_named_in_the_docs_are_declared():
    header = (ROOT / "include" / "stamp_b200.h").read_text()
    declared = set(re.findall(r"\b(stamp_[a-z0-9_]+)\s*\(", re.sub(r"/\*.*?\*/", "", header, flags=re.S)))
    named = set(re.findall(r"`(stamp_[a-z0-9_]+)`", _text()))
    named = {n for n in named if not n.startswith("stamp_b200") or n in declared}   # `stamp_b200` is the package name
    missing = sorted(n for n in named if n not in declared and not (ROOT / n).exists())
    assert not missing, missing


def test_files_named_in_the_docs_exist():
    text = _text()
    paths = set(re.findall(r"`((?:profiles|tests/golden|oracle|scripts|stamp_b200|include)/[A-Za-z0-9_./-]+\.[a-z0-9]+)`", text))
    paths |= {"profiles/" + m for m in re.findall(r"`(r1_[A-Za-z0-9_]+\.(?:csv|json|txt))`", text)}
    paths |= {"tests/golden/" + m for m in re.findall(r"`((?:mil|chief|texture)_[A-Za-z0-9_]+\.npz)`", text)}
    missing = sorted(p for p in paths if "*" not in p and not (ROOT / p).exists())
    assert not missing, missing
    for src in set(re.findall(r"`([a-z_0-9]+\.(?:cu|cuh))`", text)):
        assert (ROOT / "stamp_b200" / "csrc" / src).exists(), src
    for mod in set(re.findall(r"`stamp_b200\.([a-z_]+)(?:\.[A-Za-z_]+)*`", text)):
        assert (ROOT / "stamp_b200" / f"{mod}.py").exists(), mod


def test_every_golden_fixture_has_its_generating_script():
    scripts = "\n".join(p.read_text() for p in (ROOT / "oracle").glob("make_golden*.py"))
    for f in (ROOT / "tests" / "golden").glob("*.npz"):
        stem = f.stem
        assert stem in scripts or re.sub(r"_(alibi|mha).*", "", stem) in scripts or f"mil_{{name}}" in scripts, stem
