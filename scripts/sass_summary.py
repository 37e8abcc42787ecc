"""Per-kernel SASS evidence of the Blackwell-native paths: counts of the mnemonics that prove tcgen05 / TMEM / TMA use
(B200_PROFILING.md: tcgen05.mma -> UTC*MMA, tcgen05.ld / .st -> LDTM / STTM, cp.async.bulk.tensor -> UTMALDG / UTMASTG /
UTMAREDG, mma.sync -> HMMA) in every kernel of libstamp_b200.so.

    python scripts/sass_summary.py > profiles/r2_sass_summary.txt
"""
import re
import subprocess
import sys
from collections import OrderedDict
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
LIB = ROOT / "stamp_b200" / "libstamp_b200.so"
PATTERNS = OrderedDict([
    ("UTC*MMA", re.compile(r"\bUTC[A-Z]*MMA\b")), ("UTC*MMA.2CTA", re.compile(r"\bUTC[A-Z]*MMA\.2CTA|UTC[A-Z]*MMA[.\w]*\.2CTA")),
    ("LDTM", re.compile(r"\bLDTM\b")), ("STTM", re.compile(r"\bSTTM\b")),
    ("UTMALDG", re.compile(r"\bUTMALDG\b")), ("UTMASTG", re.compile(r"\bUTMASTG\b")), ("UTMAREDG", re.compile(r"\bUTMAREDG\b")),
    ("HMMA", re.compile(r"\bHMMA\b")), ("MUFU", re.compile(r"\bMUFU\b")),
])


def demangle(name: str) -> str:
    try:
        out = subprocess.run(["cu++filt", name], capture_output=True, text=True).stdout.strip()
    except OSError:
        out = name
    out = re.sub(r"\(anonymous namespace\)::|sb::|<unnamed>::", "", out)
    return out.split("(")[0][:70]


def main() -> None:
    sass = subprocess.run(["cuobjdump", "-sass", str(LIB)], capture_output=True, text=True, check=True).stdout
    kernels: "OrderedDict[str, dict]" = OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = kernels.setdefault(demangle(m.group(1)), {k: 0 for k in PATTERNS} | {"instr": 0})
            continue
        if cur is None or "/*" not in line:
            continue
        body = line.split("*/", 1)[-1]
        if re.search(r"\b[A-Z][A-Z0-9_.]+\b", body):
            cur["instr"] += 1
        for k, pat in PATTERNS.items():
            if pat.search(body):
                cur[k] += 1
    cols = list(PATTERNS)
    print(f"# cuobjdump -sass {LIB.relative_to(ROOT)} (sm_100a): mnemonic counts per kernel")
    print(f"{'kernel':70s} " + " ".join(f"{c:>12s}" for c in cols))
    tot = {c: 0 for c in cols}
    for name, d in sorted(kernels.items()):
        if not any(d[c] for c in cols if c != "MUFU"):
            continue
        print(f"{name:70s} " + " ".join(f"{d[c]:12d}" for c in cols))
        for c in cols:
            tot[c] += d[c]
    print(f"{'TOTAL (kernels listed)':70s} " + " ".join(f"{tot[c]:12d}" for c in cols))
    plain = [n for n, d in kernels.items() if not any(d[c] for c in cols if c != "MUFU")]
    print(f"# {len(plain)} further kernels use none of these (row-wise / HBM-bound CUDA-core kernels): " + ", ".join(sorted(plain)))


if __name__ == "__main__":
    sys.exit(main())
