"""Summarise an ncu launch list (--metrics gpu__time_duration.sum,... --csv): per-kernel totals / shares and one block."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
hdr, data = rows[hi], rows[hi + 1:]
ki, mi, vi, idi = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("ID")
launch = collections.OrderedDict()
for r in data:
    if len(r) <= vi:
        continue
    d = launch.setdefault(r[idi], {"name": r[ki]})
    d[r[mi]] = float(r[vi].replace(",", ""))
T, TA = "gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"
agg, seq = collections.OrderedDict(), []
for d in launch.values():
    n = d["name"].split("(")[0][-44:]
    seq.append((n, d))
    a = agg.setdefault(n, [0, 0.0, 0.0])
    a[0] += 1
    a[1] += d[T]
    a[2] += d.get(TA, 0) * d[T]
tot = sum(a[1] for a in agg.values())
for n, a in agg.items():
    print(f"{n:46s} n={a[0]:4d} total={a[1] / 1e3:9.1f}us share={a[1] / tot * 100:5.1f}% tensor={a[2] / max(a[1], 1):5.1f}%")
print(f"total {tot / 1e3:.1f} us over {len(seq)} launches")
lo = int(sys.argv[2]) if len(sys.argv) > 2 else 0
hi_ = int(sys.argv[3]) if len(sys.argv) > 3 else 0
for n, d in seq[lo:hi_]:
    print(f"{d[T] / 1e3:8.1f}us tensor {d.get(TA, 0):5.1f}% issue {d.get('smsp__issue_active.avg.pct_of_peak_sustained_active', 0):5.1f}% "
          f"rd {d.get('dram__bytes_read.sum', 0) / 1e6:7.1f}MB wr {d.get('dram__bytes_write.sum', 0) / 1e6:7.1f}MB {n}")
