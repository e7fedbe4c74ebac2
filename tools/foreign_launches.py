"""Print the per-kernel times of an ncu launch list (gpu__time_duration.sum) of the foreign-stream pipeline."""
import csv
import sys

for path in sys.argv[1:]:
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    scale = {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3, "s": 1e6, "second": 1e6, "nsecond": 1e-3}
    print("==", path)
    total = 0.0
    for r in rows[1:]:
        v = float(r[vi].replace(",", "")) * scale.get(r[ui], 1.0)
        total += v
        print(f"  {r[ki].split('(')[0][-28:]:30s} {v:10.1f} us")
    print(f"  total {total / 1e3:.2f} ms")
