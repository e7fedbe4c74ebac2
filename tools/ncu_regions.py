#!/usr/bin/env python
"""ncu_regions.py <source.csv> <file> <name:first-last> ... -- executed warp instructions per source-line range of one file, from
`ncu -i X.ncu-rep --page source --csv --print-source cuda,sass` (profile reading aid)."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
want = sys.argv[2]
per = collections.Counter()
cur = None
for r in rows:
    if r and r[0] == "File Path":
        cur = r[1].split("/")[-1]
        continue
    if cur and len(r) > 8 and r[0].strip().isdigit() and r[2].strip() == "-":
        per[(cur, int(r[0]))] += int(r[7] or 0)
tot = sum(per.values())
print("all files: %.1f M warp instructions" % (tot / 1e6))
for f in sorted({k[0] for k in per}):
    print("  %-36s %8.1f M" % (f, sum(v for (ff, l), v in per.items() if ff == f) / 1e6))
for spec in sys.argv[3:]:
    name, rg = spec.rsplit(":", 1)
    a, b = [int(x) for x in rg.split("-")]
    v = sum(v for (ff, l), v in per.items() if ff == want and a <= l <= b)
    print("%-36s %8.1f M  %5.1f %%" % (name + " " + rg, v / 1e6, 100.0 * v / tot))
