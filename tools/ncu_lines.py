#!/usr/bin/env python
"""Summarise an ncu source page (ncu -i X.ncu-rep --page source --csv --print-source cuda,sass) per source line:
instructions executed and stall samples, for one kernel.  Usage: ncu_lines.py src.csv <kernel substring> [top N]"""
import csv
import sys
from collections import defaultdict

path, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
rows = list(csv.reader(open(path)))
i = 0
blocks = []
while i < len(rows):
    if rows[i] and rows[i][0] == "File Path":
        fpath = rows[i][1]
        func = rows[i + 1][1]
        hdr = rows[i + 2]
        j = i + 3
        body = []
        while j < len(rows) and not (rows[j] and rows[j][0] == "File Path"):
            body.append(rows[j])
            j += 1
        blocks.append((fpath, func, hdr, body))
        i = j
    else:
        i += 1
seen = set()
for fpath, func, hdr, body in blocks:
    if kern not in func:
        continue
    key = (fpath, func)
    if key in seen:
        continue
    seen.add(key)
    col = {h: k for k, h in enumerate(hdr)}
    if "Line No" not in col:
        continue
    # rows with a Line No are source lines with aggregated metrics
    tot_inst = tot_samp = 0
    lines = []
    for r in body:
        if len(r) < len(hdr) or not r[col["Line No"]].strip():
            continue
        try:
            inst = float(r[col["Instructions Executed"]] or 0)
            samp = float(r[col["# Samples"]] or 0)
        except ValueError:
            continue
        if inst == 0 and samp == 0:
            continue
        stalls = {h: float(r[col[h]] or 0) for h in hdr if h.startswith("stall_") and "Not Issued" not in h}
        topst = sorted(stalls.items(), key=lambda kv: -kv[1])[:2]
        lines.append((inst, samp, r[col["Line No"]], r[1].strip()[:90], topst))
        tot_inst += inst
        tot_samp += samp
    print(f"== {fpath}  [{func[:60]}]  inst={tot_inst:.3g} samples={tot_samp:.0f}")
    for inst, samp, ln, src, topst in sorted(lines, key=lambda x: -x[1])[:top]:
        st = " ".join(f"{k[6:]}={v:.0f}" for k, v in topst if v)
        print(f"  {100*inst/max(tot_inst,1):5.1f}%i {100*samp/max(tot_samp,1):5.1f}%s  L{ln:>4} {src}   [{st}]")
