#!/usr/bin/env python
"""Reads dram__bytes_read.sum + dram__bytes_write.sum per launch out of ncu reports (ncu -i X --page raw --csv)
and writes profiles/traffic.json, which bench.py copies into roofline.traffic.
usage: ncu_traffic.py <encode.ncu-rep> <decode.ncu-rep> "<description of the profiled command>" """
import csv
import io
import json
import os
import subprocess
import sys

UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def traffic(rep, kernel_substr):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    vals = []
    for r in rows[2:]:
        if kernel_substr not in r[idx["Kernel Name"]]:
            continue
        tot = 0.0
        for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            tot += float(r[idx[m]]) * UNIT[units[idx[m]]]
        vals.append((tot, float(r[idx["gpu__time_duration.sum"]]), units[idx["gpu__time_duration.sum"]], r[idx["Grid Size"]]))
    if not vals:
        return None
    return {"dram_bytes_per_launch": sum(v[0] for v in vals) / len(vals), "launches_averaged": len(vals),
            "duration_under_ncu": "%.3f %s" % (sum(v[1] for v in vals) / len(vals), vals[0][2]), "grid": vals[0][3]}


if __name__ == "__main__":
    enc, dec, desc = sys.argv[1], sys.argv[2], sys.argv[3]
    e, d = traffic(enc, "qoi_encode_kernel"), traffic(dec, "qoi_decode_kernel")
    res = {"encode": e and e["dram_bytes_per_launch"], "decode": d and d["dram_bytes_per_launch"],
           "detail": {"encode": e, "decode": d}, "source": desc}
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "traffic.json")
    json.dump(res, open(path, "w"), indent=1)
    print(json.dumps(res, indent=1))
