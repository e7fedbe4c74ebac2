#!/usr/bin/env python
"""Turns the ncu reports in gpurun_out/ (prof_enc.ncu-rep, prof_dec.ncu-rep, launches.csv) into the text summaries
committed under profiles/<name>/.  Runs where ncu is installed (no GPU needed: it only reads the reports).
usage: make_profile_summary.py <name> "<profiled command>" """
import csv
import io
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
name, cmd = sys.argv[1], sys.argv[2]
out = os.path.join(ROOT, "profiles", name)
os.makedirs(out, exist_ok=True)
WANT = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.per_cycle_active", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__cycles_elapsed.max"]


def raw_summary(rep, fh):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        fh.write("-----\n")
        for w in WANT:
            if w in idx:
                fh.write(f"{w} = {r[idx[w]]} {units[idx[w]]}\n")
        st = [h for h in hdr if "issue_stalled" in h and h.endswith("_per_issue_active.ratio")]
        for v, h in sorted([(float(r[idx[h]] or 0), h) for h in st], reverse=True)[:8]:
            fh.write("  stall %.2f %s\n" % (v, h.split("issue_stalled_")[1]))


with open(os.path.join(out, "ncu_raw_summary.txt"), "w") as fh:
    fh.write(f"ncu --set full --clock-control none --import-source on, B200; profiled command: {cmd}\n")
    fh.write("(durations under ncu are cold-cache and serialised; bench.py's CUDA-event numbers are the reported ones)\n")
    for rep in ("prof_enc.ncu-rep", "prof_dec.ncu-rep"):
        p = os.path.join(ROOT, "gpurun_out", rep)
        if os.path.exists(p):
            fh.write(f"\n===== {rep} =====\n")
            raw_summary(p, fh)
for rep, kern, fn in (("prof_enc.ncu-rep", "qoi_encode_kernel", "encode_lines.txt"), ("prof_dec.ncu-rep", "qoi_decode_kernel", "decode_lines.txt")):
    p = os.path.join(ROOT, "gpurun_out", rep)
    if not os.path.exists(p):
        continue
    src = subprocess.run(["ncu", "-i", p, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
    tmp = os.path.join(out, "_src.csv")
    open(tmp, "w").write(src)
    res = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_lines.py"), tmp, kern, "30"], capture_output=True, text=True).stdout
    open(os.path.join(out, fn), "w").write("per source line: share of executed warp instructions (i) and of warp stall samples (s)\n" + res)
    os.unlink(tmp)
# the foreign-stream (any-stream) decoder: launch list of its pipeline + batch throughput per distribution
fl = os.path.join(ROOT, "gpurun_out", "foreign_launches_d4.csv")
if os.path.exists(fl):
    res = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "foreign_launches.py"), fl], capture_output=True, text=True).stdout
    extra = ""
    fb = os.path.join(ROOT, "gpurun_out", "foreign_batch32.txt")
    if os.path.exists(fb):
        extra = "\nCUDA-event times of qoi_gpu_decode(use_index=0), not under ncu (tools/foreign_prof.py <dist> <batch> 3):\n" + open(fb).read()
    open(os.path.join(out, "foreign_pipeline.txt"), "w").write(
        "per-kernel times (ncu launch list, cold-cache and serialised) of one foreign decode of 8 x 4096^2 RGBA photo streams\n" + res + extra)
for f in ("launches.csv", "bench.json", "bench_ref.json", "bench_dist0.json", "bench_dist2.json", "bench_dist6.json", "bench_dist7.json"):
    p = os.path.join(ROOT, "gpurun_out", f)
    if os.path.exists(p) and os.path.getsize(p):
        shutil.copy(p, os.path.join(out, f))
subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_traffic.py"), os.path.join(ROOT, "gpurun_out", "prof_enc.ncu-rep"),
                os.path.join(ROOT, "gpurun_out", "prof_dec.ncu-rep"), cmd])
print("wrote", out)
