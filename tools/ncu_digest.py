#!/usr/bin/env python
"""ncu_digest.py <report.ncu-rep> [> profiles/<round>/<name>.txt]

Text digest of one `ncu --set full --import-source on` capture, for the summaries committed under profiles/:
launch + headline counters, stall reasons per issue, pipe utilisation, executed-instruction histogram by opcode and
the SASS instructions / source lines that collect the stall samples.  Needs ncu (reads the report; no GPU)."""
import collections
import csv
import io
import subprocess
import sys

WANT = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.per_cycle_active", "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "launch__occupancy_limit_warps", "launch__waves_per_multiprocessor", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__cycles_elapsed.max"]
PIPES = ["adu", "alu", "fma", "lsu", "xu", "cbu_pred_on_any", "uniform"]


def run(args):
    return subprocess.run(["ncu"] + args, capture_output=True, text=True).stdout


def main():
    rep = sys.argv[1]
    rows = list(csv.reader(io.StringIO(run(["-i", rep, "--page", "raw", "--csv"]))))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        u = dict(zip(hdr, units))
        for w in WANT:
            if w in d:
                print(f"{w} = {d[w]} {u[w]}")
        st = [h for h in hdr if "issue_stalled" in h and h.endswith("_per_issue_active.ratio")]
        for v, h in sorted([(float(d[h] or 0), h) for h in st], reverse=True)[:8]:
            print("  stall %.2f %s" % (v, h.split("issue_stalled_")[1]))
        for p in PIPES:
            for suffix in (".avg.pct_of_peak_sustained_active", ".avg.pct_of_peak_sustained_elapsed"):
                k = "sm__inst_executed_pipe_" + p + suffix
                if k in d:
                    print(f"  pipe {p:16s} {float(d[k]):6.1f} % of peak")
                    break
    src = list(csv.reader(io.StringIO(run(["-i", rep, "--page", "source", "--csv", "--print-source", "sass"]))))
    if len(src) < 3:
        return
    h2 = src[1]
    ix = {h: i for i, h in enumerate(h2)}
    data = [r for r in src[2:] if len(r) > ix["Instructions Executed"]]
    cnt, smp = collections.Counter(), collections.Counter()
    for r in data:
        toks = r[ix["Source"]].split()
        if not toks:
            continue
        op = (toks[1] if toks[0].startswith("@") and len(toks) > 1 else toks[0]).split(".")[0]
        cnt[op] += int(r[ix["Instructions Executed"]] or 0)
        smp[op] += int(r[ix["# Samples"]] or 0)
    tot, tots = sum(cnt.values()) or 1, sum(smp.values()) or 1
    print(f"\nexecuted warp instructions by opcode (total {tot / 1e6:.1f} M, {tots} stall samples)")
    for op, c in cnt.most_common(24):
        print("  %-10s %9.1f M %5.1f %%   samples %5.1f %%" % (op, c / 1e6, 100.0 * c / tot, 100.0 * smp[op] / tots))
    print("\nSASS instructions with the most stall samples")
    for r in sorted(data, key=lambda r: -int(r[ix["# Samples"]] or 0))[:16]:
        print("  %5.1f %%s %5.2f %%i  %s" % (100.0 * int(r[ix["# Samples"]] or 0) / tots, 100.0 * int(r[ix["Instructions Executed"]] or 0) / tot,
                                          " ".join(r[ix["Source"]].split())[:80]))
    # per CUDA source line
    src = list(csv.reader(io.StringIO(run(["-i", rep, "--page", "source", "--csv", "--print-source", "cuda"]))))
    if len(src) < 3:
        return
    # the cuda view prints one block per file: "File Path" header rows interleave; keep rows that parse
    h3 = None
    lines = []
    for r in src:
        if "Source" in r and "# Samples" in r:
            h3 = {h: i for i, h in enumerate(r)}
            continue
        if h3 and len(r) > h3["# Samples"]:
            try:
                lines.append((int(r[h3["# Samples"]] or 0), int(r[h3["Instructions Executed"]] or 0), r[h3.get("#", 0)], " ".join(r[h3["Source"]].split())[:90]))
            except ValueError:
                pass
    if lines:
        ts, ti = sum(l[0] for l in lines) or 1, sum(l[1] for l in lines) or 1
        print("\nCUDA source lines with the most stall samples (s) / executed instructions (i)")
        for s_, i_, no, text in sorted(lines, reverse=True)[:22]:
            print("  %5.1f %%s %5.1f %%i  L%-5s %s" % (100.0 * s_ / ts, 100.0 * i_ / ti, no, text))


if __name__ == "__main__":
    main()
