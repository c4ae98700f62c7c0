#!/usr/bin/env python3
"""Summarise an .ncu-rep (raw page) into the handful of numbers the roofline discussion needs.
usage: ncu_summary.py report.ncu-rep [source]   ("source" adds the opcode mix and hottest code regions)"""
import collections
import csv
import io
import re
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "lts__t_bytes.sum", "l1tex__data_bank_conflicts_pipe_lsu.sum"]


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    return {h: (v, u) for h, u, v in zip(hdr, units, vals)}


def main():
    rep = sys.argv[1]
    m = raw(rep)
    print("kernel:", m.get("Kernel Name", ("?", ""))[0])
    for k in KEYS:
        if k in m:
            print("  %-70s %s %s" % (k, m[k][0], m[k][1]))
    for k, (v, u) in sorted(m.items()):
        if k.startswith("smsp__average_warps_issue_stalled") and k.endswith("per_issue_active.ratio"):
            try:
                if float(v) >= 0.2:
                    print("  stall %-64s %s" % (k.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), v))
            except ValueError:
                pass
    if len(sys.argv) > 2:
        out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(out)))
        hdr, data = rows[1], rows[2:]
        isrc, iex, ith = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed")
        tot = sum(int(r[iex]) for r in data if r[iex].isdigit())
        h = collections.Counter()
        for r in data:
            if r[iex].isdigit():
                op = re.sub(r"^@!?U?P\w+\s+", "", r[isrc].strip()).split()[0].split(".")[0]
                h[op] += int(r[iex])
        print("  opcode mix (share of %d executed warp instructions):" % tot)
        print("   ", ", ".join("%s %.1f%%" % (op, 100.0 * c / tot) for op, c in h.most_common(18)))


if __name__ == "__main__":
    main()
