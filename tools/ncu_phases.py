#!/usr/bin/env python3
"""Executed warp instructions and stall samples of a kernel grouped by source-line ranges ("phases").
usage: ncu_phases.py report.ncu-rep name:lo-hi[,lo-hi...] ...   (lines not covered are reported as "other")"""
import csv
import io
import subprocess
import sys


def num(s):
    try:
        return int(s)
    except ValueError:
        return 0


def main():
    out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    phases = []
    for a in sys.argv[2:]:
        name, spec = a.split(":")
        phases.append((name, [tuple(int(v) for v in r.split("-")) for r in spec.split(",")]))
    hdr = None
    acc = {}
    for r in rows:
        if r and r[0] == "Line No":
            hdr = r
            iline, iex, ismp = hdr.index("Line No"), hdr.index("Instructions Executed"), hdr.index("# Samples")
            continue
        if hdr and len(r) > ismp and r[iline].strip().isdigit() and num(r[iex]) > 0:
            ln = int(r[iline])
            name = "other"
            for n, rs in phases:
                if any(lo <= ln <= hi for lo, hi in rs):
                    name = n
                    break
            e = acc.setdefault(name, [0, 0])
            e[0] += num(r[iex]); e[1] += num(r[ismp])
    te = sum(v[0] for v in acc.values()) or 1
    ts = sum(v[1] for v in acc.values()) or 1
    for n, (e, s) in sorted(acc.items(), key=lambda kv: -kv[1][0]):
        print("%-12s %6.1f%% of executed instructions (%.0f M)   %6.1f%% of samples" % (n, 100.0 * e / te, e / 1e6, 100.0 * s / ts))


if __name__ == "__main__":
    main()
