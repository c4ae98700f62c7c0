#!/usr/bin/env python3
"""Hottest CUDA source lines of a kernel from an .ncu-rep (needs -lineinfo).  usage: ncu_lines.py report.ncu-rep [top]"""
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
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    hdr = None
    lines = []
    for r in rows:
        if r and r[0] == "Line No":
            hdr = r
            iline, isrc, iex, ismp, ith = hdr.index("Line No"), hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index(
                "Thread Instructions Executed")
            continue
        if hdr and len(r) > ith and r[iline].strip().isdigit() and num(r[iex]) > 0:
            lines.append((num(r[iex]), num(r[ismp]), num(r[ith]), int(r[iline]), r[isrc].strip()[:120]))
    tot = sum(l[0] for l in lines) or 1
    tots = sum(l[1] for l in lines) or 1
    print("executed warp instructions attributed to source lines: %d, samples %d" % (tot, tots))
    for ex, sm, th, ln, src in sorted(lines, reverse=True)[:top]:
        print("%5.1f%% exec %5.1f%% smp  thr %4.1f  L%-4d %s" % (100.0 * ex / tot, 100.0 * sm / tots, th / ex, ln, src))


if __name__ == "__main__":
    main()
