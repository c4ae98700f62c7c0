#!/usr/bin/env python3
"""Source lines of a kernel ranked by executed warp instructions, with their share of stall samples and the dominant stall reason.
usage: ncu_toplines.py report.ncu-rep [N]"""
import collections
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
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = None
    per = collections.OrderedDict()
    for r in rows:
        if r and r[0] == "Line No":
            hdr = r
            il, isrc, iex, ismp = 0, 1, hdr.index("Instructions Executed"), hdr.index("# Samples")
            stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
            continue
        if hdr and len(r) > ismp and r[il].strip().isdigit():
            ln = int(r[il])
            e = per.setdefault(ln, {"src": r[isrc], "ex": 0, "smp": 0, "st": collections.Counter()})
            e["ex"] += num(r[iex]); e["smp"] += num(r[ismp])
            for i, h in stall_cols:
                e["st"][h] += num(r[i])
    tot = sum(e["ex"] for e in per.values()) or 1
    ts = sum(e["smp"] for e in per.values()) or 1
    print("total executed warp instructions %d, samples %d" % (tot, ts))
    for ln, e in sorted(per.items(), key=lambda kv: -kv[1]["ex"])[:n]:
        top = ", ".join("%s %.0f%%" % (k.replace("stall_", ""), 100.0 * v / max(e["smp"], 1)) for k, v in e["st"].most_common(2))
        print("%5d %5.1f%% instr %5.1f%% samples  [%s]  %s" % (ln, 100.0 * e["ex"] / tot, 100.0 * e["smp"] / ts, top, e["src"].strip()[:100]))


if __name__ == "__main__":
    main()
